"""Import shim (oracle scaffolding only): lets `import Starfish` succeed without h5py.
Nothing on the log-likelihood path touches HDF5."""


class File:  # pragma: no cover - never instantiated on the path
    def __init__(self, *a, **k):
        raise ImportError("h5py is not installed; this is an import shim for the oracle")


class Group:  # pragma: no cover
    pass


class Dataset:  # pragma: no cover
    pass
