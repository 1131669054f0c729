"""Import shim (oracle scaffolding only): `NDArray[float]` annotations must be subscriptable."""


class _Sub(type):
    def __getitem__(cls, item):
        return cls


class NDArray(metaclass=_Sub):
    pass
