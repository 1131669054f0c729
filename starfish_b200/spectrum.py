"""In-memory spectrum containers exposing what the likelihood path reads from the reference's
``Starfish.spectrum`` objects: ``Spectrum(waves, fluxes, sigmas, masks, name)`` indexable into per-order
views with masked ``wave`` / ``flux`` / ``sigma`` and the raw ``_wave`` / ``_flux`` / ``_sigma`` / ``mask``
arrays that callers (and the reference's tests, tests/test_models/test_models.py:226) overwrite in place.
Defaults follow Starfish/spectrum.py:36 (order σ = 0) and :103 (spectrum σ = 1).  HDF5 load/save is out of
scope (h5py is not in the image)."""
import numpy as np


class Order:
    """One echelle order: full arrays plus a boolean pixel mask applied on read."""

    __slots__ = ("_wave", "_flux", "_sigma", "mask")

    def __init__(self, _wave, _flux, _sigma=None, mask=None):
        self._wave = _wave
        self._flux = _flux
        self._sigma = np.zeros_like(_flux) if _sigma is None else _sigma
        self.mask = np.ones_like(_wave, dtype=bool) if mask is None else mask

    def _masked(self, arr):
        return arr[self.mask]

    wave = property(lambda self: self._masked(self._wave), doc="masked wavelengths")
    flux = property(lambda self: self._masked(self._flux), doc="masked fluxes")
    sigma = property(lambda self: self._masked(self._sigma), doc="masked flux uncertainties")

    def __len__(self):
        return len(self._wave)


class Spectrum:
    """A rectangular stack of orders (1-D input becomes a single order)."""

    def __init__(self, waves, fluxes, sigmas=None, masks=None, name="Spectrum"):
        stack = [np.atleast_2d(waves), np.atleast_2d(fluxes)]
        stack.append(np.ones_like(stack[1]) if sigmas is None else np.atleast_2d(sigmas))
        stack.append(np.ones_like(stack[0], dtype=bool) if masks is None else np.atleast_2d(masks).astype(bool))
        if len({a.shape for a in stack}) != 1:
            raise AssertionError("wave, flux, sigma and mask arrays must share one shape")
        self.orders = [Order(*rows) for rows in zip(*stack)]
        self.name = name

    def __len__(self):
        return len(self.orders)

    def __getitem__(self, index):
        return self.orders[index]

    def __setitem__(self, index, order):
        if len(order) != len(self.orders[0]):
            raise ValueError("Invalid order length; no ragged spectra allowed")
        self.orders[index] = order

    def _stacked(self, attr):
        return np.array([getattr(o, attr) for o in self.orders])

    waves = property(lambda self: self._stacked("wave"))
    fluxes = property(lambda self: self._stacked("flux"))
    sigmas = property(lambda self: self._stacked("sigma"))
    masks = property(lambda self: np.array([o.mask for o in self.orders]))

    @property
    def shape(self):
        return (len(self.orders), len(self.orders[0]))

    def reshape(self, shape):
        """New Spectrum with every array reshaped to ``shape`` (orders × pixels)."""
        raw = [np.array([getattr(o, a) for o in self.orders]).reshape(shape) for a in ("_wave", "_flux", "_sigma")]
        return Spectrum(raw[0], raw[1], raw[2], self.masks.reshape(shape), name=self.name)
