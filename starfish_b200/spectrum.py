"""In-memory data containers with the attribute surface the likelihood path reads
(``wave``, ``flux``, ``sigma``, ``mask``, ``name``); mirrors Starfish/spectrum.py:8-62, :65-130
without the HDF5 I/O (h5py is absent; I/O is out of scope)."""
from dataclasses import dataclass
from typing import Optional

import numpy as np


@dataclass
class Order:
    _wave: np.ndarray
    _flux: np.ndarray
    _sigma: Optional[np.ndarray] = None
    mask: Optional[np.ndarray] = None

    def __post_init__(self):
        if self._sigma is None:
            self._sigma = np.zeros_like(self._flux)
        if self.mask is None:
            self.mask = np.ones_like(self._wave, dtype=bool)

    @property
    def wave(self):
        return self._wave[self.mask]

    @property
    def flux(self):
        return self._flux[self.mask]

    @property
    def sigma(self):
        return self._sigma[self.mask]

    def __len__(self):
        return len(self._wave)


class Spectrum:
    def __init__(self, waves, fluxes, sigmas=None, masks=None, name="Spectrum"):
        waves = np.atleast_2d(waves)
        fluxes = np.atleast_2d(fluxes)
        sigmas = np.ones_like(fluxes) if sigmas is None else np.atleast_2d(sigmas)
        masks = np.ones_like(waves, dtype=bool) if masks is None else np.atleast_2d(masks).astype(bool)
        if not (fluxes.shape == waves.shape == sigmas.shape == masks.shape):
            raise AssertionError("wave/flux/sigma/mask arrays have incompatible shapes")
        self.orders = [Order(w, f, s, m) for w, f, s, m in zip(waves, fluxes, sigmas, masks)]
        self.name = name

    def __getitem__(self, index):
        return self.orders[index]

    def __setitem__(self, index, order):
        if len(order) != len(self.orders[0]):
            raise ValueError("Invalid order length; no ragged spectra allowed")
        self.orders[index] = order

    def __len__(self):
        return len(self.orders)

    @property
    def waves(self):
        return np.array([o.wave for o in self.orders])

    @property
    def fluxes(self):
        return np.array([o.flux for o in self.orders])

    @property
    def sigmas(self):
        return np.array([o.sigma for o in self.orders])

    @property
    def shape(self):
        return (len(self), len(self.orders[0]))
