"""Function seam of the hot path: drop-in replacements for the two covariance-kernel builders the
reference imports by name (Starfish/models/spectrum_model.py:23, Starfish/models/utils.py:9).

* ``global_covariance_matrix(wave, amplitude, lengthscale)`` — Starfish/models/kernels.py:7-41
* ``local_covariance_matrix(wave, amplitude, mu, sigma)``   — Starfish/models/kernels.py:44-81

Same arguments, same dense ``[N, N]`` numpy result; the matrix is produced by the fused sm_100a build
kernel through ``sfb_build_cov`` (pass ``as_tensor=True`` to keep it on the GPU as a torch tensor).
There is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .engine import LikelihoodEngine

_ENGINES = {}


def _engine_for(n_pix: int, device: int = 0) -> LikelihoodEngine:
    key = (int(n_pix), int(device))
    eng = _ENGINES.get(key)
    if eng is None:
        if len(_ENGINES) >= 4:  # handles own GPU workspace; keep only a few sizes alive
            _ENGINES.pop(next(iter(_ENGINES))).close()
        eng = LikelihoodEngine(n_pix, 0, 32, 1, device=device, workspace_walkers=-1)  # build-only: no N×N slots
        _ENGINES[key] = eng
    return eng


def _build(wave, glob, loc, as_tensor, device):
    wave = np.ascontiguousarray(wave, dtype=np.float64)
    eng = _engine_for(wave.size, device)
    zeros = np.zeros(wave.size)
    eng.set_data(wave, zeros, zeros)
    C = eng.build_covariance(None, None, glob=glob, loc=loc, n_walkers=1)[0]
    return C if as_tensor else C.cpu().numpy()


def global_covariance_matrix(wave, amplitude: float, lengthscale: float, as_tensor=False, device=0):
    """Matérn-3/2 kernel on the velocity separation of the wavelengths, Hann-tapered to zero at 6ℓ."""
    return _build(wave, np.array([[float(amplitude), float(lengthscale)]]), None, as_tensor, device)


def local_covariance_matrix(wave, amplitude: float, mu: float, sigma: float, as_tensor=False, device=0):
    """Gaussian kernel localised at ``mu`` (σ in km/s), Hann-tapered to zero at 4σ."""
    loc = np.array([[[float(amplitude), float(mu), float(sigma)]]])
    return _build(wave, None, loc, as_tensor, device)
