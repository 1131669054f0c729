"""Wavelength-grid helpers (host side; mirror Starfish/utils.py:8-22 and :44-88)."""
import numpy as np

from .constants import c_kms


def calculate_dv(wave):
    """Smallest pixel-to-pixel velocity step of ``wave`` in km/s (Starfish/utils.py:8-22)."""
    wave = np.asarray(wave, dtype=np.float64)
    return c_kms * np.min(np.diff(wave) / wave[:-1])


def create_log_lam_grid(dv, start, end):
    """Log-λ grid from ``start`` to ``end`` with a power-of-two length and spacing <= ``dv``.

    Mirrors Starfish/utils.py:44-88 (same FITS-style keys in the returned dict).
    """
    if start >= end:
        raise ValueError("Wavelength must be increasing, but start >= end")
    if start <= 0 or end <= 0:
        raise ValueError("Cannot have negative or 0 wavelength")
    step = np.log10(dv / c_kms + 1.0)
    lo, hi = np.log10(start), np.log10(end)
    needed = (hi - lo) / step
    npix = 2
    while npix < needed:
        npix *= 2
    cdelt = (hi - lo) / (npix - 1)
    wl = 10 ** (lo + cdelt * np.arange(npix))
    return {"wl": wl, "CRVAL1": lo, "CDELT1": cdelt, "NAXIS1": npix}
