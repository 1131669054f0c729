"""Host-side spectral transforms upstream of the hot path ("next").

These produce the per-walker inputs of the GPU stage boundary (``X[M,N]``, ``model_flux[N]``).
They mirror the behaviour of Starfish/transforms.py (function names, argument meaning, errors):
``resample`` :11-42, ``instrumental_broaden`` :45-90, ``rotational_broaden`` :93-134,
``doppler_shift`` :137-158, ``rescale`` :209-230, ``renorm`` :233-262, ``chebyshev_correct`` :271-304.
``extinct`` is not provided: the ``extinction`` package is absent from this image.
"""
import numpy as np
from numpy.polynomial.chebyshev import chebval
from scipy.interpolate import InterpolatedUnivariateSpline
from scipy.special import j1

from .constants import c_kms
from .utils import calculate_dv


def resample(wave, flux, new_wave):
    """Quintic-spline interpolation of ``flux`` (1-D or rows of 2-D) onto ``new_wave``."""
    new_wave = np.asarray(new_wave)
    if np.any(new_wave <= 0):
        raise ValueError("Wavelengths must be positive")
    flux = np.asarray(flux)
    if flux.ndim == 1:
        return InterpolatedUnivariateSpline(wave, flux, k=5)(new_wave)
    return np.array([InterpolatedUnivariateSpline(wave, row, k=5)(new_wave) for row in flux])


def _fourier_filter(wave, flux, transfer):
    dv = calculate_dv(wave)
    npix = flux.shape[-1]
    freq = np.fft.rfftfreq(npix, d=dv)
    spec = np.fft.rfft(flux)
    spec *= transfer(freq)
    return np.fft.irfft(spec, n=npix)


def instrumental_broaden(wave, flux, fwhm):
    """Gaussian line-spread function of the given FWHM [km/s], applied in Fourier space."""
    if fwhm < 0:
        raise ValueError("FWHM must be non-negative")
    sig = fwhm / 2.355
    return _fourier_filter(wave, flux, lambda f: np.exp(-2 * (np.pi * sig * f) ** 2))


def rotational_broaden(wave, flux, vsini):
    """Gray (2005) rotational kernel for ``vsini`` [km/s], applied in Fourier space."""
    if vsini <= 0:
        raise ValueError("vsini must be positive")

    def transfer(freq):
        ub = 2.0 * np.pi * vsini * freq[1:]
        sb = j1(ub) / ub - 3 * np.cos(ub) / (2 * ub**2) + 3.0 * np.sin(ub) / (2 * ub**3)
        return np.insert(sb, 0, 1.0)

    return _fourier_filter(wave, flux, transfer)


def doppler_shift(wave, vz):
    """λ·sqrt((c+vz)/(c−vz))."""
    return wave * np.sqrt((c_kms + vz) / (c_kms - vz))


def rescale(flux, scale):
    """flux·Ω; an array ``scale`` broadcasts over the leading (batch) axis."""
    scale = np.atleast_1d(scale)
    if len(scale) > 1:
        scale = scale[:, np.newaxis]
    return flux * scale


def _get_renorm_factor(wave, flux, reference_flux):
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return trapz(reference_flux, wave) / trapz(flux, wave, axis=-1)


def renorm(wave, flux, reference_flux):
    """Scale ``flux`` so its integral over ``wave`` equals that of ``reference_flux``."""
    return rescale(flux, _get_renorm_factor(wave, flux, reference_flux))


def chebyshev_correct(wave, flux, coeffs):
    """Multiply by a Chebyshev series in λ/λmax; for a single spectrum c0 must be 1."""
    coeffs = np.asarray(coeffs)
    if coeffs.ndim == 1 and coeffs[0] != 1:
        raise ValueError(
            "For single spectrum the linear Chebyshev coefficient (c[0]) must be 1"
        )
    return flux * chebval(wave / wave.max(), coeffs, tensor=False)
