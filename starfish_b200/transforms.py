"""Host-side spectral transforms upstream of the hot path ("next").

These produce the per-walker inputs of the GPU stage boundary (``X[M,N]``, ``model_flux[N]``).
They mirror the behaviour of Starfish/transforms.py (function names, argument meaning, errors):
``resample`` :11-42, ``instrumental_broaden`` :45-90, ``rotational_broaden`` :93-134,
``doppler_shift`` :137-158, ``rescale`` :209-230, ``renorm`` :233-262, ``chebyshev_correct`` :271-304.
``extinct`` :161-206 for the two closed-form laws (ccm89, odonnell94); the reference delegates to the
``extinction`` package (absent from this image) — its spline-based laws (fitzpatrick99, fm07) and calzetti00 are
not provided.
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


def _ccm89_ab(x, optical="ccm89"):
    """a(x), b(x) of Cardelli, Clayton & Mathis (1989, ApJ 345, 245) eqs. 2-5, x = 1/λ [µm⁻¹] in [0.3, 11];
    ``optical="odonnell94"`` swaps in O'Donnell's (1994, ApJ 422, 158) optical/NIR polynomials (1.1 <= x < 3.3)."""
    x = np.asarray(x, dtype=np.float64)
    a = np.empty_like(x)
    b = np.empty_like(x)
    if np.any((x < 0.3) | (x > 11.0)):
        raise ValueError("ccm89 is defined for 0.3 <= 1/λ[µm] <= 11 (909 Å … 33333 Å)")
    ir = x < 1.1
    a[ir] = 0.574 * x[ir] ** 1.61
    b[ir] = -0.527 * x[ir] ** 1.61
    op = (x >= 1.1) & (x < 3.3)
    y = x[op] - 1.82
    if optical == "odonnell94":
        ca = [1.0, 0.104, -0.609, 0.701, 1.137, -1.718, -0.827, 1.647, -0.505]
        cb = [0.0, 1.952, 2.908, -3.989, -7.985, 11.102, 5.491, -10.805, 3.347]
    else:
        ca = [1.0, 0.17699, -0.50447, -0.02427, 0.72085, 0.01979, -0.77530, 0.32999]
        cb = [0.0, 1.41338, 2.28305, 1.07233, -5.38434, -0.62251, 5.30260, -2.09002]
    pa = np.zeros_like(y)
    pb = np.zeros_like(y)
    for c_a, c_b in zip(ca[::-1], cb[::-1]):   # Horner
        pa = pa * y + c_a
        pb = pb * y + c_b
    a[op], b[op] = pa, pb
    uv = (x >= 3.3) & (x < 8.0)
    xu = x[uv]
    d = np.where(xu >= 5.9, xu - 5.9, 0.0)
    a[uv] = 1.752 - 0.316 * xu - 0.104 / ((xu - 4.67) ** 2 + 0.341) + (-0.04473 * d**2 - 0.009779 * d**3)
    b[uv] = -3.090 + 1.825 * xu + 1.206 / ((xu - 4.62) ** 2 + 0.263) + (0.2130 * d**2 + 0.1207 * d**3)
    fuv = x >= 8.0
    z = x[fuv] - 8.0
    a[fuv] = -1.073 - 0.628 * z + 0.137 * z**2 - 0.070 * z**3
    b[fuv] = 13.670 + 4.257 * z - 0.420 * z**2 + 0.374 * z**3
    return a, b


def extinct(wave, flux, Av, Rv=3.1, law="ccm89"):
    """``flux · 10^(−0.4·A_λ)``, ``A_λ = Av·(a(x) + b(x)/Rv)`` (Starfish/transforms.py:161-206).  The reference
    evaluates the law with the ``extinction`` package; here the two closed-form laws are restated from the papers."""
    if law not in ["ccm89", "odonnell94", "calzetti00", "fitzpatrick99", "fm07"]:
        raise ValueError("Invalid extinction law given")
    if Rv <= 0:
        raise ValueError("Rv must be positive")
    if law not in ("ccm89", "odonnell94"):
        raise NotImplementedError(f"extinction law {law!r} needs the `extinction` package (absent from this image)")
    a, b = _ccm89_ab(1e4 / np.asarray(wave, dtype=np.float64), optical=law)
    return flux * 10 ** (-0.4 * (Av * (a + b / Rv)))


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
