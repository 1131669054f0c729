"""CPU oracle for the stages UPSTREAM of the covariance.  TEST INFRASTRUCTURE ONLY.

numpy/scipy restatement of what ``SpectrumModel.__call__`` does before the rank-M term, on plain arrays,
following the reference's operation order.  Imported by ``tests/`` only; the product never imports it.

=============================  ==========================================================
oracle function                reference lines followed
=============================  ==========================================================
``calculate_dv``               ``Starfish/utils.py:8-22``
``create_log_lam_grid``        ``Starfish/utils.py:44-88``
``resample``                   ``Starfish/transforms.py:11-42`` (scipy FITPACK quintic spline)
``rotational_broaden``         ``Starfish/transforms.py:93-134``
``doppler_shift``              ``Starfish/transforms.py:137-158``
``chebyshev_correct``          ``Starfish/transforms.py:271-304``
``renorm_factor``              ``Starfish/transforms.py:265-268``
``batch_kernel``               ``Starfish/emulator/kernels.py:5-49``
``emulator_predict``           ``Starfish/emulator/emulator.py:382-388``
``model_setup``                ``Starfish/models/spectrum_model.py:149-156``
``model_call``                 ``Starfish/models/spectrum_model.py:287-332``
=============================  ==========================================================

Pinning: ``tests/test_oracle_upstream.py`` checks ``model_call`` against the stage inputs recorded from
inside the unmodified reference's ``__call__`` (``tests/golden/model_*.npz`` and ``upstream_*.npz``, written
by ``oracle/make_golden.py``) and re-runs the live reference whenever ``/root/reference`` exists.

Third-party arithmetic the reference delegates to: scipy ``InterpolatedUnivariateSpline`` (FITPACK
``fpcurf``/``splev``), ``scipy.special.j1`` (cephes), ``numpy.fft`` (pocketfft), ``numpy.linalg.solve``
(LAPACK dgesv), ``scipy.spatial.distance.cdist``; the oracle calls the same routines.
"""
from __future__ import annotations

import numpy as np
from numpy.polynomial.chebyshev import chebval
from scipy.interpolate import InterpolatedUnivariateSpline
from scipy.linalg import block_diag
from scipy.spatial.distance import cdist
from scipy.special import j1

C_KMS = 2.99792458e5  # Starfish/constants.py:7


def calculate_dv(wave):
    wave = np.asarray(wave)
    return C_KMS * np.min(np.diff(wave) / wave[:-1])


def create_log_lam_grid(dv, start, end):
    cdelt_temp = np.log10(dv / C_KMS + 1.0)
    crval1, crvaln = np.log10(start), np.log10(end)
    n = (crvaln - crval1) / cdelt_temp
    naxis1 = 2
    while naxis1 < n:
        naxis1 *= 2
    cdelt1 = (crvaln - crval1) / (naxis1 - 1)
    return 10 ** (crval1 + cdelt1 * np.arange(naxis1))


def resample(wave, flux, new_wave):
    return np.array([InterpolatedUnivariateSpline(wave, fl, k=5)(new_wave) for fl in flux])


def rotational_broaden(wave, flux, vsini):
    if vsini <= 0:
        raise ValueError("vsini must be positive")
    dv = calculate_dv(wave)
    freq = np.fft.rfftfreq(flux.shape[-1], dv)
    flux_ff = np.fft.rfft(flux)
    ub = 2.0 * np.pi * vsini * freq
    ub = ub[1:]
    sb = j1(ub) / ub - 3 * np.cos(ub) / (2 * ub**2) + 3.0 * np.sin(ub) / (2 * ub**3)
    flux_ff *= np.insert(sb, 0, 1.0)
    return np.fft.irfft(flux_ff, n=flux.shape[-1])


def doppler_shift(wave, vz):
    return wave * np.sqrt((C_KMS + vz) / (C_KMS - vz))


def chebyshev_correct(wave, flux, coeffs):
    return flux * chebval(wave / wave.max(), np.asarray(coeffs), tensor=False)


def renorm_factor(wave, flux, reference_flux):
    trapz = getattr(np, "trapezoid", None) or np.trapz
    return trapz(reference_flux, wave) / trapz(flux, wave, axis=-1)


def batch_kernel(X, Z, variances, lengthscales):
    return block_diag(*[v * np.exp(-0.5 * cdist(X / l, Z / l, "sqeuclidean"))
                        for v, l in zip(variances, lengthscales)])


def emulator_predict(grid_points, variances, lengthscales, v11, w_hat, params):
    """(mu[M], cov[M,M]) of the emulator weights at ``params`` (R&W eqs 2.18/2.19, as coded)."""
    params = np.atleast_2d(params)
    v12 = batch_kernel(grid_points, params, variances, lengthscales)
    v22 = batch_kernel(params, params, variances, lengthscales)
    v21 = v12.T
    mu = v21 @ np.linalg.solve(v11, w_hat)
    cov = v22 - v21 @ np.linalg.solve(v11, v12)
    return mu, cov


def model_setup(emu_wl, emu_bulk_fluxes, data_wave):
    """(min_dv_wave, bulk_fluxes on it) as the constructor prepares them."""
    dv = calculate_dv(data_wave)
    fine = create_log_lam_grid(dv, emu_wl.min(), emu_wl.max())
    return fine, resample(emu_wl, emu_bulk_fluxes, fine)


def ccm89(wave, a_v, r_v=3.1):
    """A_λ of Cardelli, Clayton & Mathis (1989) — what ``extinction.ccm89(wave, a_v, r_v)`` returns and
    Starfish/transforms.py:199-205 multiplies in.  The ``extinction`` package (v0.4.x, a dependency of the reference,
    setup.py) is ABSENT from this image, so this restates the published law (eqs. 2a-5b of the paper, Horner
    evaluation as in the package's C source): PARITY UNPINNED against the package itself; pinned against the paper's
    Table 3 (tests/test_oracle_upstream.py::test_ccm89_matches_the_papers_table3)."""
    x = 1e4 / np.asarray(wave, dtype=np.float64)
    a = np.zeros_like(x)
    b = np.zeros_like(x)
    for i, xi in enumerate(x):          # plain per-pixel restatement, deliberately not sharing code with the product
        if 0.3 <= xi < 1.1:
            a[i], b[i] = 0.574 * xi**1.61, -0.527 * xi**1.61
        elif xi < 3.3:
            y = xi - 1.82
            a[i] = ((((((0.32999 * y - 0.77530) * y + 0.01979) * y + 0.72085) * y - 0.02427) * y - 0.50447) * y
                    + 0.17699) * y + 1.0
            b[i] = ((((((-2.09002 * y + 5.30260) * y - 0.62251) * y - 5.38434) * y + 1.07233) * y + 2.28305) * y
                    + 1.41338) * y
        elif xi < 8.0:
            d = xi - 5.9 if xi >= 5.9 else 0.0
            a[i] = 1.752 - 0.316 * xi - 0.104 / ((xi - 4.67) ** 2 + 0.341) - 0.04473 * d**2 - 0.009779 * d**3
            b[i] = -3.090 + 1.825 * xi + 1.206 / ((xi - 4.62) ** 2 + 0.263) + 0.2130 * d**2 + 0.1207 * d**3
        elif xi <= 11.0:
            z = xi - 8.0
            a[i] = -1.073 - 0.628 * z + 0.137 * z**2 - 0.070 * z**3
            b[i] = 13.670 + 4.257 * z - 0.420 * z**2 + 0.374 * z**3
        else:
            raise ValueError("ccm89: wavelength out of range")
    return a_v * (a + b / r_v)


def model_call(fine_wave, bulk_fluxes, data_wave, data_flux, weights, *, vsini=None, vz=None, cheb=None,
               log_scale=None, norm=1.0, Av=None):
    """-> (flux[N], X[M,N], log_scale) for one walker; ``None`` = parameter absent from the model."""
    wave, fluxes = fine_wave, bulk_fluxes
    if vsini is not None:
        fluxes = rotational_broaden(wave, fluxes, vsini)
    if vz is not None:
        wave = doppler_shift(wave, vz)
    fluxes = resample(wave, fluxes, data_wave)
    if Av is not None:                                   # spectrum_model.py:298-299, transforms.py:199-205
        fluxes = fluxes * 10 ** (-0.4 * ccm89(data_wave, Av))
    if cheb is not None:
        fluxes = chebyshev_correct(data_wave, fluxes, [1, *cheb])
    *eigenspectra, flux_mean, flux_std = fluxes
    X = eigenspectra * flux_std
    flux = weights @ X + flux_mean
    if log_scale is None:
        scale = renorm_factor(data_wave, flux * norm, data_flux)
        log_scale = np.log(scale)
        scale = scale * norm
    else:
        scale = np.exp(log_scale) * norm
    return flux * scale, X * scale, log_scale
