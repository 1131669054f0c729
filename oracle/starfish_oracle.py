"""CPU oracle for the Starfish per-step log-likelihood hot path.  TEST INFRASTRUCTURE ONLY.

This file is a numpy/scipy restatement of the reference's algorithm for the path named in
BASELINE.json (``north_star``).  It exists so that the CUDA path can be checked on a GPU box where
``/root/reference`` is not mounted.  Only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it; the product package
``starfish_b200`` never does (``tests/test_no_oracle_in_product.py`` enforces this).

Pinning: the reference ships NO golden vectors for this path (its tests are property checks only,
``tests/test_models/test_kernels.py:9-38``, ``tests/test_models/test_models.py:223-229``).  The
oracle is therefore pinned against outputs of the *reference itself*, imported unmodified in the
build container through ``oracle/ref_loader.py`` and recorded by ``oracle/make_golden.py`` into
``tests/golden/*.npz``; ``tests/test_oracle.py`` replays those fixtures (everywhere) and re-runs the
live comparison whenever ``/root/reference`` is present.

Every function follows the reference's operation ORDER (not just its maths) so results are
bit-comparable with the reference's numpy expressions:

============================  =====================================================
oracle function               reference lines followed
============================  =====================================================
``global_covariance_matrix``  ``Starfish/models/kernels.py:27-40``
``local_covariance_matrix``   ``Starfish/models/kernels.py:70-80``
``emulator_term``             ``Starfish/models/spectrum_model.py:334-335``
``assemble_covariance``       ``Starfish/models/spectrum_model.py:334-363``
``log_likelihood``            ``Starfish/models/spectrum_model.py:399-405``
``C_KMS``                     ``Starfish/constants.py:7``
============================  =====================================================

Third-party arithmetic the reference delegates to (not under /root/reference): scipy
``cho_factor``/``cho_solve`` (LAPACK dpotrf/dpotrs; pinned ``scipy>=1.3.0,<2`` in the reference's
``setup.py:51``; 1.18.x in this image) and numpy elementwise maths (pinned ``numpy>=1.16,<2``;
2.3.x here).  The oracle calls the same routines.
"""
from __future__ import annotations

import numpy as np
from scipy.linalg import cho_factor, cho_solve

C_KMS = 2.99792458e5  # Starfish/constants.py:7
JITTER = 1e-10  # Starfish/models/spectrum_model.py:399


def global_covariance_matrix(wave, amplitude, lengthscale):
    """Matérn-3/2 × Hann taper on the reference's "velocity" metric (kernels.py:27-40).

    r_ij = (c/2)·|(λj−λi)/(λj+λi)|, r0 = 6ℓ, K = [r<=r0]·(½+½cos(πr/r0))·a·(1+√3r/ℓ)·exp(−√3r/ℓ).
    """
    wave = np.asarray(wave, dtype=np.float64)
    col = wave[np.newaxis, :]  # == meshgrid "wx"
    row = wave[:, np.newaxis]  # == meshgrid "wy"
    r = C_KMS / 2 * np.abs((col - row) / (col + row))
    r0 = 6 * lengthscale
    out = np.zeros((wave.size, wave.size))
    inside = r <= r0
    rin = r[inside]
    hann = 0.5 + 0.5 * np.cos(np.pi * rin / r0)
    out[inside] = (
        hann
        * amplitude
        * (1 + np.sqrt(3) * rin / lengthscale)
        * np.exp(-np.sqrt(3) * rin / lengthscale)
    )
    return out


def local_covariance_matrix(wave, amplitude, mu, sigma):
    """Gaussian × Hann taper on the max-metric (kernels.py:70-80).

    m_i = (c/μ)|λi−μ|, r0 = 4σ, K = [max(m_i,m_j)<=r0]·(½+½cos(π·max/r0))·A·exp(−½(m_i²+m_j²)/σ²).
    """
    wave = np.asarray(wave, dtype=np.float64)
    metric = C_KMS / mu * np.abs(wave - mu)
    mx = np.broadcast_to(metric[np.newaxis, :], (wave.size, wave.size))
    my = np.broadcast_to(metric[:, np.newaxis], (wave.size, wave.size))
    r_tap = np.maximum(mx, my)
    r2 = mx**2 + my**2
    r0 = 4 * sigma
    out = np.zeros((wave.size, wave.size))
    inside = r_tap <= r0
    hann = 0.5 + 0.5 * np.cos(np.pi * r_tap[inside] / r0)
    out[inside] = hann * amplitude * np.exp(-0.5 * r2[inside] / sigma**2)
    return out


def emulator_term(X, weights_cov):
    """``Xᵀ Σ_w⁻¹ X`` exactly as coded at spectrum_model.py:334-335 (NOT ΦΣ_wΦᵀ)."""
    fac = cho_factor(np.array(weights_cov, dtype=np.float64), overwrite_a=True)
    return X.T @ cho_solve(fac, X)


def assemble_covariance(wave, sigma, X, weights_cov, global_cov=None, local_cov=()):
    """Covariance returned by ``SpectrumModel.__call__`` (spectrum_model.py:334-363), no jitter.

    global_cov: None or (amplitude, lengthscale) — already exponentiated.
    local_cov : iterable of (amplitude, mu, sigma) — already exponentiated.
    X may be None (config 2: global kernel + σ² only).
    """
    n = len(wave)
    if X is not None:
        cov = emulator_term(X, weights_cov)
    else:
        cov = np.zeros((n, n))
    np.fill_diagonal(cov, cov.diagonal() + np.asarray(sigma) ** 2)
    if global_cov is not None:
        cov += global_covariance_matrix(wave, global_cov[0], global_cov[1])
    if len(local_cov):
        loc = 0
        for amp, mu, sig in local_cov:
            loc = loc + local_covariance_matrix(wave, amp, mu, sig)
        cov += loc
    return cov


def log_likelihood(cov, model_flux, data_flux):
    """lnL of spectrum_model.py:399-405 (no priors, no N·log2π).  Destroys ``cov``.

    Raises numpy.linalg.LinAlgError when cov+1e-10·I is not positive definite, as scipy does.
    Returns (lnL, logdet, sqmah, R).
    """
    np.fill_diagonal(cov, cov.diagonal() + JITTER)
    factor, flag = cho_factor(cov, overwrite_a=True)
    logdet = 2 * np.sum(np.log(factor.diagonal()))
    R = model_flux - data_flux
    sqmah = R @ cho_solve((factor, flag), R)
    return -(logdet + sqmah) / 2, logdet, sqmah, R


def stage_log_likelihood(wave, sigma, data_flux, X, weights_cov, model_flux,
                         global_cov=None, local_cov=()):
    """Whole stage boundary of SURVEY §8d: inputs → lnL (float)."""
    cov = assemble_covariance(wave, sigma, X, weights_cov, global_cov, local_cov)
    return log_likelihood(cov, model_flux, data_flux)[0]
