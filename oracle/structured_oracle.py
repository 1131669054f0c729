"""Independent large-N checker for the log-likelihood (TEST INFRASTRUCTURE ONLY).

The dense oracle needs O(N³) time and ~10 N² bytes — fine up to N≈4096, too slow to run many cases at the
benchmark sizes (N=8192, 16384).  This module evaluates the SAME quantity

    lnL = −½ [ log det(C) + Rᵀ C⁻¹ R ],   C = S + XᵀAX,   S = diag(σ²+jitter) + K_global + Σ K_local

through the structure of C: on a sorted wavelength grid S is banded (the Matérn taper
and the local blocks are compactly supported), so  S = L_b L_bᵀ  by banded Cholesky (LAPACK dpbtrf via
scipy), and the rank-M term enters through the matrix determinant lemma and the Woodbury identity:

    log det C = log det S + log det(I + A·G),          G = X S⁻¹ Xᵀ   (M×M)
    Rᵀ C⁻¹ R  = RᵀS⁻¹R − uᵀ (I + A·G)⁻¹ A u,           u = X S⁻¹ R

It shares no code path with the CUDA kernels (no dense factorisation at all) and its kernel entries come
from the dense oracle's formulas evaluated only inside the band.  tests/test_oracle.py checks it against the
dense oracle (itself pinned to the reference) at N ≤ 2048.
"""
import numpy as np
from scipy.linalg import cholesky_banded, cho_solve_banded

from .starfish_oracle import C_KMS, JITTER


def _band_halfwidth(wave, glob, loc):
    """Largest |i−j| with a non-zero kernel entry (wave strictly increasing)."""
    n = wave.size
    hw = 0
    if glob is not None and glob[0] > 0:
        r0 = 6 * glob[1]
        # r(i,j) grows with |i-j|: find, for a coarse set of rows, the farthest column inside r0
        for i in np.unique(np.linspace(0, n - 1, 64).astype(int)):
            r = C_KMS / 2 * np.abs((wave - wave[i]) / (wave + wave[i]))
            idx = np.flatnonzero(r <= r0)
            hw = max(hw, i - idx.min(), idx.max() - i)
        hw += 2
    for amp, mu, sig in loc:
        m = C_KMS / mu * np.abs(wave - mu)
        idx = np.flatnonzero(m <= 4 * sig)
        if idx.size:
            hw = max(hw, idx.max() - idx.min())
    return int(min(hw, n - 1))


def banded_S(wave, sigma, glob, loc, jitter=JITTER):
    """Lower banded storage ab[k, j] = S[j+k, j] (scipy `lower=True` convention) and the half-width."""
    wave = np.asarray(wave, dtype=np.float64)
    n = wave.size
    if np.any(np.diff(wave) <= 0):
        raise ValueError("structured oracle needs a strictly increasing wavelength grid")
    hw = _band_halfwidth(wave, glob, loc)
    ab = np.zeros((hw + 1, n))
    ab[0] = np.asarray(sigma) ** 2
    for k in range(hw + 1):
        wi, wj = wave[k:], wave[: n - k]          # S[j+k, j]: row j+k, column j
        val = np.zeros(n - k)
        if glob is not None and glob[0] > 0:
            amp, ls = glob
            r = C_KMS / 2 * np.abs((wj - wi) / (wj + wi))
            r0 = 6 * ls
            ins = r <= r0
            rin = r[ins]
            val[ins] += ((0.5 + 0.5 * np.cos(np.pi * rin / r0)) * amp * (1 + np.sqrt(3) * rin / ls)
                         * np.exp(-np.sqrt(3) * rin / ls))
        lsum = np.zeros(n - k)
        for amp, mu, sig in loc:
            mi = C_KMS / mu * np.abs(wi - mu)
            mj = C_KMS / mu * np.abs(wj - mu)
            rt = np.maximum(mi, mj)
            r0 = 4 * sig
            ins = rt <= r0
            lsum[ins] += (0.5 + 0.5 * np.cos(np.pi * rt[ins] / r0)) * amp * np.exp(
                -0.5 * (mi[ins] ** 2 + mj[ins] ** 2) / sig**2)
        ab[k, : n - k] += val + lsum
    ab[0] += jitter
    return ab, hw


def stage_log_likelihood(wave, sigma, data_flux, X, A, model_flux, glob=None, loc=()):
    """lnL of the stage boundary; `A` is the M×M matrix of the rank-M term (Σ_w⁻¹ in parity mode)."""
    ab, _ = banded_S(wave, sigma, glob, loc)
    cb = cholesky_banded(ab, lower=True)
    logdet = 2.0 * np.sum(np.log(cb[0]))
    R = np.asarray(model_flux) - np.asarray(data_flux)
    SiR = cho_solve_banded((cb, True), R)
    quad = R @ SiR
    if X is not None:
        X = np.asarray(X)
        A = np.asarray(A)
        SiXt = cho_solve_banded((cb, True), X.T)          # N×M
        G = X @ SiXt                                       # M×M
        K = np.eye(X.shape[0]) + A @ G
        sign, ld = np.linalg.slogdet(K)
        if sign <= 0:
            raise np.linalg.LinAlgError("capacitance matrix not positive definite")
        logdet += ld
        u = X @ SiR
        quad -= u @ np.linalg.solve(K, A @ u)
    return -(logdet + quad) / 2
