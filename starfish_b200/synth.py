"""Seeded synthetic workload of SURVEY.md §8d (there is no network for PHOENIX libraries).

Pure numpy, no Starfish objects: the tests and the golden-fixture generator feed the *same* arrays to
the reference classes, to the CPU checker and to this package, so "identical inputs" is literal.
"""
from __future__ import annotations

import numpy as np

from .constants import c_kms

PARAM_NAMES = ["T", "logg", "Z"]


def log_uniform_wave(n_pix, lo=5000.0, hi=5600.0):
    """``wave_i = lo·(hi/lo)^(i/(N−1))`` — strictly increasing log-λ grid."""
    return lo * (hi / lo) ** (np.arange(n_pix) / (n_pix - 1))


def multi_order_wave(n_orders=8, n_per_order=2048, start=5000.0, width=70.0, gap=5.0):
    """Config 5: consecutive non-overlapping log-uniform windows, concatenated (strictly increasing)."""
    parts = []
    for o in range(n_orders):
        a = start + o * (width + gap)
        parts.append(log_uniform_wave(n_per_order, a, a + width))
    return np.concatenate(parts)


def make_data(n_pix, wave=None, seed=0):
    """(wave, data_flux, sigma): σ_i = 0.01, flux = 1 + 0.1 sin(λ/7) + 0.01·N(0,1)."""
    wave = log_uniform_wave(n_pix) if wave is None else np.asarray(wave, dtype=np.float64)
    rng = np.random.default_rng(seed)
    flux = 1.0 + 0.1 * np.sin(wave / 7.0) + 0.01 * rng.standard_normal(wave.size)
    sigma = np.full(wave.size, 0.01)
    return wave, flux, sigma


def make_emulator_arrays(n_comp=6, seed=0, lo=4900.0, hi=5700.0, dv=2.0):
    """Keyword arguments for an ``Emulator`` constructor (reference's or ours): 3×3×3 grid (G=27),
    M orthonormal random eigenspectra on a ``dv`` km/s log-λ grid, unit-normal weights."""
    rng = np.random.default_rng(seed)
    n_wl = int(np.ceil(np.log(hi / lo) / np.log1p(dv / c_kms))) + 1
    wl = lo * (hi / lo) ** (np.arange(n_wl) / (n_wl - 1))
    q, _ = np.linalg.qr(rng.standard_normal((n_wl, n_comp)))
    eigenspectra = np.ascontiguousarray(q.T)
    axes = [np.array([6000.0, 6100.0, 6200.0]), np.array([4.0, 4.5, 5.0]), np.array([-0.5, 0.0, 0.5])]
    grid_points = np.array(np.meshgrid(*axes, indexing="ij")).reshape(3, -1).T.copy()
    weights = rng.standard_normal((grid_points.shape[0], n_comp))
    return dict(
        grid_points=grid_points,
        param_names=list(PARAM_NAMES),
        wavelength=wl,
        weights=weights,
        eigenspectra=eigenspectra,
        w_hat=np.ascontiguousarray(weights.T).ravel(),  # component-major, as get_w_hat orders it
        flux_mean=1.0 + 0.1 * np.sin(wl / 7.0),
        flux_std=0.05 + 0.01 * np.cos(wl / 11.0) ** 2,
        factors=np.ones(grid_points.shape[0]),
    )


def walker_params(b, n_local=2, with_global=True, stress=None):
    """Nested parameter dict of walker ``b`` (``default_rng(1000+b)``), SpectrumModel keyword form.

    ``stress=(amp, ls)`` overrides the global kernel (the ill-conditioned parity set, cond → 7e5).
    """
    rng = np.random.default_rng(1000 + b)
    grid = [rng.uniform(6000.0, 6200.0), rng.uniform(4.0, 5.0), rng.uniform(-0.5, 0.5)]
    p = dict(vsini=5.0, vz=0.0, log_scale=0.0, cheb=[0.01, -0.01])
    z1, z2 = rng.standard_normal(2)
    if with_global:
        if stress is None:
            p["global_cov"] = dict(log_amp=np.log(1e-4) + 0.3 * z1, log_ls=np.log(20.0) + 0.1 * z2)
        else:
            p["global_cov"] = dict(log_amp=np.log(stress[0]), log_ls=np.log(stress[1]))
    if n_local:
        loc = []
        for k in range(n_local):
            za, zs = rng.standard_normal(2)
            loc.append(dict(mu=5100.0 + 200.0 * k, log_amp=np.log(1e-4) + 0.3 * za,
                            log_sigma=np.log(30.0) + 0.1 * zs))
        p["local_cov"] = loc
    return grid, p


def walker_params_orders(b, order_centres, n_local_per_order=2):
    """Config 5: K local kernels inside every order (16 total for 8 orders)."""
    grid, p = walker_params(b, n_local=0)
    rng = np.random.default_rng(5000 + b)
    loc = []
    for c in order_centres:
        for k in range(n_local_per_order):
            za, zs = rng.standard_normal(2)
            loc.append(dict(mu=float(c) + 15.0 * (2 * k - 1), log_amp=np.log(1e-4) + 0.3 * za,
                            log_sigma=np.log(30.0) + 0.1 * zs))
    p["local_cov"] = loc
    return grid, p


def stage_inputs_direct(n_pix, n_walkers, n_comp=6, n_local=2, with_global=True, wave=None, seed=0,
                        first_walker=0):
    """Stage-boundary inputs generated *without* the upstream transforms — for the
    bench and the large-size property tests, where running FFT+splines for 256 walkers on the host
    would only add set-up time.  Shapes/scales match what ``SpectrumModel`` produces:

    X[b] (M×N): smooth orthonormal-ish eigenspectra × flux_std × scale; A[b] = Σ_w⁻¹ (M×M SPD);
    model_flux[b] = data-like continuum; kernel hyper-parameters from ``walker_params``.
    """
    wave, data_flux, sigma = make_data(n_pix, wave=wave, seed=seed)
    N = wave.size
    rng = np.random.default_rng(seed + 7)
    basis = np.empty((n_comp, N))
    for m in range(n_comp):
        basis[m] = np.sin(wave / (3.0 + 1.7 * m) + m) * np.sqrt(2.0 / N)
    flux_std = 0.05 + 0.01 * np.cos(wave / 11.0) ** 2
    B = n_walkers
    X = np.empty((B, n_comp, N)) if n_comp else None
    A = np.empty((B, n_comp, n_comp)) if n_comp else None
    model_flux = np.empty((B, N))
    glob = np.zeros((B, 2))
    nloc = np.zeros(B, dtype=np.int32)
    loc = np.zeros((B, max(n_local, 1), 3))
    for i in range(B):
        b = first_walker + i
        _, p = walker_params(b, n_local=n_local, with_global=with_global)
        wr = np.random.default_rng(9000 + b)
        scale = np.exp(0.05 * wr.standard_normal())
        if n_comp:
            X[i] = basis * flux_std * scale * (1.0 + 0.01 * wr.standard_normal((n_comp, 1)))
            s = 1e-4 * np.exp(0.2 * wr.standard_normal((n_comp, n_comp)))
            wcov = s @ s.T + 1e-4 * np.eye(n_comp)   # Σ_w, SPD
            A[i] = np.linalg.inv(wcov)
        model_flux[i] = scale * (1.0 + 0.1 * np.sin(wave / 7.0 + 0.001 * wr.standard_normal()))
        if with_global:
            glob[i] = np.exp(p["global_cov"]["log_amp"]), np.exp(p["global_cov"]["log_ls"])
        nloc[i] = n_local
        for k in range(n_local):
            kk = p["local_cov"][k]
            loc[i, k] = np.exp(kk["log_amp"]), kk["mu"], np.exp(kk["log_sigma"])
    return dict(wave=wave, sigma=sigma, data_flux=data_flux, X=X, A=A, model_flux=model_flux,
                glob=glob, nloc=nloc, loc=loc)


def stage_inputs_orders(n_walkers, n_orders=8, n_per_order=2048, n_comp=6, n_local_per_order=2, first_walker=0):
    """Config 5 of BASELINE.json: one concatenated multi-order spectrum (8 × 2048 = 16384 px) with
    ``n_local_per_order`` local kernels inside every order (16 per walker)."""
    wave = multi_order_wave(n_orders, n_per_order)
    d = stage_inputs_direct(wave.size, n_walkers, n_comp=n_comp, n_local=0, wave=wave, first_walker=first_walker)
    centres = [wave[o * n_per_order:(o + 1) * n_per_order].mean() for o in range(n_orders)]
    K = n_orders * n_local_per_order
    loc = np.zeros((n_walkers, K, 3))
    for i in range(n_walkers):
        _, p = walker_params_orders(first_walker + i, centres, n_local_per_order)
        for k, kk in enumerate(p["local_cov"]):
            loc[i, k] = np.exp(kk["log_amp"]), kk["mu"], np.exp(kk["log_sigma"])
    d["loc"] = loc
    d["nloc"] = np.full(n_walkers, K, dtype=np.int32)
    return d
