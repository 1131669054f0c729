"""Generate ``tests/golden/*.npz`` by running the UNMODIFIED reference (through ``ref_loader``).

Run in the build container (needs ``/root/reference``):  ``python -m oracle.make_golden``

The reference's own tests hold no golden vectors for this path, so these files —
outputs of the reference's ``global_covariance_matrix``, ``local_covariance_matrix``,
``SpectrumModel.__call__`` and ``SpectrumModel.log_likelihood`` on the seeded synthetic inputs of
``starfish_b200.synth`` — are what pins both the oracle and the CUDA path.  The per-walker stage
inputs (``X``, ``weights_cov``, ``model_flux``) are captured from inside the reference's ``__call__``
by wrapping the ``cho_factor``/``cho_solve`` names in its module namespace with recorders (the
reference source is not touched).
"""
import os
import sys
import time
import warnings

import numpy as np

_ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if _ROOT not in sys.path:
    sys.path.insert(0, _ROOT)

from oracle import ref_loader  # noqa: E402
from starfish_b200 import synth  # noqa: E402

OUT = os.path.join(_ROOT, "tests", "golden")


class _Recorder:
    """Wraps scipy's cho_factor/cho_solve inside Starfish.models.spectrum_model to capture the
    first factorised matrix (weights_cov) and first right-hand side (X) of one ``__call__``."""

    def __init__(self, module):
        self.m = module
        self.f0, self.s0 = module.cho_factor, module.cho_solve
        self.weights_cov = None
        self.X = None

    def __enter__(self):
        def cho_factor(a, *args, **kw):
            if self.weights_cov is None:
                self.weights_cov = np.array(a, copy=True)
            return self.f0(a, *args, **kw)

        def cho_solve(c, b, *args, **kw):
            if self.X is None:
                self.X = np.array(b, copy=True)
            return self.s0(c, b, *args, **kw)

        self.m.cho_factor, self.m.cho_solve = cho_factor, cho_solve
        return self

    def __exit__(self, *exc):
        self.m.cho_factor, self.m.cho_solve = self.f0, self.s0


def run_reference_model(n_pix, walker, wave=None, mus=None, stress=None, n_local=2,
                        with_global=True, emu=None):
    """Returns dict of stage inputs + the reference's cov and lnL for one walker."""
    import Starfish.models.spectrum_model as sm

    emu = emu or synth.make_emulator_arrays()
    w, f, s = synth.make_data(n_pix, wave=wave)
    grid, p = synth.walker_params(walker, n_local=n_local, with_global=with_global, stress=stress)
    if mus is not None:
        for k, mu in enumerate(mus):
            p["local_cov"][k]["mu"] = float(mu)
    model = ref_loader.build_reference_model(emu, w, f, s, grid, p)
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        with _Recorder(sm) as rec:
            flux, cov = model()
        t0 = time.perf_counter()
        lnl = model.log_likelihood()
        dt = time.perf_counter() - t0
    out = dict(wave=w, data_flux=f, sigma=s, X=rec.X, weights_cov=rec.weights_cov, model_flux=flux,
               cov=cov, lnL=np.float64(lnl), labels=np.array(model.labels),
               param_vector=model.get_param_vector(), ref_seconds=np.float64(dt))
    if with_global:
        out["glob"] = np.array([np.exp(p["global_cov"]["log_amp"]), np.exp(p["global_cov"]["log_ls"])])
    else:
        out["glob"] = np.zeros(2)
    out["loc"] = np.array([[np.exp(k["log_amp"]), k["mu"], np.exp(k["log_sigma"])]
                           for k in p.get("local_cov", [])]).reshape(-1, 3)
    return out


def main():
    os.makedirs(OUT, exist_ok=True)
    ref_loader.load_reference()
    from Starfish.models.kernels import global_covariance_matrix, local_covariance_matrix

    # ---- 1. kernel functions on a small dense grid (full matrices) ---------------------------
    wave = synth.log_uniform_wave(192, 5096.0, 5104.0)  # ~2.4 km/s pixels
    g_params = np.array([[1e-4, 20.0], [3.0, 5.0], [100.0, 1.0], [1e-2, 40.0], [7.5e-5, 80.0]])
    l_params = np.array([[1e-4, 5100.0, 30.0], [2e-3, 5097.5, 12.0], [5.0, 5103.9, 50.0],
                         [1e-4, 5200.0, 30.0]])  # last one: centre outside the grid -> all zero
    np.savez_compressed(
        os.path.join(OUT, "kernels_n192.npz"), wave=wave, g_params=g_params, l_params=l_params,
        g=np.array([global_covariance_matrix(wave, a, l) for a, l in g_params]),
        l=np.array([local_covariance_matrix(wave, a, m, s) for a, m, s in l_params]))

    # the reference's own test grid (tests/test_models/test_kernels.py:9-38): 1000 px, 1e4..2e4 Å
    wt = np.linspace(1e4, 2e4, 1000)
    gt = global_covariance_matrix(wt, 100.0, 1.0)
    lt = local_covariance_matrix(wt, 100.0, 1.5e4, 1e3)
    np.savez_compressed(os.path.join(OUT, "kernels_reftest.npz"), g_diag=gt.diagonal().copy(),
                        g_offdiag_absmax=np.abs(gt - np.diag(gt.diagonal())).max(),
                        l_rows=lt[495:505].copy(), l_sum=lt.sum(), l_nnz=np.count_nonzero(lt))

    # ---- 2. whole model, small N, full covariance ---------------------------------------------
    wsmall = synth.log_uniform_wave(256, 5092.0, 5108.0)
    for walker in (0, 1):
        r = run_reference_model(256, walker, wave=wsmall, mus=(5098.0, 5103.0))
        np.savez_compressed(os.path.join(OUT, f"model_n256_w{walker}.npz"), **r)
        print("n256", walker, r["lnL"])

    # ---- 3. config 1 (N=2048, K=2) + the ill-conditioned stress set ---------------------------
    rows = np.array([0, 1, 63, 64, 127, 128, 339, 340, 341, 700, 1023, 1024, 1025, 1663, 2046, 2047])
    for walker in (0, 3):
        r = run_reference_model(2048, walker)
        cov = r.pop("cov")
        r["cov_rows_idx"] = rows
        r["cov_rows"] = cov[rows].copy()
        r["cov_diag"] = cov.diagonal().copy()
        np.savez_compressed(os.path.join(OUT, f"model_n2048_w{walker}.npz"), **r)
        print("n2048", walker, r["lnL"], "ref s", r["ref_seconds"])
    stress = []
    for amp in (1e-2, 1.0, 1e2):
        for ls in (40.0, 60.0, 80.0):
            r = run_reference_model(2048, 0, stress=(amp, ls))
            stress.append((amp, ls, float(r["lnL"])))
            print("stress", stress[-1])
    base = run_reference_model(2048, 0)
    np.savez_compressed(os.path.join(OUT, "stress_n2048.npz"), stress=np.array(stress),
                        **{k: base[k] for k in ("wave", "data_flux", "sigma", "X", "weights_cov",
                                                "model_flux", "loc")})

    # ---- 4. config 2 shape (global only, no emulator term is not expressible through the
    #         reference's SpectrumModel, which always adds XᵀAX) -> kernel-level golden instead
    w4 = synth.log_uniform_wave(4096)
    g4 = global_covariance_matrix(w4, 1.3e-4, 21.0)
    np.savez_compressed(os.path.join(OUT, "global_n4096_rows.npz"), params=np.array([1.3e-4, 21.0]),
                        rows_idx=rows * 2, rows=g4[rows * 2].copy())

    # ---- 5. larger sizes: inputs + lnL only ----------------------------------------------------
    for n in (4096,):
        r = run_reference_model(n, 5)
        r.pop("cov")
        np.savez_compressed(os.path.join(OUT, f"model_n{n}_w5.npz"), **r)
        print(f"n{n}", r["lnL"], "ref s", r["ref_seconds"])


UPSTREAM_VARIANTS = {
    # name: (n_pix, wave window, walker, overrides; None removes the parameter from the model)
    "a": (256, (5092.0, 5108.0), 2, dict(vsini=12.5, vz=-37.0, log_scale=None, cheb=[0.02, -0.01, 0.005])),
    "b": (256, (5092.0, 5108.0), 4, dict(vsini=None, vz=None, log_scale=-0.2, cheb=None)),
    "c": (256, (5092.0, 5108.0), 6, dict(vsini=None, vz=55.0, cheb=[0.03])),
    "d": (300, (5400.0, 5690.0), 7, dict(vsini=31.0, vz=120.0)),                 # coarse pixels up to the grid edge
    "e": (2048, (5000.0, 5070.0), 8, dict(vsini=2.5, vz=-8.0, log_scale=None)),  # 2 km/s pixels -> nf = 32768
}


def upstream_variant_params(name):
    n_pix, window, walker, over = UPSTREAM_VARIANTS[name]
    grid, p = synth.walker_params(walker)
    mid = 0.5 * (window[0] + window[1])
    half = 0.5 * (window[1] - window[0])
    for k, kern in enumerate(p["local_cov"]):
        kern["mu"] = float(mid + (0.4 * k - 0.2) * half)
    for key, val in over.items():
        if val is None:
            p.pop(key, None)
        else:
            p[key] = val
    return n_pix, synth.log_uniform_wave(n_pix, *window), grid, p


def main_upstream():
    """Stage inputs recorded inside the reference's __call__ for parameter sets that exercise every
    upstream transform (rows f1/f2): Doppler shift, strong/weak/no rotation, renormalisation, Chebyshev
    orders, and a fine-pixel case whose internal grid has 32768 points."""
    os.makedirs(OUT, exist_ok=True)
    ref_loader.load_reference()
    import Starfish.models.spectrum_model as sm

    for name in UPSTREAM_VARIANTS:
        n_pix, wave, grid, p = upstream_variant_params(name)
        w, f, s = synth.make_data(n_pix, wave=wave)
        model = ref_loader.build_reference_model(synth.make_emulator_arrays(), w, f, s, grid, p)
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            with _Recorder(sm) as rec:
                flux, cov = model()
            weights, wcov = model.emulator(model.grid_params)
            lnl = model.log_likelihood()
        np.savez_compressed(
            os.path.join(OUT, f"upstream_{name}.npz"), wave=w, data_flux=f, sigma=s, X=rec.X,
            weights=weights, weights_cov=rec.weights_cov, model_flux=flux, lnL=np.float64(lnl),
            log_scale=np.float64(model._log_scale), n_fine=np.int64(len(model.min_dv_wave)),
            cov_diag=cov.diagonal().copy(), labels=np.array(model.labels), param_vector=model.get_param_vector())
        print("upstream", name, n_pix, len(model.min_dv_wave), float(lnl), float(model._log_scale))


if __name__ == "__main__":
    if "upstream" in sys.argv[1:]:
        main_upstream()
    else:
        main()
