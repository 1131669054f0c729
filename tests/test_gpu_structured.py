"""GPU: the structure-exploiting solver (csrc/band.cu) against the dense CUDA path, the dense
CPU oracle and the independent banded+Woodbury CPU checker.  Same stage boundary, same tolerance as the
dense path: |ΔlnL| <= 1e-10·max(1,|lnL|)."""
import numpy as np
import pytest

from oracle import starfish_oracle as O
from oracle import structured_oracle as S
from starfish_b200 import synth

pytestmark = pytest.mark.gpu

TOL = 1e-10


def _engine(N, M, K, B, solver="structured", **kw):
    from starfish_b200.engine import LikelihoodEngine

    eng = LikelihoodEngine(N, M, K, B, **kw)
    eng.set_solver(solver)
    return eng


def _dense_ref(d, b):
    loc = d["loc"][b][: d["nloc"][b]]
    glob = d["glob"][b] if d["glob"][b][0] > 0 else None
    cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, glob, loc)
    if d["X"] is not None:
        cov += d["X"][b].T @ d["A"][b] @ d["X"][b]
    return O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]


def _run(eng, d, **kw):
    X, A = (d["X"], d["A"]) if d["X"] is not None else (None, None)
    lnL, info = eng.log_likelihood(X, A, d["model_flux"], glob=d["glob"], nloc=d["nloc"], loc=d["loc"], **kw)
    return lnL.cpu().numpy(), info.cpu().numpy()


def test_every_window_class_and_dense_fallback_n2048():
    """ℓ chosen so that the band (half-widths 30, 85, 117, 140, 182, 218, 247 and 291 pixels) needs each of the
    64/96/128/160/192/256-pixel windows and, for the last walker, more than any window (dense path inside
    the same call)."""
    N, B = 2048, 9
    d = synth.stage_inputs_direct(N, B, n_comp=6, n_local=2)
    d["glob"][:, 1] = [20.0, 58.0, 80.0, 96.0, 125.0, 150.0, 170.0, 200.0, 20.0]
    d["glob"][:, 0] = [1e-4, 2e-4, 1e-4, 3e-4, 1e-4, 2e-4, 1e-4, 1e-4, 5e-3]
    eng = _engine(N, 6, 2, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    classes = eng.band_classes()
    assert classes == {64: 2, 96: 1, 128: 1, 160: 1, 192: 1, 256: 2, 0: 1}, classes
    assert (info == 0).all()
    for b in range(B):
        ref = _dense_ref(d, b)
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    # the dense CUDA path on the same handle gives the same numbers
    eng.set_solver("dense")
    lnL_d, info_d = _run(eng, d)
    assert np.abs(lnL - lnL_d).max() <= TOL * np.abs(lnL_d).max()
    # residuals come back identically in both modes
    eng.set_solver("structured")
    out = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], nloc=d["nloc"], loc=d["loc"],
                             return_residuals=True)
    assert np.array_equal(out[2].cpu().numpy(), d["model_flux"] - d["data_flux"])
    eng.close()


def test_band_within_three_pixels_of_every_window():
    """Half-bandwidths WD−4 … WD−1 for WD = 64, 96, 128, 160 (length scales from the support test itself): the band
    fills its window, and for b > WD−4 the rank-4 kernel's entering rows reach pivots of the block that has just
    been eliminated (late-row correction).  Same classes as the rank-1 kernel, same numbers as the dense path."""
    from _helpers import ls_for_half_bandwidth

    N, B = 2048, 16
    targets = [WD - k for WD in (64, 96, 128, 160) for k in (4, 3, 2, 1)]
    d = synth.stage_inputs_direct(N, B, n_comp=6, n_local=2)
    d["glob"][:, 1] = [ls_for_half_bandwidth(d["wave"], b) for b in targets]
    eng = _engine(N, 6, 2, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    classes = eng.band_classes()
    assert classes == {64: 4, 96: 4, 128: 4, 160: 4, 192: 0, 256: 0, 0: 0}, classes
    assert (info == 0).all()
    for b in range(B):
        ref = _dense_ref(d, b)
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref)), (b, targets[b], lnL[b], ref)
    eng.set_solver("dense")
    lnL_d, _ = _run(eng, d)
    assert np.abs(lnL - lnL_d).max() <= TOL * np.abs(lnL_d).max()
    eng.close()


def test_config2_shape_global_only_no_emulator_term():
    N, B = 4096, 5
    d = synth.stage_inputs_direct(N, B, n_comp=0, n_local=0)
    eng = _engine(N, 0, 1, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    assert (info == 0).all() and eng.band_classes()[0] == 0
    for b in range(B):
        ref = S.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], None, None, d["model_flux"][b],
                                     glob=d["glob"][b], loc=())
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref))
    eng.close()


def test_config3_shape_n8192_against_cpu_checker_and_dense_cuda():
    N, B = 8192, 12
    d = synth.stage_inputs_direct(N, B, n_comp=6, n_local=2)
    eng = _engine(N, 6, 2, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    assert (info == 0).all()
    classes = eng.band_classes()
    assert classes[0] == 0 and sum(classes.values()) == B and classes[128] > 0 and classes[160] > 0, classes
    assert classes[192] == 0 and classes[256] == 0
    for b in range(B):
        ref = S.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], d["A"][b], d["model_flux"][b],
                                     glob=d["glob"][b], loc=d["loc"][b])
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    eng.set_solver("dense")
    lnL_d, _ = _run(eng, d)
    assert np.abs(lnL - lnL_d).max() <= TOL * np.abs(lnL_d).max()
    eng.close()


def test_shared_hyper_rows_ragged_kernels_and_wide_emulator():
    N, B, K, M = 1000, 5, 4, 12      # N not a multiple of anything convenient; MAXNR = 17 instantiation
    wave = synth.log_uniform_wave(N, 5090.0, 5125.0)
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K, wave=wave)
    d["loc"][:, :, 1] = np.linspace(5095.0, 5120.0, K)
    d["nloc"] = np.array([0, 4, 2, 1, 3], dtype=np.int32)
    d["glob"][1] = 0.0               # walker without a global kernel: S = diag + local blocks
    eng = _engine(N, M, K, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    assert (info == 0).all()
    for b in range(B):
        ref = _dense_ref(d, b)
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    # one shared hyper-parameter row (frozen kernel groups)
    lnL_s, info_s = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"][2:3], nloc=d["nloc"][2:3],
                                       loc=d["loc"][2:3], shared_hyper=True)
    d2 = dict(d)
    d2["glob"], d2["nloc"], d2["loc"] = (np.repeat(d[k][2:3], B, axis=0) for k in ("glob", "nloc", "loc"))
    for b in range(B):
        ref = _dense_ref(d2, b)
        assert abs(lnL_s.cpu().numpy()[b] - ref) <= TOL * max(1.0, abs(ref))
    eng.close()


def test_fallbacks_unsorted_grid_and_wide_local_block():
    N, B = 512, 3
    wave = synth.log_uniform_wave(N, 5090.0, 5125.0)
    d = synth.stage_inputs_direct(N, B, n_comp=3, n_local=1, wave=wave)
    d["loc"][:, 0, 1] = 5107.0
    d["loc"][1, 0, 2] = 400.0        # 4σ = 1600 km/s on 4 km/s pixels: block of ~390 px > any window
    eng = _engine(N, 3, 1, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    assert eng.band_classes()[0] == 1
    for b in range(B):
        if info[b] == 0:
            ref = _dense_ref(d, b)
            assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref))
    assert info[0] == 0 and info[2] == 0
    # unsorted grid: everything takes the dense path, results equal the dense oracle
    perm = np.random.default_rng(2).permutation(N)
    du = synth.stage_inputs_direct(N, B, n_comp=3, n_local=1, wave=wave)
    du["loc"][:, 0, 1] = 5107.0
    for k in ("wave", "sigma", "data_flux"):
        du[k] = du[k][perm]
    du["X"], du["model_flux"] = du["X"][:, :, perm], du["model_flux"][:, perm]
    eng.set_data(du["wave"], du["sigma"], du["data_flux"])
    before = eng.band_classes()[0]
    lnL, info = _run(eng, du)
    assert eng.band_classes()[0] == before + B and (info == 0).all()
    for b in range(B):
        ref = _dense_ref(du, b)
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref))
    eng.close()


def test_not_positive_definite_is_reported():
    N, B = 300, 2
    wave = synth.log_uniform_wave(N, 5095.0, 5106.0)
    d = synth.stage_inputs_direct(N, B, n_comp=2, n_local=1, wave=wave)
    d["sigma"][:] = 0.0
    d["glob"][:] = 0.0
    d["loc"][:, 0] = [5.0, 5100.5, 2.0]   # the reference's local kernel is not PSD-guaranteed (test_kernels.py:37-38)
    d["A"][1] = -np.eye(2) * 1e9          # and an indefinite rank-M term
    eng = _engine(N, 2, 1, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    eng.set_solver("dense")
    lnL_d, info_d = _run(eng, d)
    assert ((info != 0) == (info_d != 0)).all()
    assert np.isnan(lnL[info != 0]).all()
    eng.close()


def test_model_level_structured_solver(golden_dir):
    import os

    from _helpers import make_model

    g = dict(np.load(os.path.join(golden_dir, "model_n2048_w0.npz"), allow_pickle=False))
    m = make_model(2048, 0, solver="structured")
    lnl = m.log_likelihood()
    assert abs(lnl - g["lnL"]) <= 1e-10 * abs(g["lnL"])
    md = make_model(2048, 0)
    P = np.tile(m.get_param_vector(), (4, 1))
    P[1:, list(m.labels).index("vz")] = [5.0, -20.0, 60.0]
    P[2, list(m.labels).index("global_cov:log_ls")] += 1.2
    a, b = m.log_likelihood_batch(P), md.log_likelihood_batch(P)
    assert np.abs(a - b).max() <= 1e-9 * np.abs(b).max()


def test_config5_shape_n16384_multi_order_takes_the_256_window():
    """configs[4]: 8 concatenated orders (16384 px, 2 km/s pixels → half-bandwidths 266, 225, 233), 16 local
    kernels: two walkers fit the 256-pixel window, the first one takes the dense path in the same call."""
    B = 3
    d = synth.stage_inputs_orders(B)
    eng = _engine(16384, 6, 16, B, workspace_walkers=2)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = _run(eng, d)
    classes = eng.band_classes()
    assert (info == 0).all() and classes[0] == 1 and classes[256] == 2, classes
    for b in range(B):
        ref = S.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], d["A"][b], d["model_flux"][b],
                                     glob=d["glob"][b], loc=d["loc"][b])
        assert abs(lnL[b] - ref) <= TOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    eng.close()


def test_randomised_hyper_parameters_dense_vs_structured():
    """24 walkers with hyper-parameters drawn over two decades of amplitude, ℓ ∈ [5, 90] km/s, 0–3 local kernels of
    random position/width (some overlapping, some outside the grid): both solvers, same numbers."""
    N, B, K, M = 1536, 24, 3, 5
    rng = np.random.default_rng(21)
    wave = synth.log_uniform_wave(N, 5050.0, 5250.0)
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K, wave=wave)
    d["glob"][:, 0] = 10.0 ** rng.uniform(-5.0, -3.0, B)
    d["glob"][:, 1] = rng.uniform(5.0, 90.0, B)
    d["glob"][rng.integers(0, B, 3), 0] = 0.0                  # a few walkers without a global kernel
    d["nloc"] = rng.integers(0, K + 1, B).astype(np.int32)
    d["loc"][:, :, 0] = 10.0 ** rng.uniform(-5.0, -3.5, (B, K))
    d["loc"][:, :, 1] = rng.uniform(5040.0, 5260.0, (B, K))    # some centres fall off the grid
    d["loc"][:, :, 2] = rng.uniform(8.0, 60.0, (B, K))
    eng = _engine(N, M, K, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL_s, info_s = _run(eng, d)
    eng.set_solver("dense")
    lnL_d, info_d = _run(eng, d)
    assert (info_s == 0).all() and (info_d == 0).all()
    assert np.abs(lnL_s - lnL_d).max() <= TOL * np.abs(lnL_d).max()
    for b in (0, 7, 19):
        ref = _dense_ref(d, b)
        assert abs(lnL_s[b] - ref) <= TOL * max(1.0, abs(ref))
    eng.close()
