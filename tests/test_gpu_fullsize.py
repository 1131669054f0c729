"""GPU tests at BASELINE.json's full sizes, checked through size-independent routes: an independent
structure-exploiting CPU evaluation (banded Cholesky + Woodbury, oracle/structured_oracle.py), closed forms,
factor reconstruction L·Lᵀ = C on sampled rows, composition of the API-seam calls, and run-to-run determinism."""
import numpy as np
import pytest

from oracle import structured_oracle as SO
from starfish_b200 import synth

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-10


def _engine(N, M, K, B, **kw):
    from starfish_b200.engine import LikelihoodEngine

    return LikelihoodEngine(N, M, K, B, **kw)


def _check_against_structured(d, lnL, rows):
    for b in rows:
        X = d["X"][b] if d["X"] is not None else None
        A = d["A"][b] if d["A"] is not None else None
        ref = SO.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], X, A, d["model_flux"][b],
                                      d["glob"][b], d["loc"][b][: d["nloc"][b]])
        assert abs(lnL[b] - ref) <= LNL_RTOL * max(1.0, abs(ref)), (b, lnL[b], ref)


def test_config3_n8192_global_local_emulator():
    """configs[2]: N=8192, M=6, K=2 — a slice of the benchmark ensemble, two chunks on two lanes."""
    B = 10
    d = synth.stage_inputs_direct(8192, B)
    eng = _engine(8192, 6, 2, B, workspace_walkers=6)   # 3 slots per lane -> 4 chunks, both lanes, look-ahead
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    lnL1, info = lnL.cpu().numpy(), info.cpu().numpy()
    assert (info == 0).all()
    _check_against_structured(d, lnL1, range(B))
    lnL2, _ = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    assert np.array_equal(lnL1, lnL2.cpu().numpy())          # deterministic, bit for bit
    eng.close()


@pytest.mark.timeout(300)
def test_n8192_two_walkers_against_the_dense_oracle_itself():
    """Headline size, DIRECT comparison: two bench walkers at N=8192 (M=6, K=2) against oracle/starfish_oracle.py — the
    numpy kernels and scipy cho_factor/cho_solve of the reference path, no banded shortcut in between — for both
    trailing-update modes of the dense solver."""
    from oracle import starfish_oracle as O

    B = 2
    d = synth.stage_inputs_direct(8192, B)
    eng = _engine(8192, 6, 2, B, workspace_walkers=2)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    ref = []
    for b in range(B):
        wcov = np.linalg.inv(d["A"][b])       # the oracle takes Σ_w; the stage carries A = Σ_w⁻¹ (as the reference codes it)
        ref.append(O.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], wcov, d["model_flux"][b],
                                          d["glob"][b], d["loc"][b][: d["nloc"][b]]))
    ref = np.array(ref)
    for solver in ("dense_i8", "dense"):
        eng.set_solver(solver)
        lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
        assert (info.cpu().numpy() == 0).all()
        rel = np.abs(lnL.cpu().numpy() - ref) / np.maximum(1.0, np.abs(ref))
        print(f"N=8192 {solver}: max |lnL - dense oracle| / |lnL| = {rel.max():.2e}")
        assert rel.max() <= LNL_RTOL, (solver, rel)
    eng.close()


def test_config2_n4096_global_only():
    """configs[1]: N=4096, global kernel + σ² only (X = NULL)."""
    B = 8
    d = synth.stage_inputs_direct(4096, B, n_comp=0, n_local=0)
    eng = _engine(4096, 0, 1, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(None, None, d["model_flux"], glob=d["glob"])
    assert (info.cpu().numpy() == 0).all()
    _check_against_structured(d, lnL.cpu().numpy(), range(B))
    eng.close()


def test_config5_n16384_multi_order_16_local_kernels():
    """configs[4]: 8 concatenated orders (16384 px), 2 local kernels per order."""
    B = 3
    d = synth.stage_inputs_orders(B)
    eng = _engine(16384, 6, 16, B, workspace_walkers=2)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], nloc=d["nloc"], loc=d["loc"])
    assert (info.cpu().numpy() == 0).all()
    _check_against_structured(d, lnL.cpu().numpy(), range(B))
    eng.close()


def test_closed_form_diagonal_n8192():
    """No kernels, no emulator term: C = diag(σ²+1e-10) and lnL has a closed form."""
    N, B = 8192, 3
    rng = np.random.default_rng(5)
    wave = synth.log_uniform_wave(N)
    sigma = 0.01 + 0.02 * rng.random(N)
    data = rng.standard_normal(N)
    flux = data + 0.03 * rng.standard_normal((B, N))
    eng = _engine(N, 0, 1, B)
    eng.set_data(wave, sigma, data)
    lnL, info = eng.log_likelihood(None, None, flux)
    var = sigma**2 + 1e-10
    ref = -0.5 * (np.log(var).sum() + (((flux - data) ** 2) / var).sum(axis=1))
    assert (info.cpu().numpy() == 0).all()
    assert np.abs(lnL.cpu().numpy() - ref).max() <= 1e-12 * np.abs(ref).max()
    eng.close()


def test_seam_composition_and_factor_reconstruction_n4096():
    """build_cov → cho_factor → solve_lower (the function-seam calls) reproduce the fused path, and the
    factor reconstructs the covariance on sampled rows."""
    N, B = 4096, 2
    d = synth.stage_inputs_direct(N, B)
    eng = _engine(N, 6, 2, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    fused, _ = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    C = eng.build_covariance(d["X"], d["A"], glob=d["glob"], loc=d["loc"], jitter=1e-10)
    C0 = C.cpu().numpy().copy()
    # both triangles written; XᵀAX is symmetric only up to rounding (Σ_m X[m,i]·(AX)[m,j]), like the reference's BLAS
    assert np.abs(C0[0] - C0[0].T).max() <= 1e-15 * C0[0].diagonal().max()
    _, info, logdet = eng.cho_factor(C, return_logdet=True)
    assert (info.cpu().numpy() == 0).all()
    R = d["model_flux"] - d["data_flux"]
    z = eng.solve_lower(C, R).cpu().numpy()
    lnl = -(logdet.cpu().numpy() + (z * z).sum(axis=1)) / 2
    assert np.abs(lnl - fused.cpu().numpy()).max() <= LNL_RTOL * np.abs(lnl).max()
    L = np.tril(C.cpu().numpy()[0])
    rows = np.array([0, 1, 127, 128, 129, 1000, 2047, 2048, 4095])
    rec = L[rows] @ L.T
    assert np.abs(rec - C0[0][rows]).max() <= 1e-13 * C0[0].diagonal().max()
    eng.close()
