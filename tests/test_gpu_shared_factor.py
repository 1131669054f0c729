"""GPU: the shared-factor path for frozen kernel groups (sfb_loglike with shared_hyper=1, B >= 2).

Reference behaviour: with "global_cov" / "local_cov" frozen, SpectrumModel reuses the cached kernel matrices
(Starfish/models/spectrum_model.py:341-363) — every walker of an ensemble step then has the same
S = diag(σ²+1e-10) + K_global + ΣK_local.  The library factorises S once and solves all walkers' right-hand sides
against it; the result must equal the dense oracle (which builds and factorises every walker's full covariance)
to the stage tolerance 1e-10, for both dense trailing-update modes, with and without the emulator term, and must
report non-positive-definite cases."""
import numpy as np
import pytest

from oracle import starfish_oracle as O
from oracle import structured_oracle as SO
from starfish_b200 import synth

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-10


def _engine(N, M, K, B, solver="dense", **kw):
    from starfish_b200.engine import LikelihoodEngine

    eng = LikelihoodEngine(N, M, K, B, **kw)
    eng.set_solver(solver)
    return eng


def _oracle(d, b, glob, loc):
    cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, glob, loc)
    if d["X"] is not None:
        cov += d["X"][b].T @ d["A"][b] @ d["X"][b]
    return O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]


@pytest.mark.parametrize("solver", ["dense", "dense_i8"])
@pytest.mark.parametrize("N,M,K,B", [(640, 6, 2, 5), (1000, 3, 1, 9), (1536, 0, 2, 4), (2048, 6, 2, 19)])
def test_frozen_kernels_match_dense_oracle(solver, N, M, K, B):
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K)
    glob, loc = d["glob"][2 % B], d["loc"][2 % B]          # one hyper-parameter row for everybody
    eng = _engine(N, M, K, B, solver)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    calls0 = eng.shared_factor_calls
    lnL, info, resid = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=glob[None], loc=loc[None],
                                          shared_hyper=True, return_residuals=True)
    assert eng.shared_factor_calls == calls0 + 1           # the shared-factor path ran, not B factorisations
    lnL, info, resid = lnL.cpu().numpy(), info.cpu().numpy(), resid.cpu().numpy()
    assert (info == 0).all()
    worst = 0.0
    for b in range(B):
        ref = _oracle(d, b, glob, loc)
        worst = max(worst, abs(lnL[b] - ref) / max(1.0, abs(ref)))
        assert np.array_equal(resid[b], d["model_flux"][b] - d["data_flux"])
    print(f"{solver} N={N} M={M} B={B}: max |dlnL|/|lnL| vs dense oracle = {worst:.2e}")
    assert worst <= LNL_RTOL
    # the same call with the path switched off factorises every walker's covariance: same answer to rounding
    eng.set_shared_factor(False)
    l2, i2 = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=glob[None], loc=loc[None], shared_hyper=True)
    assert eng.shared_factor_calls == calls0 + 1
    rel = np.abs(l2.cpu().numpy() - lnL) / np.maximum(1.0, np.abs(lnL))
    assert rel.max() <= LNL_RTOL
    eng.close()


def test_host_buffer_entry_takes_the_shared_path():
    N, M, K, B = 768, 6, 2, 6
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K)
    eng = _engine(N, M, K, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    loc = np.zeros((1, eng.K, 3))
    loc[0, :K] = d["loc"][0]
    lnl_h, info_h = np.zeros(B), np.zeros(B, dtype=np.int32)
    eng.log_likelihood_host(d["X"], d["A"], d["model_flux"], np.ascontiguousarray(d["glob"][:1]),
                            np.ascontiguousarray(d["nloc"][:1].astype(np.int32)), loc, lnl_h, info_h, shared_hyper=True)
    assert eng.shared_factor_calls == 1 and (info_h == 0).all()
    for b in range(B):
        ref = _oracle(d, b, d["glob"][0], d["loc"][0])
        assert abs(lnl_h[b] - ref) <= LNL_RTOL * max(1.0, abs(ref))
    eng.close()


def test_not_positive_definite_is_reported():
    """(a) S itself not PD (a large negative-amplitude local kernel) -> every walker gets S's LAPACK info;
    (b) S fine but S + XᵀAX not PD for ONE walker (negative-definite A) -> only that walker is flagged."""
    N, M, K, B = 512, 4, 1, 4
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K)
    eng = _engine(N, M, K, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    bad_loc = d["loc"][:1].copy()
    bad_loc[0, 0, 0] = -1.0
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"][:1], loc=bad_loc, shared_hyper=True)
    info = info.cpu().numpy()
    assert (info > 0).all() and len(set(info.tolist())) == 1 and np.isnan(lnL.cpu().numpy()).all()
    A = d["A"].copy()
    A[1] = -1e6 * np.eye(M)
    lnL, info = eng.log_likelihood(d["X"], A, d["model_flux"], glob=d["glob"][:1], loc=d["loc"][:1], shared_hyper=True)
    info, lnL = info.cpu().numpy(), lnL.cpu().numpy()
    assert info[1] != 0 and np.isnan(lnL[1]) and (np.delete(info, 1) == 0).all() and np.isfinite(np.delete(lnL, 1)).all()
    eng.close()


def test_frozen_ensemble_n8192_against_independent_checker():
    """Headline size: 24 walkers sharing the kernels, int8 factorisation of S, checked with the banded+Woodbury CPU
    evaluation (oracle/structured_oracle.py)."""
    B = 24
    d = synth.stage_inputs_direct(8192, B)
    eng = _engine(8192, 6, 2, B, "dense_i8", workspace_walkers=2)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"][:1], loc=d["loc"][:1], shared_hyper=True)
    lnL, info = lnL.cpu().numpy(), info.cpu().numpy()
    assert (info == 0).all() and eng.shared_factor_calls == 1
    for b in range(0, B, 5):
        ref = SO.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], d["A"][b], d["model_flux"][b],
                                      d["glob"][0], d["loc"][0][: d["nloc"][0]])
        assert abs(lnL[b] - ref) <= LNL_RTOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    eng.close()
