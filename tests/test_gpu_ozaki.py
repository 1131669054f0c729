"""GPU parity tests of the int8 tensor-core trailing update (solver 'dense_i8', csrc/ozaki.cu).

The mode changes only HOW A_ij −= L_ik·L_jkᵀ is evaluated (fixed-point operands, exact int8×int8→int32 products
on tcgen05, fp64 recombination); diagonal-tile factorisation, triangular solves, logdet and the forward solve stay
true fp64.  It must therefore meet the same bar as the fp64 path: |ΔlnL| ≤ 1e-10·max(1,|lnL|) against the dense
CPU oracle (scipy cho_factor — the reference's own routine), including the ill-conditioned stress fixture recorded
from the unmodified reference, and factors that agree with LAPACK's element-wise.  The CPU experiment
tools/ozaki_experiment.py predicts ≤ 1.6e-12 for this slice/anti-diagonal choice.
"""
import os

import numpy as np
import pytest

from oracle import starfish_oracle as O
from oracle import structured_oracle as SO
from starfish_b200 import synth

pytestmark = pytest.mark.gpu

LNL_RTOL = 1e-10


def _engine(N, M, K, B, **kw):
    from starfish_b200.engine import LikelihoodEngine

    eng = LikelihoodEngine(N, M, K, B, **kw)
    eng.set_solver("dense_i8")
    return eng


@pytest.mark.parametrize("N", [384, 1000, 1280])
def test_factor_matches_lapack_elementwise(N):
    """cho_factor seam: L from the int8 mode against scipy's, and against the fp64 (DMMA) mode of the same handle.
    N=384 has one 3-column outer block (strips only), 1000 is ragged (identity padding), 1280 spans two outer
    blocks (trailing update + look-ahead + both sliced-panel buffers)."""
    import torch
    from scipy.linalg import cho_factor

    rng = np.random.default_rng(11)
    B = 3
    eng = _engine(N, 0, 1, B)
    mats = []
    for b in range(B):
        a = rng.standard_normal((N, N)) * (10.0 ** rng.uniform(-3, 3, size=(N, 1)))   # badly scaled rows
        mats.append(a @ a.T + np.diag((a * a).sum(1)) * 0.5)
    mats = np.array(mats)
    Cd = torch.from_numpy(mats.copy()).cuda()
    _, info, logdet = eng.cho_factor(Cd, return_logdet=True)
    Lg = np.tril(Cd.cpu().numpy())
    assert info.cpu().tolist() == [0] * B
    eng.set_solver("dense")
    Cd2 = torch.from_numpy(mats.copy()).cuda()
    eng.cho_factor(Cd2)
    L64 = np.tril(Cd2.cpu().numpy())
    worst = 0.0
    for b in range(B):
        Lr = np.tril(cho_factor(mats[b], lower=True)[0])
        scale = np.sqrt(np.diag(mats[b]))[:, None]          # |L_ik| <= sqrt(C_ii): the row-wise scale of the factor
        e_i8 = (np.abs(Lg[b] - Lr) / scale).max()
        e_64 = (np.abs(L64[b] - Lr) / scale).max()
        worst = max(worst, e_i8)
        assert e_i8 <= 1e-12, (b, e_i8, e_64)
        assert abs(logdet.cpu().numpy()[b] - 2 * np.log(np.diag(Lr)).sum()) <= 1e-10 * N
    print(f"N={N}: max row-scaled |L_i8 - L_lapack| = {worst:.2e}")
    eng.close()


def test_stress_set_ill_conditioned(golden_dir):
    """cond up to 7e5; lnL recorded from the unmodified reference."""
    g = dict(np.load(os.path.join(golden_dir, "stress_n2048.npz")))
    N, M = g["wave"].size, g["X"].shape[0]
    S = g["stress"]
    B = len(S)
    eng = _engine(N, M, 2, B)
    eng.set_data(g["wave"], g["sigma"], g["data_flux"])
    A = np.linalg.inv(g["weights_cov"])
    X = np.broadcast_to(g["X"], (B, M, N)).copy()
    F = np.broadcast_to(g["model_flux"], (B, N)).copy()
    lnL, info = eng.log_likelihood(X, np.broadcast_to(A, (B, M, M)).copy(), F, glob=S[:, :2],
                                   loc=np.broadcast_to(g["loc"], (B, 2, 3)).copy())
    lnL = lnL.cpu().numpy()
    assert (info.cpu().numpy() == 0).all()
    err = np.abs(lnL - S[:, 2]) / np.maximum(1.0, np.abs(S[:, 2]))
    print("stress set, int8 mode: max |dlnL|/|lnL| =", err.max())
    assert err.max() <= LNL_RTOL, err
    eng.close()


@pytest.mark.parametrize("N,M,K", [(640, 6, 2), (2048, 6, 2), (1111, 3, 1)])
def test_batch_vs_dense_oracle(N, M, K):
    B = 4
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K)
    eng = _engine(N, M, K, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    lnL, info = lnL.cpu().numpy(), info.cpu().numpy()
    assert (info == 0).all()
    for b in range(B):
        cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, d["glob"][b], d["loc"][b])
        cov += d["X"][b].T @ d["A"][b] @ d["X"][b]
        ref = O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]
        assert abs(lnL[b] - ref) <= LNL_RTOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    lnL2, _ = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    assert np.array_equal(lnL, lnL2.cpu().numpy())     # integer accumulation: deterministic bit for bit
    eng.close()


def test_not_positive_definite_reports_lapack_info():
    import torch

    N = 512
    eng = _engine(N, 0, 1, 2)
    C = torch.eye(N, dtype=torch.float64, device="cuda").repeat(2, 1, 1).contiguous()
    C[1, 300, 300] = -1.0
    _, info = eng.cho_factor(C)
    assert info.cpu().tolist() == [0, 301]
    eng.close()


def test_config3_n8192_against_independent_checker_and_fp64_mode():
    """Headline size: int8 mode vs the banded+Woodbury CPU checker and vs the fp64 DMMA mode of the same handle."""
    B = 6
    d = synth.stage_inputs_direct(8192, B)
    eng = _engine(8192, 6, 2, B, workspace_walkers=4)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    lnL, info = lnL.cpu().numpy(), info.cpu().numpy()
    assert (info == 0).all()
    for b in range(B):
        ref = SO.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], d["A"][b], d["model_flux"][b],
                                      d["glob"][b], d["loc"][b][: d["nloc"][b]])
        assert abs(lnL[b] - ref) <= LNL_RTOL * max(1.0, abs(ref)), (b, lnL[b], ref)
    eng.set_solver("dense")
    l64, _ = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    rel = np.abs(lnL - l64.cpu().numpy()) / np.abs(lnL)
    print("N=8192: max |lnL_i8 - lnL_fp64| / |lnL| =", rel.max())
    assert rel.max() <= 1e-11
    eng.close()


def test_zero_digit_slabs_are_skipped_exactly():
    """The update kernel skips every int8 product with an all-zero digit slab.  (a) a matrix whose factor has large
    entries everywhere (no slab is zero): every MMA is issued; (b) a strongly diagonally dominant matrix (off-diagonal
    factor entries below 2^-8 of the row scale: leading slab zero): fewer are issued — and both factors are exact."""
    import torch
    from scipy.linalg import cho_factor

    rng = np.random.default_rng(5)
    N, B = 1280, 2
    eng = _engine(N, 0, 1, B)
    a = rng.standard_normal((N, N))
    dense_like = a @ a.T / N + 0.5 * np.eye(N)                     # correlations ~ N^-1/2 ≈ 0.03 > 2^-8
    sparse_like = 1e-4 * (a @ a.T) / N + np.eye(N)                 # off-diagonal ~ 3e-6 of the diagonal
    fracs = []
    for mat in (dense_like, sparse_like):
        Cd = torch.from_numpy(np.stack([mat, mat])).cuda()
        eng.i8_mma_counts()
        _, info = eng.cho_factor(Cd)
        issued, dense = eng.i8_mma_counts()
        assert info.cpu().tolist() == [0, 0] and dense > 0
        fracs.append(issued / dense)
        Lr = np.tril(cho_factor(mat, lower=True)[0])
        Lg = np.tril(Cd.cpu().numpy()[0])
        assert (np.abs(Lg - Lr) / np.sqrt(np.diag(mat))[:, None]).max() <= 1e-12
    print("issued / dense-pattern int8 MMAs:", fracs)
    assert fracs[0] > 0.95 and fracs[1] < 0.75
    eng.close()


@pytest.mark.timeout(120)
def test_banded_matrices_with_all_zero_operand_blocks_many_rounds():
    """Global kernel only (M = 0): S is banded, so most 64 x 32 operand blocks of the factor are EXACT zeros — the
    producers of the update kernel have nothing to fetch for long stretches and the MMA warp nothing to issue.  A
    producer that is not needed for a pipeline stage to complete must still keep in step with it (it once fell two
    barrier phases behind and the kernel deadlocked, seen only at bench size): many walkers, repeated calls, int8 mode,
    results against the independent checker and bit-identical between calls."""
    B = 24
    d = synth.stage_inputs_direct(4096, B, n_comp=0, n_local=0)
    eng = _engine(4096, 0, 1, B)
    eng.set_solver("dense_i8")
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    first = None
    for _ in range(6):
        lnL, info = eng.log_likelihood(None, None, d["model_flux"], glob=d["glob"])
        lnL, info = lnL.cpu().numpy(), info.cpu().numpy()
        assert (info == 0).all()
        if first is None:
            first = lnL.copy()
            for b in (0, B // 2, B - 1):
                ref = SO.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], None, None, d["model_flux"][b],
                                              d["glob"][b], d["loc"][b][: d["nloc"][b]])
                assert abs(lnL[b] - ref) <= LNL_RTOL * max(1.0, abs(ref)), (b, lnL[b], ref)
        assert np.array_equal(lnL, first)
    issued, dense = eng.i8_mma_counts()
    assert issued < 0.05 * dense       # almost every product has an all-zero slab
    eng.close()
