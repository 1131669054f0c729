"""How many int8 slices does an exact-product (Ozaki-style) trailing update need?  CPU experiment.

The dense path's ceiling is the fp64 DMMA rate (201 evals/s per GPU at N=8192).  The only way past it is to run
the trailing update ``A_ij -= L_ik L_jk^T`` (N^3/3 of the FLOPs) on the int8 tensor path (tcgen05 kind::i8,
exact int32 accumulation) with fp64 recombination.  This is NOT iterative refinement: the panel work (diagonal
tile factorisation, triangular solve, forward substitution of the residual, logdet) stays true fp64; only the
operands of the update are replaced by a fixed-point restatement of the already computed fp64 L:

    row scale   2^e_i  >= 2*sqrt(C_ii)            (|L_ik| <= sqrt(C_ii), so |L_ik / 2^e_i| <= 1/2)
    q_ik      = rint(L_ik / 2^e_i * 2^(8S-1))      (an (8S)-bit signed integer)
    q_ik      = sum_t b_ikt 256^(S-1-t),  b in [-128, 127]   (balanced radix-256 digits, t = 0 most significant)
    L_ik L_jk ~ 2^(e_i+e_j) 2^(-16S+2) sum_{s+t<=D} 256^(2S-2-s-t) (sum_k b_iks b_jkt)

Every digit product b*b is exact in int8 x int8 -> int32 and so is its sum over k <= 1024 for all pairs of one
anti-diagonal s+t = d (|sum| < 7*2^24), hence ONE int32 accumulator per anti-diagonal.  The experiment below
emulates exactly this arithmetic in numpy (digit matmuls in float64 are exact: |sum| < 2^53) inside the blocked
right-looking Cholesky the GPU runs, and reports |dlnL|/|lnL| against the dense oracle (scipy cho_factor) for

  * the ill-conditioned stress fixture (tests/golden/stress_n2048.npz, cond up to 7e5), and
  * synthetic bench walkers (cond 10-20) at N=2048.

Usage:  python tools/ozaki_experiment.py [--n 2048] [--walkers 4]
Output: a table  S (slices) x D (highest anti-diagonal kept) -> max relative lnL error, number of int8 GEMMs.
"""
from __future__ import annotations

import argparse
import os
import sys

import numpy as np
from scipy.linalg import solve_triangular

sys.path.insert(0, os.path.join(os.path.dirname(__file__), ".."))
from oracle import starfish_oracle as so  # noqa: E402
from starfish_b200 import synth  # noqa: E402

NB = 128


def row_exponents(diag):
    """e_i with 2^e_i in (2 sqrt(C_ii), 4 sqrt(C_ii)] -> |L_ik| / 2^e_i < 1/2."""
    return np.floor(np.log2(np.sqrt(diag))).astype(np.int64) + 2


def digits(L, e, S):
    """Balanced radix-256 digits of rint(L / 2^e * 2^(8S-1)); returns list of float64 arrays, t=0 most significant."""
    q = np.rint(np.ldexp(L, (8 * S - 1) - e[:, None])).astype(np.int64)
    out = []
    for _ in range(S):
        d = ((q + 128) & 0xFF) - 128      # in [-128, 127]
        out.append(d.astype(np.float64))
        q = (q - d) >> 8
    assert np.all(q == 0), "top digit overflow"
    return out[::-1]


def ozaki_product(La, ea, Lb, eb, S, D):
    """sum_k La[i,k] Lb[j,k] through the sliced restatement; anti-diagonals 0..D kept."""
    da, db = digits(La, ea, S), digits(Lb, eb, S)
    acc = np.zeros((La.shape[0], Lb.shape[0]))
    for d in range(D, -1, -1):  # Horner from the least significant anti-diagonal
        Sd = np.zeros_like(acc)
        for s in range(max(0, d - S + 1), min(d, S - 1) + 1):
            Sd += da[s] @ db[d - s].T   # exact (integers < 2^53)
        acc = acc / 256.0 + Sd if d != D else Sd
    # acc = sum_d Sd 256^(-d); the weight of anti-diagonal d is 256^(2S-2-d) * 2^(-16S+2) = 256^(-d) * 2^-14
    return np.ldexp(acc, ea[:, None] + eb[None, :] - 14)


def chol_lnl(C, R, S=None, D=None):
    """Blocked right-looking Cholesky (panel 128) + folded forward solve.  S=None: plain fp64 updates."""
    A = C.copy()
    n = A.shape[0]
    e = row_exponents(np.diag(C).copy())
    rhs = R.copy()
    logdet = 0.0
    sq = 0.0
    for k0 in range(0, n, NB):
        k1 = min(n, k0 + NB)
        Lkk = np.linalg.cholesky(A[k0:k1, k0:k1])
        logdet += 2.0 * np.sum(np.log(np.diag(Lkk)))
        z = solve_triangular(Lkk, rhs[k0:k1], lower=True)
        sq += z @ z
        if k1 == n:
            break
        P = solve_triangular(Lkk, A[k1:, k0:k1].T, lower=True).T   # L_ik
        rhs[k1:] -= P @ z
        if S is None:
            A[k1:, k1:] -= P @ P.T
        else:
            A[k1:, k1:] -= ozaki_product(P, e[k1:], P, e[k1:], S, D)
    return -(logdet + sq) / 2


def cases(n, walkers):
    out = []
    f = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "stress_n2048.npz"))
    if n == 2048:
        for amp, ls, _ in f["stress"]:
            cov = so.assemble_covariance(f["wave"], f["sigma"], f["X"], f["weights_cov"], (amp, ls),
                                         [tuple(r) for r in f["loc"]])
            out.append((f"stress amp={amp:g} ls={ls:g}", cov, f["model_flux"] - f["data_flux"]))
    d = synth.stage_inputs_direct(n, walkers)
    for b in range(walkers):
        wc = np.linalg.inv(d["A"][b])
        cov = so.assemble_covariance(d["wave"], d["sigma"], d["X"][b], wc, tuple(d["glob"][b]),
                                     [tuple(r) for r in d["loc"][b]])
        out.append((f"synthetic walker {b}", cov, d["model_flux"][b] - d["data_flux"]))
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--n", type=int, default=2048)
    ap.add_argument("--walkers", type=int, default=3)
    a = ap.parse_args()
    cs = cases(a.n, a.walkers)
    configs = [(None, None)] + [(S, D) for S in (5, 6, 7, 8) for D in (S - 1, S)]
    worst = {c: (0.0, "") for c in configs}
    print(f"N={a.n}; {len(cs)} matrices; tolerance 1e-10 relative")
    for name, cov, R in cs:
        ref = so.log_likelihood(cov.copy(), R, np.zeros_like(R))[0]
        cond = np.linalg.cond(cov + 1e-10 * np.eye(cov.shape[0]))
        C = cov + 1e-10 * np.eye(cov.shape[0])
        row = []
        for c in configs:
            v = chol_lnl(C, R, *c)
            rel = abs(v - ref) / max(1.0, abs(ref))
            row.append(rel)
            if rel >= worst[c][0]:
                worst[c] = (rel, name)
        print(f"{name:28s} cond={cond:9.3g} lnL={ref:14.6f}  " + " ".join(f"{r:8.1e}" for r in row))
    print("\nconfig (S slices, D top anti-diagonal) -> int8 GEMMs per fp64 GEMM, worst |dlnL|/|lnL|")
    for c in configs:
        if c[0] is None:
            print(f"  fp64 blocked (no slicing)            : {worst[c][0]:.2e}  ({worst[c][1]})")
        else:
            S, D = c
            n_gemm = sum(min(d, S - 1) - max(0, d - S + 1) + 1 for d in range(D + 1))
            print(f"  S={S} D={D}: {n_gemm:2d} GEMMs, {D + 1} accumulators : {worst[c][0]:.2e}  ({worst[c][1]})")


if __name__ == "__main__":
    main()
