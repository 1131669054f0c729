"""numpy restatement of the DEVICE algorithm of csrc/band.cu (development aid, not product, not oracle).

Band storage Sb[i, d] = S[i, i-d]; a WD×WD window addressed circularly (slot = index mod WD) slides along the
diagonal; per pivot: publish column j, rank-1 update of every slot, the row of index j+WD replaces the slots of
the retiring index j; the M+1 right-hand sides ride along and only the Gram matrix of Z = L⁻¹[R | Xᵀ] is kept;
the epilogue is the M×M capacitance system.  Restated so the index algebra (slot ↔ index mapping, the entering
row's band offsets, the Gram/capacitance formulas) can be checked on a CPU against the structured oracle.
"""
import numpy as np


def band_storage(S, WD):
    """Sb[i, d] = S[i, i-d] for d < WD (zero where i-d < 0); raises if the band does not fit."""
    N = S.shape[0]
    Sb = np.zeros((N, WD))
    for i in range(N):
        for d in range(min(WD, i + 1)):
            Sb[i, d] = S[i, i - d]
        if i - WD >= 0 and np.any(S[i, : i - WD + 1] != 0.0):
            raise ValueError("band wider than the window")
    return Sb


def window_loglike(Sb, rhs, A=None):
    """rhs: [N, NR] with column 0 = R and columns 1.. = Xᵀ.  Returns (lnL, info) like band_chol_kernel."""
    N, WD = Sb.shape
    NR = rhs.shape[1]
    M = NR - 1

    def row_of(i):          # band row of index i (identity padding past the end)
        if i < N:
            return Sb[i]
        r = np.zeros(WD)
        r[0] = 1.0
        return r

    def rhs_of(i):
        return rhs[i] if i < N else np.zeros(NR)

    W = np.zeros((WD, WD))            # slot (r, c) = element (i, k), k <= i, i ≡ r, k ≡ c (mod WD)
    for i in range(WD):
        for k in range(i + 1):
            W[i, k] = row_of(i)[i - k]
    rw = np.array([rhs_of(i) for i in range(WD)])
    gram = np.zeros((NR, NR))
    logdet, info = 0.0, 0
    for j in range(N):
        jr = j % WD
        col = W[:, jr].copy()         # every slot of column-residue jr holds (i, j) for the window's i
        pj = col[jr]
        if not pj > 0 and info == 0:
            info = j + 1
        logdet += np.log(pj) if pj > 0 else np.nan
        inv = 1.0 / pj
        W -= np.outer(col, col * inv)  # the kernel updates every slot; dead ones are overwritten before use
        z = rw[jr].copy()
        rw -= np.outer(col, z * inv)
        gram += np.outer(z, z) * inv
        new = row_of(j + WD)           # index j+WD takes the slots of index j
        for c in range(WD):
            t = (c - jr - 1) % WD
            W[jr, c] = new[WD - 1 - t]
        rw[jr] = rhs_of(j + WD)
    quad = gram[0, 0]
    if M > 0 and info == 0:
        G = gram[1:, 1:]
        u = gram[1:, 0]
        K = np.eye(M) + A @ G
        sign, ldk = np.linalg.slogdet(K)
        Gc = np.linalg.cholesky(G + 1e-300 * np.eye(M))
        try:
            np.linalg.cholesky(np.eye(M) + Gc.T @ A @ Gc)
        except np.linalg.LinAlgError:
            info = N
        if sign <= 0:
            info = N
        logdet += ldk
        quad -= u @ np.linalg.solve(K, A @ u)
    return (-(logdet + quad) / 2 if info == 0 else np.nan), info
