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


def window_loglike_blocked(Sb, rhs, A=None):
    """The rank-4 (DMMA) variant, band_mma_kernel: four pivots per step.  Publish the raw columns of the four
    pivots; factor their 4x4 diagonal block (LDL^T); panel solve per row (W = unscaled, L = W/d); one rank-4
    update of every slot; the four rows j+WD .. j+WD+3 enter together, each corrected for the pivots of the block
    it reaches (b <= WD-1 like the rank-1 kernel).  Returns (lnL, info) like window_loglike."""
    N, WD = Sb.shape
    NR = rhs.shape[1]
    M = NR - 1

    def row_of(i):
        if i < N:
            return Sb[i]
        r = np.zeros(WD)
        r[0] = 1.0
        return r

    def rhs_of(i):
        return rhs[i] if i < N else np.zeros(NR)

    W = np.zeros((WD, WD))
    for i in range(WD):
        for k in range(i + 1):
            W[i, k] = row_of(i)[i - k]
    rw = np.array([rhs_of(i) for i in range(WD)])
    gram = np.zeros((NR, NR))
    logdet, info = 0.0, 0
    for j in range(0, N, 4):
        jr = j % WD
        P = W[:, jr:jr + 4].copy()          # raw panel, P[x, t] = slot (x, jr+t)
        zraw = rw[jr:jr + 4].copy()
        D = P[jr:jr + 4]                    # rows of the four pivots
        Lp = np.zeros((4, 4))
        inv = np.zeros(4)
        Wp = np.zeros((WD, 4))
        for t in range(4):                  # panel solve, row-parallel on the device
            Wp[:, t] = P[:, t]
            for k in range(t):
                Wp[:, t] -= Wp[:, k] * Lp[t, k]
            d = Wp[jr + t, t]
            if not d > 0 and info == 0:
                info = j + t + 1
            logdet += np.log(d) if d > 0 else np.nan
            inv[t] = 1.0 / d
            Lp[:, t] = Wp[jr:jr + 4, t] * inv[t]
        L = Wp * inv
        zW = np.zeros((4, NR))
        for t in range(4):
            zW[t] = zraw[t]
            for k in range(t):
                zW[t] -= Wp[jr + t, k] * (zW[k] * inv[k])
        zL = zW * inv[:, None]
        W -= Wp @ L.T                        # the DMMA update: C[r][c] -= sum_t W[r][t] L[c][t]
        rw -= Wp @ zL
        gram += zL.T @ zW
        for s4 in range(4):                  # rows j+WD+s4 take the slots of the retiring indices j+s4
            new = row_of(j + WD + s4)
            # Late-row correction: row j+WD+s4 may reach the pivots j+s4+1 .. j+3 of THIS block (band offsets
            # WD-3+s4 .. WD-1, non-zero when b > WD-4); it was not in the window when they were eliminated, so
            # their updates are applied as it enters: W' = its panel values, then raw - sum W'[t] L[c][t].
            Wl = np.zeros(4)
            for t in range(s4 + 1, 4):
                Wl[t] = new[WD + s4 - t]
                for k in range(s4 + 1, t):
                    Wl[t] -= Wl[k] * Lp[t, k]
            for c in range(WD):
                v = new[(jr + s4 - c) % WD]
                if not (0 <= c - jr < 4):    # the block's own residues: new columns / dead columns, no update
                    for t in range(1, 4):
                        v -= Wl[t] * L[c, t]
                W[jr + s4, c] = v
            rw[jr + s4] = rhs_of(j + WD + s4) - Wl @ zL
    quad = gram[0, 0]
    if M > 0 and info == 0:
        G = gram[1:, 1:]
        u = gram[1:, 0]
        K = np.eye(M) + A @ G
        sign, ldk = np.linalg.slogdet(K)
        if sign <= 0:
            info = N
        logdet += ldk
        quad -= u @ np.linalg.solve(K, A @ u)
    return (-(logdet + quad) / 2 if info == 0 else np.nan), info


def window_loglike_lookahead(Sb, rhs, A=None):
    """The SCHEDULE of band_mma_la_kernel at matrix level: block n's rank-4 update is applied to the tile column
    (8 residues) that holds block n+1's pivots first, block n's entering rows enter in that tile column, the raw
    panel of block n+1 is taken, its panel is solved, and only then the rest of block n's update and entering rows
    follow.  Same arithmetic as window_loglike_blocked — this restatement pins the ordering argument (an entered row
    must not receive block n's update; the entered values belong in the next panel)."""
    N, WD = Sb.shape
    NR = rhs.shape[1]
    M = NR - 1

    def row_of(i):
        if i < N:
            return Sb[i]
        r = np.zeros(WD)
        r[0] = 1.0
        return r

    def rhs_of(i):
        return rhs[i] if i < N else np.zeros(NR)

    W = np.zeros((WD, WD))
    for i in range(WD):
        for k in range(i + 1):
            W[i, k] = row_of(i)[i - k]
    rw = np.array([rhs_of(i) for i in range(WD)])
    gram = np.zeros((NR, NR))
    state = {"logdet": 0.0, "info": 0}

    def panel(j, jr, P, zraw):
        Lp = np.zeros((4, 4))
        inv = np.zeros(4)
        Wp = np.zeros((WD, 4))
        for t in range(4):
            Wp[:, t] = P[:, t]
            for k in range(t):
                Wp[:, t] -= Wp[:, k] * Lp[t, k]
            d = Wp[jr + t, t]
            if not d > 0 and state["info"] == 0:
                state["info"] = j + t + 1
            state["logdet"] += np.log(d) if d > 0 else np.nan
            inv[t] = 1.0 / d
            Lp[:, t] = Wp[jr:jr + 4, t] * inv[t]
        zW = np.zeros((4, NR))
        for t in range(4):
            zW[t] = zraw[t]
            for k in range(t):
                zW[t] -= Wp[jr + t, k] * (zW[k] * inv[k])
        return Wp, Wp * inv, Lp, zW, zW * inv[:, None]

    def enter(j, jr, cols, L, Lp):
        for s4 in range(4):
            new = row_of(j + WD + s4)
            Wl = np.zeros(4)
            for t in range(s4 + 1, 4):
                Wl[t] = new[WD + s4 - t]
                for k in range(s4 + 1, t):
                    Wl[t] -= Wl[k] * Lp[t, k]
            for c in cols:
                v = new[(jr + s4 - c) % WD]
                if not (0 <= c - jr < 4):
                    for t in range(1, 4):
                        v -= Wl[t] * L[c, t]
                W[jr + s4, c] = v
        return

    cur = panel(0, 0, W[:, 0:4].copy(), rw[0:4].copy())
    for j in range(0, N, 4):
        jr = j % WD
        Wp, L, Lp, zW, zL = cur
        jn, jrn = j + 4, (jr + 4) % WD
        has_next = jn < N
        rw -= Wp @ zL                                   # S1
        gram += zL.T @ zW
        first = list(range(8 * (jrn // 8), 8 * (jrn // 8) + 8)) if has_next else []
        rest = [c for c in range(WD) if c not in first]
        if has_next:                                    # S2: the next panel's tile column first
            W[:, first] -= Wp @ L[first].T
            enter(j, jr, first, L, Lp)
            cur = panel(jn, jrn, W[:, jrn:jrn + 4].copy(), rw[jrn:jrn + 4].copy())   # NB + S4
        W[:, rest] -= Wp @ L[rest].T                    # S5
        enter(j, jr, rest, L, Lp)
        for s4 in range(4):
            new = row_of(j + WD + s4)
            Wl = np.zeros(4)
            for t in range(s4 + 1, 4):
                Wl[t] = new[WD + s4 - t]
                for k in range(s4 + 1, t):
                    Wl[t] -= Wl[k] * Lp[t, k]
            rw[jr + s4] = rhs_of(j + WD + s4) - Wl @ zL
    logdet, info = state["logdet"], state["info"]
    quad = gram[0, 0]
    if M > 0 and info == 0:
        G = gram[1:, 1:]
        u = gram[1:, 0]
        K = np.eye(M) + A @ G
        sign, ldk = np.linalg.slogdet(K)
        if sign <= 0:
            info = N
        logdet += ldk
        quad -= u @ np.linalg.solve(K, A @ u)
    return (-(logdet + quad) / 2 if info == 0 else np.nan), info
