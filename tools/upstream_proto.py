"""numpy prototype of the DEVICE algorithm of csrc/upstream.cu (development aid, not product, not oracle).

It restates, step by step, what the kernels do — static spectrum of the bulk fluxes, half-length complex
inverse FFT with S-way decimation, windowed banded-inverse FIR for the quintic-spline coefficients,
de Boor evaluation with Doppler-scaled knots — so the index algebra can be checked on a CPU against the
reference-following host transforms before spending GPU time.
"""
import numpy as np

K = 5
W = 48


def knot(fw, l):
    nf = len(fw)
    if l < 6:
        return fw[0]
    if l >= nf:
        return fw[nf - 1]
    return fw[l - 3]


def interval(fw, x, scale=1.0):
    """l with t[l] <= x < t[l+1], clamped to [5, nf-1] (FITPACK splev search)."""
    nf = len(fw)
    lo, hi = 5, nf - 1          # invariant: t[lo] <= x or lo == 5 ; answer in [lo, hi]
    while lo < hi:
        mid = (lo + hi + 1) // 2
        if knot(fw, mid) * scale <= x:
            lo = mid
        else:
            hi = mid - 1
    return lo


def bspl(fw, x, l, scale=1.0):
    h = np.zeros(6)
    hh = np.zeros(6)
    h[0] = 1.0
    for j in range(1, K + 1):
        hh[:j] = h[:j]
        h[0] = 0.0
        for i in range(j):
            li = l + i + 1
            lj = li - j
            tli, tlj = knot(fw, li) * scale, knot(fw, lj) * scale
            f = hh[i] / (tli - tlj)
            h[i] = h[i] + f * (tli - x)
            h[i + 1] = f * (x - tlj)
    return h


def collocation_band(fw):
    """band[i, d] = B[i, i-5+d], d=0..10."""
    nf = len(fw)
    band = np.zeros((nf, 11))
    for i in range(nf):
        l = interval(fw, fw[i])
        h = bspl(fw, fw[i], l)
        for m in range(6):
            c = l - 5 + m
            band[i, c - i + 5] = h[m]
    return band


def inverse_band(band, W=W, WW=2 * W):
    """Ginv[j, d] = (B^-1)[j, j-W+d], via windowed solves of B^T g = e_j (no pivoting)."""
    nf = band.shape[0]
    G = np.zeros((nf, 2 * W + 1))
    for j in range(nf):
        a, b = max(0, j - WW), min(nf, j + WW + 1)
        n = b - a
        # T = (B^T)[a:b, a:b]: T[r, c] = B[a+c, a+r]
        T = np.zeros((n, n))
        for c in range(n):
            i = a + c
            for d in range(11):
                r = i - 5 + d - a
                if 0 <= r < n:
                    T[r, c] = band[i, d]
        e = np.zeros(n)
        e[j - a] = 1.0
        g = np.linalg.solve(T, e)
        for d in range(2 * W + 1):
            i = j - W + d
            if a <= i < b:
                G[j, d] = g[i - a]
    return G


def irfft_device(Xs, n, n_sub_max):
    """Xs[0..n/2] -> real x[n], the way broaden_kernel does it."""
    n2 = n // 2
    S = 1
    while n2 // S > n_sub_max:
        S *= 2
    npr = n2 // S
    T = np.exp(2j * np.pi * np.arange(n2) / n)
    k = np.arange(n2)
    Xk, Xr = Xs[k].copy(), np.conj(Xs[n2 - k])
    Xk[0] = Xk[0].real
    Xr[0] = Xs[n2].real
    E = (Xk + Xr) * 0.5
    O = (Xk - Xr) * T[k] * 0.5
    Z = E + 1j * O
    z = np.zeros(n2, dtype=complex)
    bits = int(np.log2(npr))
    for h in range(S):
        kp = np.arange(npr)
        v = np.zeros(npr, dtype=complex)
        for q in range(S):
            v += Z[kp + q * npr] * np.exp(2j * np.pi * h * q / S)
        # e^{2 pi i h k'/n2} = T[2 h k' mod ...]: index 2*h*k' can exceed n2 -> use symmetry
        idx = 2 * h * kp
        v *= np.exp(2j * np.pi * idx / n)
        sm = np.zeros(npr, dtype=complex)
        rev = np.array([int(format(i, f"0{bits}b")[::-1], 2) for i in kp]) if bits else kp
        sm[rev] = v
        ln = 2
        while ln <= npr:
            half = ln // 2
            t = np.arange(npr // 2)
            grp, pos = t // half, t % half
            i = grp * ln + pos
            j = i + half
            w = T[pos * (n // ln)]
            a, bb = sm[i], sm[j] * w
            sm[i], sm[j] = a + bb, a - bb
            ln *= 2
        z[S * kp + h] = sm
    x = np.empty(n)
    x[0::2] = z.real / n2
    x[1::2] = z.imag / n2
    return x
