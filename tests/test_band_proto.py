"""CPU: the structured solver's algorithm (csrc/band.cu) restated in numpy against the dense and the
banded+Woodbury oracles — pins the circular-window index algebra without a GPU."""
import os
import sys

import numpy as np

from oracle import starfish_oracle as O
from oracle import structured_oracle as S
from starfish_b200 import synth

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))


def test_sliding_window_restatement_matches_oracles():
    import band_proto as P

    N, B, M = 200, 3, 3
    wave = synth.log_uniform_wave(N, 5096.0, 5104.0)
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=1, wave=wave)
    d["loc"][:, 0, 1] = 5100.3
    d["loc"][:, 0, 2] = [6.0, 9.0, 4.0]
    d["glob"][:, 1] = [8.0, 14.0, 3.0]
    for b, WD in zip(range(B), (96, 160, 64)):   # half-bandwidths 83, 144, 32
        glob, loc = d["glob"][b], d["loc"][b]
        Smat = O.assemble_covariance(d["wave"], d["sigma"], None, None, glob, loc)
        np.fill_diagonal(Smat, Smat.diagonal() + O.JITTER)
        Sb = P.band_storage(Smat, WD)
        rhs = np.column_stack([d["model_flux"][b] - d["data_flux"], d["X"][b].T])
        lnl, info = P.window_loglike(Sb, rhs, d["A"][b])
        ref = S.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], d["A"][b], d["model_flux"][b],
                                     glob=glob, loc=loc)
        cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, glob, loc) + d["X"][b].T @ d["A"][b] @ d["X"][b]
        dense = O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]
        assert info == 0
        assert abs(lnl - ref) <= 1e-11 * abs(ref) and abs(lnl - dense) <= 1e-11 * abs(dense)
        # the rank-4 variant (four pivots per step, rows enter four at a time: b <= WD - 4)
        lnl4, info4 = P.window_loglike_blocked(Sb, rhs, d["A"][b])
        assert info4 == 0 and abs(lnl4 - dense) <= 1e-11 * abs(dense)
        # ... and the look-ahead schedule (next panel's tile column updated, entered and published first)
        lnla, infola = P.window_loglike_lookahead(Sb, rhs, d["A"][b])
        assert infola == 0 and abs(lnla - lnl4) <= 1e-13 * abs(lnl4)
    # M = 0 (config-2 shape) and a band that does not fit
    lnl0, info0 = P.window_loglike(Sb, rhs[:, :1])
    ref0 = S.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], None, None, d["model_flux"][B - 1],
                                  glob=d["glob"][B - 1], loc=d["loc"][B - 1])
    assert info0 == 0 and abs(lnl0 - ref0) <= 1e-11 * abs(ref0)
    try:
        P.band_storage(Smat, 8)
        assert False, "expected the band not to fit"
    except ValueError:
        pass
    # an indefinite rank-M term is reported, not silently accepted
    _, info_bad = P.window_loglike(Sb, rhs, -1e9 * np.eye(M))
    assert info_bad == N


def test_bench_ensemble_covers_the_late_row_path():
    """bench.py reports max |lnL_structured - lnL_dense| / |lnL| over its 256 walkers; that figure pins the rank-4
    kernel's late-row correction only if some walkers have a band within three pixels of their window (b > WD - 4).
    The synthetic ensemble does, in every class it uses (host arithmetic only: the band widths follow from the
    kernel hyper-parameters)."""
    N, B, c = 8192, 256, 2.99792458e5
    wave = synth.log_uniform_wave(N)
    tight = {96: 0, 128: 0, 160: 0}
    counts = {96: 0, 128: 0, 160: 0}
    for b in range(B):
        _, p = synth.walker_params(b, n_local=2, with_global=True)
        r0 = 6 * np.exp(p["global_cov"]["log_ls"])
        bw = 0
        for d in range(40, 200):
            r = c / 2 * np.abs((wave[:-d] - wave[d:]) / (wave[:-d] + wave[d:]))
            if not np.any(r <= r0):
                break
            bw = d
        for lk in p["local_cov"]:
            m = c / lk["mu"] * np.abs(wave - lk["mu"])
            bw = max(bw, int(np.sum(m <= 4 * np.exp(lk["log_sigma"]))) - 1)
        WD = next(w for w in (64, 96, 128, 160, 192, 256) if bw + 1 <= w)
        assert WD in counts, (b, bw)
        counts[WD] += 1
        tight[WD] += bw > WD - 4
    assert counts == {96: 14, 128: 206, 160: 36}      # the routing bench.py reports (window_classes / 23 passes)
    assert all(v >= 1 for v in tight.values()), tight


def test_tight_band_inputs_of_the_gpu_test_and_the_rank4_restatement():
    """The inputs of tests/test_gpu_structured.py::test_band_within_three_pixels_of_every_window, checked without a
    GPU: the length scales give exactly the half-bandwidths WD-4 .. WD-1, and the numpy restatement of the rank-4
    kernel (late-row correction included) reproduces the dense oracle on the tightest walker of two classes."""
    import band_proto as P
    from _helpers import ls_for_half_bandwidth

    N, B = 2048, 16
    targets = [WD - k for WD in (64, 96, 128, 160) for k in (4, 3, 2, 1)]
    d = synth.stage_inputs_direct(N, B, n_comp=6, n_local=2)
    d["glob"][:, 1] = [ls_for_half_bandwidth(d["wave"], b) for b in targets]
    for b in range(B):
        Smat = O.assemble_covariance(d["wave"], d["sigma"], None, None, d["glob"][b], d["loc"][b])
        i, k = np.nonzero(Smat)
        assert int(np.max(i - k)) == targets[b], (b, targets[b])
        if b not in (7, 15):          # b = 95 in the 96-pixel window, b = 159 in the 160-pixel window
            continue
        cov = Smat + d["X"][b].T @ d["A"][b] @ d["X"][b]
        dense = O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]
        np.fill_diagonal(Smat, Smat.diagonal() + O.JITTER)
        Sb = P.band_storage(Smat, targets[b] + 1)
        rhs = np.column_stack([d["model_flux"][b] - d["data_flux"], d["X"][b].T])
        lnl4, info4 = P.window_loglike_blocked(Sb, rhs, d["A"][b])
        assert info4 == 0 and abs(lnl4 - dense) <= 1e-11 * abs(dense)
        # ... and the look-ahead schedule (next panel's tile column updated, entered and published first)
        lnla, infola = P.window_loglike_lookahead(Sb, rhs, d["A"][b])
        assert infola == 0 and abs(lnla - lnl4) <= 1e-13 * abs(lnl4)
