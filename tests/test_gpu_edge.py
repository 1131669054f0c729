"""GPU edge cases: ragged local-kernel counts, unsorted / odd-length grids, wide emulators, tiny sizes,
non-finite input, empty batch."""
import numpy as np
import pytest

from oracle import starfish_oracle as O
from starfish_b200 import synth

pytestmark = pytest.mark.gpu


def _engine(N, M, K, B, **kw):
    from starfish_b200.engine import LikelihoodEngine

    return LikelihoodEngine(N, M, K, B, **kw)


def _dense_ref(d, b, nloc=None):
    loc = d["loc"][b][: (d["nloc"][b] if nloc is None else nloc)]
    glob = d["glob"][b] if d["glob"][b][0] > 0 else None
    cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, glob, loc)
    if d["X"] is not None:
        cov += d["X"][b].T @ d["A"][b] @ d["X"][b]
    keep = cov.copy()   # log_likelihood factorises in place
    return O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0], keep


@pytest.mark.parametrize("N,M", [(1, 0), (5, 2), (127, 1), (129, 3), (333, 6), (257, 8), (300, 12), (260, 16)])
def test_odd_sizes_and_emulator_widths(N, M):
    """Odd N (scalar-store and plain-load fallbacks), N not a multiple of the tile, every M code path."""
    B = 3
    wave = synth.log_uniform_wave(max(N, 2), 5095.0, 5106.0)[:N]
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=2, wave=wave)
    d["loc"][:, :, 1] = [5098.0, 5103.0]
    eng = _engine(N, M, 2, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    X, A = (d["X"], d["A"]) if M else (None, None)
    lnL, info = eng.log_likelihood(X, A, d["model_flux"], glob=d["glob"], loc=d["loc"])
    C = eng.build_covariance(X, A, glob=d["glob"], loc=d["loc"]).cpu().numpy()
    lnL = lnL.cpu().numpy()
    assert (info.cpu().numpy() == 0).all()
    for b in range(B):
        ref, cov = _dense_ref(d, b)
        assert abs(lnL[b] - ref) <= 1e-10 * max(1.0, abs(ref))
        assert np.abs(C[b] - cov).max() <= 1e-13 * cov.diagonal().max()
    eng.close()


def test_ragged_local_kernel_counts_and_no_global():
    N, B, K = 384, 4, 5
    wave = synth.log_uniform_wave(N, 5095.0, 5106.0)
    d = synth.stage_inputs_direct(N, B, n_comp=4, n_local=K, wave=wave)
    d["loc"][:, :, 1] = np.linspace(5096.5, 5104.5, K)
    d["nloc"] = np.array([0, 5, 2, 1], dtype=np.int32)
    d["glob"][1] = 0.0          # amplitude 0 => walker without a global kernel
    eng = _engine(N, 4, K, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], nloc=d["nloc"], loc=d["loc"])
    lnL = lnL.cpu().numpy()
    assert (info.cpu().numpy() == 0).all()
    for b in range(B):
        ref, _ = _dense_ref(d, b)
        assert abs(lnL[b] - ref) <= 1e-10 * max(1.0, abs(ref)), b
    eng.close()


def test_unsorted_wavelength_grid():
    """Tile rejection relies on a sorted grid; a shuffled grid must take the per-element path and still match."""
    N, B = 300, 2
    rng = np.random.default_rng(2)
    wave = rng.permutation(synth.log_uniform_wave(N, 5095.0, 5106.0))
    d = synth.stage_inputs_direct(N, B, n_comp=3, n_local=1, wave=wave)
    d["loc"][:, :, 1] = 5100.0
    eng = _engine(N, 3, 1, B)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    C = eng.build_covariance(d["X"], d["A"], glob=d["glob"], loc=d["loc"]).cpu().numpy()
    for b in range(B):
        _, cov = _dense_ref(d, b)
        assert np.abs(C[b] - cov).max() <= 1e-13 * cov.diagonal().max()
    eng.close()


def test_nan_input_is_reported_not_propagated_silently():
    N = 256
    d = synth.stage_inputs_direct(N, 2, n_comp=2, n_local=0, wave=synth.log_uniform_wave(N, 5095.0, 5106.0))
    d["X"][1, 0, 17] = np.nan
    eng = _engine(N, 2, 1, 2)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    lnL, info = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"])
    info, lnL = info.cpu().numpy(), lnL.cpu().numpy()
    assert info[0] == 0 and np.isfinite(lnL[0])
    assert info[1] > 0 and np.isnan(lnL[1])
    eng.close()


def test_empty_batch_and_oversized_batch():
    import torch

    from starfish_b200._lib import SfbError

    eng = _engine(128, 0, 1, 2)
    eng.set_data(synth.log_uniform_wave(128, 5095.0, 5100.0), np.ones(128), np.zeros(128))
    lnL, info = eng.log_likelihood(None, None, np.zeros((0, 128)))
    assert lnL.numel() == 0 and info.numel() == 0
    with pytest.raises(SfbError):
        eng.log_likelihood(None, None, np.zeros((3, 128)))     # B > Bmax
    with pytest.raises(ValueError):
        eng.log_likelihood(None, None, np.zeros((1, 64)))      # wrong N
    eng2 = _engine(128, 0, 1, 1)
    with pytest.raises(SfbError):
        eng2.log_likelihood(None, None, np.zeros((1, 128)))    # set_data not called
    eng.close(); eng2.close()


def test_host_buffer_path_matches_device_path():
    import torch

    N, B, M, K = 640, 7, 6, 2
    wave = synth.log_uniform_wave(N, 5090.0, 5130.0)
    d = synth.stage_inputs_direct(N, B, n_comp=M, n_local=K, wave=wave)
    d["loc"][:, :, 1] = [5100.0, 5120.0]
    eng = _engine(N, M, K, B, workspace_walkers=4)
    eng.set_data(d["wave"], d["sigma"], d["data_flux"])
    dev, _ = eng.log_likelihood(d["X"], d["A"], d["model_flux"], glob=d["glob"], loc=d["loc"])
    lnL = np.zeros(B); info = np.zeros(B, dtype=np.int32); resid = np.zeros((B, N))
    eng.log_likelihood_host(d["X"], d["A"], d["model_flux"], d["glob"], d["nloc"], d["loc"], lnL, info,
                            resid_out=resid)
    assert np.array_equal(lnL, dev.cpu().numpy()) and (info == 0).all()
    assert np.array_equal(resid, d["model_flux"] - d["data_flux"])
    eng.close()
