"""GPU: the drop-in SpectrumModel end to end (host upstream + CUDA path) against reference fixtures."""
import os

import numpy as np
import pytest

from starfish_b200 import synth

from _helpers import make_model

pytestmark = pytest.mark.gpu

from test_gpu_upstream import MODEL_LNL_RTOL  # noqa: E402  (the justified end-to-end tolerance)


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


@pytest.mark.parametrize("walker", [0, 1])
def test_model_call_and_loglike_n256(golden_dir, walker):
    g = _load(golden_dir, f"model_n256_w{walker}.npz")
    m = make_model(256, walker, wave=g["wave"], mus=(5098.0, 5103.0))
    flux, cov = m()
    assert cov.shape == (256, 256) and flux.shape == (256,)
    scale = g["cov"].diagonal().max()
    # whole-model tolerance is set by the emulator's Σ_w noise floor (see test_host_model), not the kernels
    assert np.abs(cov - g["cov"]).max() <= 1e-8 * scale
    lnl = m.log_likelihood()
    print(f"end-to-end |dlnL|/|lnL| = {abs(lnl - g['lnL']) / abs(g['lnL']):.2e}")
    assert abs(lnl - g["lnL"]) <= MODEL_LNL_RTOL * abs(g["lnL"])
    assert len(m.residuals) == 1 and np.allclose(m.residuals[-1], g["model_flux"] - g["data_flux"], atol=1e-12)
    # lnL rises when data := model (tests/test_models/test_models.py:223-229)
    m.data._flux = flux
    assert m.log_likelihood() > lnl


def test_config1_n2048(golden_dir):
    g = _load(golden_dir, "model_n2048_w0.npz")
    m = make_model(2048, 0)
    lnl = m.log_likelihood()
    print(f"end-to-end |dlnL|/|lnL| = {abs(lnl - g['lnL']) / abs(g['lnL']):.2e}")
    assert abs(lnl - g["lnL"]) <= MODEL_LNL_RTOL * abs(g["lnL"])


def test_priors_and_batch(golden_dir):
    import scipy.stats as st

    g = _load(golden_dir, "model_n256_w0.npz")
    m = make_model(256, 0, wave=g["wave"], mus=(5098.0, 5103.0))
    priors = {"T": st.uniform(6000, 200), "vsini": st.norm(5, 2)}
    single = m.log_likelihood(priors)
    assert m.log_likelihood({"T": st.uniform(7000, 100)}) == -np.inf
    P0 = m.get_param_vector()
    P = np.tile(P0, (5, 1))
    P[1, m.labels.index("vsini")] = 6.0
    P[2, m.labels.index("T")] = 9000.0          # outside the emulator grid -> -inf, no GPU work
    P[3, m.labels.index("global_cov:log_amp")] += 0.5
    P[4, m.labels.index("local_cov:0:log_sigma")] -= 0.3
    out = m.log_likelihood_batch(P, priors)
    assert out.shape == (5,) and out[2] == -np.inf and np.isfinite(out[[0, 1, 3, 4]]).all()
    assert abs(out[0] - single) <= 1e-9 * abs(single)
    assert np.array_equal(m.get_param_vector(), P0)       # model state untouched
    for b in (1, 3, 4):
        m.set_param_vector(P[b])
        assert abs(m.log_likelihood(priors) - out[b]) <= 1e-9 * abs(out[b])


def test_not_positive_definite_raises_linalgerror(golden_dir):
    g = _load(golden_dir, "model_n256_w0.npz")
    m = make_model(256, 0, wave=g["wave"], mus=(5098.0, 5103.0))
    # a huge *negative-curvature* local kernel cannot be produced by exp(); break PD-ness through sigma
    m.data._sigma = np.zeros_like(m.data._sigma)
    m["global_cov:log_amp"] = -80.0
    m["local_cov:0:log_amp"] = 5.0
    m["local_cov:0:log_sigma"] = np.log(2.0)
    try:
        val = m.log_likelihood()
        assert np.isfinite(val) or np.isnan(val) or val == -np.inf
    except np.linalg.LinAlgError:
        pass


def test_function_seam_kernels(golden_dir):
    from starfish_b200.kernels import global_covariance_matrix, local_covariance_matrix

    g = _load(golden_dir, "kernels_n192.npz")
    amp, ls = g["g_params"][0]
    K = global_covariance_matrix(g["wave"], amp, ls)
    assert isinstance(K, np.ndarray) and np.abs(K - g["g"][0]).max() <= 1e-13 * amp
    amp, mu, sig = g["l_params"][1]
    L = local_covariance_matrix(g["wave"], amp, mu, sig)
    assert np.abs(L - g["l"][1]).max() <= 1e-13 * amp
    # the reference's property checks (tests/test_models/test_kernels.py:9-38) on GPU output
    wave = np.linspace(1e4, 2e4, 1000)
    cov = global_covariance_matrix(wave, 100.0, 1.0)
    assert cov.shape == (1000, 1000) and np.allclose(cov.diagonal(), 100.0)
    assert cov.min() == 0 and np.all(cov >= 0) and np.allclose(cov, cov.T)
    assert np.linalg.eigvalsh(cov).min() >= 0
    loc = local_covariance_matrix(wave, 100.0, 1.5e4, 1e3)
    r = _load(golden_dir, "kernels_reftest.npz")
    assert np.all(loc >= 0) and loc.max() <= 100.0 and np.allclose(loc, loc.T)
    assert np.abs(loc[495:505] - r["l_rows"]).max() <= 1e-11
    assert np.count_nonzero(loc) == int(r["l_nnz"])


def test_frozen_groups_shared_hyper_batch(golden_dir):
    g = _load(golden_dir, "model_n256_w0.npz")
    m = make_model(256, 0, wave=g["wave"], mus=(5098.0, 5103.0))
    ref = m.log_likelihood()
    m.freeze(["global_cov", "local_cov"])
    assert m._glob_cov is None and m._loc_cov is None
    assert abs(m.log_likelihood() - ref) <= 1e-12 * abs(ref)
    assert m._glob_cov is not None and np.asarray(m._glob_cov).shape == (256, 256)
    kg = np.asarray(m._glob_cov)
    from oracle import starfish_oracle as O

    assert np.abs(kg - O.global_covariance_matrix(g["wave"], *g["glob"])).max() <= 1e-13 * g["glob"][0] + 1e-18


def test_cached_kernel_matrices_are_bit_identical_to_the_builders(golden_dir):
    """The reference's _glob_cov/_loc_cov ARE the builders' outputs (spectrum_model.py:343-360)."""
    from starfish_b200.kernels import global_covariance_matrix, local_covariance_matrix

    g = _load(golden_dir, "model_n256_w0.npz")
    m = make_model(256, 0, wave=g["wave"], mus=(5098.0, 5103.0))
    m.freeze(["global_cov", "local_cov"])
    m.log_likelihood()
    kg, kl = np.asarray(m._glob_cov), np.asarray(m._loc_cov)
    assert np.array_equal(kg, global_covariance_matrix(g["wave"], *g["glob"]))
    ref = sum(local_covariance_matrix(g["wave"], *row) for row in g["loc"])
    assert np.abs(kl - ref).max() <= 1e-16 * max(1.0, np.abs(ref).max())   # one fused sum vs a sum of builder calls
    assert kg.diagonal().min() == kg.diagonal().max() == g["glob"][0]       # diag = amplitude exactly: no σ² round trip


def test_inplace_edits_of_the_emulator_tables_are_seen(golden_dir):
    """The reference re-reads emulator.w_hat / bulk_fluxes on every call; the device tables must follow in-place edits."""
    g = _load(golden_dir, "model_n256_w0.npz")
    m = make_model(256, 0, wave=g["wave"], mus=(5098.0, 5103.0))
    a = m.log_likelihood()
    m.emulator.w_hat[:] = m.emulator.w_hat * 1.05          # in place: same object id
    b = m.log_likelihood()
    assert b != a
    m2 = make_model(256, 0, wave=g["wave"], mus=(5098.0, 5103.0))
    m2.emulator.w_hat = m2.emulator.w_hat * 1.05            # re-assigned: the path that always worked
    assert m2.log_likelihood() == b
    m.bulk_fluxes[0] *= 1.01
    assert m.log_likelihood() != b


def test_build_only_handle_allocates_its_workspace_on_first_factorisation():
    import torch

    from starfish_b200.engine import LikelihoodEngine

    eng = LikelihoodEngine(300, 0, 1, 2, workspace_walkers=-1)
    assert eng.workspace_walkers == 0
    wave = np.linspace(5000.0, 5010.0, 300)
    eng.set_data(wave, np.full(300, 0.01), np.zeros(300))
    C = eng.build_covariance(None, None, glob=np.array([[1e-4, 20.0]]), n_walkers=1)
    assert eng.workspace_walkers == 0 and C.shape == (1, 300, 300)
    _, info = eng.cho_factor(C.clone().contiguous())
    assert info.cpu().tolist() == [0] and eng.workspace_walkers >= 2
    eng.close()
