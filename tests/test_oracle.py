"""The CPU oracle against (a) the committed fixtures generated from the unmodified reference and
(b) the live reference whenever /root/reference is present (build container)."""
import os

import numpy as np
import pytest

from oracle import ref_loader
from oracle import starfish_oracle as O


def _load(golden_dir, name):
    return dict(np.load(os.path.join(golden_dir, name), allow_pickle=False))


def test_kernel_fixture_bit_exact(golden_dir):
    g = _load(golden_dir, "kernels_n192.npz")
    for k, (amp, ls) in enumerate(g["g_params"]):
        assert np.array_equal(O.global_covariance_matrix(g["wave"], amp, ls), g["g"][k])
    for k, (amp, mu, sig) in enumerate(g["l_params"]):
        assert np.array_equal(O.local_covariance_matrix(g["wave"], amp, mu, sig), g["l"][k])
    assert not g["l"][3].any()  # centre outside the grid: identically zero


def test_reference_property_checks(golden_dir):
    """The reference's own assertions (tests/test_models/test_kernels.py:9-38) on the oracle output."""
    g = _load(golden_dir, "kernels_reftest.npz")
    wave = np.linspace(1e4, 2e4, 1000)
    cov = O.global_covariance_matrix(wave, 100.0, 1.0)
    assert cov.shape == (1000, 1000)
    assert np.allclose(cov.diagonal(), 100.0) and np.array_equal(cov.diagonal(), g["g_diag"])
    assert cov.min() == 0 and np.all(cov >= 0) and np.isclose(cov.max(), 100.0)
    assert np.allclose(cov, cov.T)
    assert np.linalg.eigvalsh(cov).min() >= 0
    loc = O.local_covariance_matrix(wave, 100.0, 1.5e4, 1e3)
    assert np.all(loc >= 0) and loc.max() <= 100.0 and np.allclose(loc, loc.T)
    assert np.array_equal(loc[495:505], g["l_rows"])
    assert np.count_nonzero(loc) == int(g["l_nnz"])
    assert np.isclose(loc.sum(), float(g["l_sum"]), rtol=1e-15)


@pytest.mark.parametrize("walker", [0, 1])
def test_full_model_fixture_n256(golden_dir, walker):
    g = _load(golden_dir, f"model_n256_w{walker}.npz")
    cov = O.assemble_covariance(g["wave"], g["sigma"], g["X"], g["weights_cov"], g["glob"], g["loc"])
    assert np.abs(cov - g["cov"]).max() <= 1e-15 * g["cov"].diagonal().max()
    lnl = O.log_likelihood(cov, g["model_flux"], g["data_flux"])[0]
    assert abs(lnl - g["lnL"]) <= 1e-12 * abs(g["lnL"])


@pytest.mark.parametrize("name", ["model_n2048_w0.npz", "model_n2048_w3.npz"])
def test_config1_fixture_n2048(golden_dir, name):
    g = _load(golden_dir, name)
    cov = O.assemble_covariance(g["wave"], g["sigma"], g["X"], g["weights_cov"], g["glob"], g["loc"])
    assert np.abs(cov[g["cov_rows_idx"]] - g["cov_rows"]).max() <= 1e-15 * g["cov_diag"].max()
    lnl = O.log_likelihood(cov, g["model_flux"], g["data_flux"])[0]
    assert abs(lnl - g["lnL"]) <= 1e-12 * abs(g["lnL"])


def test_stress_fixture(golden_dir):
    g = _load(golden_dir, "stress_n2048.npz")
    for amp, ls, ref in g["stress"][[0, 4, 8]]:
        lnl = O.stage_log_likelihood(g["wave"], g["sigma"], g["data_flux"], g["X"], g["weights_cov"],
                                     g["model_flux"], (amp, ls), g["loc"])
        assert abs(lnl - ref) <= 1e-11 * max(1.0, abs(ref))


def test_not_positive_definite_raises():
    wave = np.linspace(5000, 5010, 64)
    cov = np.eye(64)
    cov[40, 40] = -1.0
    with pytest.raises(np.linalg.LinAlgError):
        O.log_likelihood(cov, np.zeros(64), np.zeros(64))


def test_empty_and_single_pixel():
    assert O.global_covariance_matrix(np.array([]), 1.0, 1.0).shape == (0, 0)
    one = O.assemble_covariance(np.array([5000.0]), np.array([0.1]), None, None, (2.0, 10.0), [(1.0, 5000.0, 5.0)])
    assert one.shape == (1, 1) and np.isclose(one[0, 0], 0.01 + 2.0 + 1.0)


@pytest.mark.skipif(not ref_loader.have_reference(), reason="reference tree not mounted")
def test_live_reference_agrees():
    """Run the unmodified reference here and compare the oracle on the same inputs."""
    import warnings

    from oracle.make_golden import run_reference_model
    from starfish_b200 import synth

    ref_loader.load_reference()
    from Starfish.models.kernels import global_covariance_matrix, local_covariance_matrix

    wave = synth.log_uniform_wave(300, 5095.0, 5106.0)
    assert np.array_equal(global_covariance_matrix(wave, 0.3, 7.0), O.global_covariance_matrix(wave, 0.3, 7.0))
    assert np.array_equal(local_covariance_matrix(wave, 0.3, 5100.0, 9.0),
                          O.local_covariance_matrix(wave, 0.3, 5100.0, 9.0))
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        r = run_reference_model(512, 7)
    cov = O.assemble_covariance(r["wave"], r["sigma"], r["X"], r["weights_cov"], r["glob"], r["loc"])
    assert np.abs(cov - r["cov"]).max() <= 1e-15 * r["cov"].diagonal().max()
    lnl = O.log_likelihood(cov, r["model_flux"], r["data_flux"])[0]
    assert abs(lnl - r["lnL"]) <= 1e-12 * abs(r["lnL"])


def test_structured_oracle_matches_dense():
    """The banded-Cholesky + Woodbury checker used for the full-size GPU tests agrees with the dense oracle."""
    from oracle import structured_oracle as SO
    from starfish_b200 import synth

    d = synth.stage_inputs_direct(768, 3)
    for b in range(3):
        cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, d["glob"][b], d["loc"][b])
        cov += d["X"][b].T @ d["A"][b] @ d["X"][b]
        ref = O.log_likelihood(cov, d["model_flux"][b], d["data_flux"])[0]
        got = SO.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], d["X"][b], d["A"][b],
                                      d["model_flux"][b], d["glob"][b], d["loc"][b])
        assert abs(got - ref) <= 1e-11 * abs(ref)
    # multi-order grid with many local kernels, no emulator term
    d = synth.stage_inputs_orders(1, n_orders=2, n_per_order=256, n_comp=0)
    cov = O.assemble_covariance(d["wave"], d["sigma"], None, None, d["glob"][0], d["loc"][0])
    ref = O.log_likelihood(cov, d["model_flux"][0], d["data_flux"])[0]
    got = SO.stage_log_likelihood(d["wave"], d["sigma"], d["data_flux"], None, None, d["model_flux"][0],
                                  d["glob"][0], d["loc"][0])
    assert abs(got - ref) <= 1e-11 * abs(ref)
