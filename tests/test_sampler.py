"""CPU: the stretch-move ensemble sampler (host glue of row f3) on an analytic target."""
import numpy as np
import pytest

from starfish_b200.sampler import EnsembleSampler


def test_recovers_gaussian_moments_and_batches_calls():
    mean = np.array([1.0, -2.0, 0.5])
    cov = np.array([[1.0, 0.6, 0.0], [0.6, 2.0, -0.3], [0.0, -0.3, 0.5]])
    icov = np.linalg.inv(cov)
    shapes = []

    def lnp(P):
        shapes.append(P.shape)
        d = P - mean
        return -0.5 * np.einsum("bi,ij,bj->b", d, icov, d)

    nw, nd = 32, 3
    s = EnsembleSampler(nw, nd, lnp, seed=1)
    p0 = mean + 0.1 * np.random.default_rng(0).standard_normal((nw, nd))
    s.run_mcmc(p0, 1500)
    assert set(shapes[1:]) == {(nw // 2, nd)}          # two half-ensemble calls per step, nothing per walker
    assert s.n_calls == 1 + 2 * 1500
    flat = s.get_chain(discard=300, flat=True)
    assert flat.shape == (1200 * nw, nd)
    assert np.abs(flat.mean(axis=0) - mean).max() < 0.1
    assert np.abs(np.cov(flat.T) - cov).max() < 0.2
    assert 0.2 < s.acceptance_fraction.mean() < 0.9
    assert s.get_log_prob().shape == (1500, nw)


def test_minus_inf_rows_are_never_accepted_and_errors():
    def lnp(P):
        out = -0.5 * np.sum(P ** 2, axis=1)
        out[P[:, 0] > 1.0] = -np.inf                    # hard wall, like a prior's support
        return out

    s = EnsembleSampler(8, 2, lnp, seed=3)
    p0 = 0.1 * np.random.default_rng(4).standard_normal((8, 2))
    s.run_mcmc(p0, 400)
    assert s.get_chain()[:, :, 0].max() <= 1.0
    with pytest.raises(ValueError):
        EnsembleSampler(3, 2, lnp)                       # odd / too few walkers
    with pytest.raises(ValueError):
        s.run_mcmc(np.full((8, 2), 5.0), 1)              # initial state outside the support
    with pytest.raises(ValueError):
        EnsembleSampler(8, 2, lambda P: np.full(len(P), np.nan)).run_mcmc(p0, 1)
