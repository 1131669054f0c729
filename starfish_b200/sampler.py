"""Affine-invariant ensemble sampler driving the batched GPU likelihood.

The reference samples with ``emcee.EnsembleSampler`` around a per-walker Python closure
(examples/single.ipynb:458-470: set_param_vector → log_likelihood(priors), one walker at a time).
emcee is not part of this image, and its per-walker call pattern is exactly what the B200 path
removes, so this module provides the same algorithm — Goodman & Weare's stretch move with the
red/blue half-ensemble split emcee uses by default (a = 2) — written around ONE vectorised
log-probability call per half-step: ``log_prob_fn(P[B/2, ndim]) -> lnp[B/2]``.
``SpectrumModel.log_likelihood_batch`` has that signature.  The call surface follows emcee 3
(``run_mcmc``, ``get_chain``, ``get_log_prob``, ``acceptance_fraction``, ``vectorize=True``) so the
reference's notebooks translate line by line; with emcee installed,
``emcee.EnsembleSampler(nwalkers, ndim, model.log_likelihood_batch, vectorize=True)`` works as well.
"""
from __future__ import annotations

from typing import Callable, Optional

import numpy as np


class EnsembleSampler:
    def __init__(self, nwalkers: int, ndim: int, log_prob_fn: Callable, a: float = 2.0, args=(),
                 kwargs=None, vectorize: bool = True, seed: Optional[int] = None):
        if nwalkers < 2 * ndim or nwalkers % 2:
            raise ValueError("nwalkers must be even and at least 2*ndim (as emcee requires)")
        if not vectorize:
            raise ValueError("this sampler drives a batched log-probability; use vectorize=True")
        self.nwalkers, self.ndim, self.a = int(nwalkers), int(ndim), float(a)
        self._fn, self._args, self._kwargs = log_prob_fn, tuple(args), dict(kwargs or {})
        self.rng = np.random.default_rng(seed)
        self.reset()

    def reset(self):
        self._chain, self._lnp = [], []
        self._accepted = np.zeros(self.nwalkers)
        self.iteration = 0
        self.n_calls = 0

    def compute_log_prob(self, coords):
        lnp = np.asarray(self._fn(coords, *self._args, **self._kwargs), dtype=np.float64)
        self.n_calls += 1
        if lnp.shape != (coords.shape[0],):
            raise ValueError("log_prob_fn must return one value per row")
        if np.any(np.isnan(lnp)):
            raise ValueError("log_prob_fn returned NaN")
        return lnp

    def _stretch(self, active, other):
        """Proposal for the walkers in ``active`` using complementary walkers drawn from ``other``."""
        ns = active.shape[0]
        zz = ((self.a - 1.0) * self.rng.random(ns) + 1.0) ** 2.0 / self.a
        partner = other[self.rng.integers(other.shape[0], size=ns)]
        return partner - (partner - active) * zz[:, None], (self.ndim - 1.0) * np.log(zz)

    def sample(self, initial_state, iterations: int = 1, log_prob0=None):
        """Generator over steps; yields (coords[B,ndim], log_prob[B]) after every full ensemble update."""
        p = np.array(initial_state, dtype=np.float64)
        if p.shape != (self.nwalkers, self.ndim):
            raise ValueError("initial_state must have shape (nwalkers, ndim)")
        lnp = self.compute_log_prob(p) if log_prob0 is None else np.array(log_prob0, dtype=np.float64)
        if not np.all(np.isfinite(lnp)):
            raise ValueError("initial state has a non-finite log-probability")
        half = self.nwalkers // 2
        halves = (np.arange(half), np.arange(half, self.nwalkers))
        for _ in range(iterations):
            for first, second in (halves, halves[::-1]):
                q, factors = self._stretch(p[first], p[second])
                new = self.compute_log_prob(q)
                accept = np.log(self.rng.random(first.size)) < factors + new - lnp[first]
                idx = first[accept]
                p[idx], lnp[idx] = q[accept], new[accept]
                self._accepted[idx] += 1
            self.iteration += 1
            self._chain.append(p.copy())
            self._lnp.append(lnp.copy())
            yield p, lnp

    def run_mcmc(self, initial_state, nsteps: int, log_prob0=None, progress: bool = False):
        state = None
        for state in self.sample(initial_state, iterations=nsteps, log_prob0=log_prob0):
            pass
        return state

    def _stack(self, store, discard, thin, flat):
        arr = np.array(store[discard::thin])
        if flat and arr.size:
            arr = arr.reshape((-1,) + arr.shape[2:])
        return arr

    def get_chain(self, discard: int = 0, thin: int = 1, flat: bool = False):
        """[steps, nwalkers, ndim] (or [steps*nwalkers, ndim] when ``flat``)."""
        return self._stack(self._chain, discard, thin, flat)

    def get_log_prob(self, discard: int = 0, thin: int = 1, flat: bool = False):
        return self._stack(self._lnp, discard, thin, flat)

    @property
    def acceptance_fraction(self):
        return self._accepted / max(self.iteration, 1)
