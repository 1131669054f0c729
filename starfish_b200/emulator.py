"""Spectral-emulator front end: produces the hot path's inputs ``(weights[M], weights_cov[M,M])``.

Host side (numpy), in-memory only.  Mirrors the call surface of the reference's
``Starfish.emulator.Emulator`` that ``SpectrumModel`` consumes:

* constructor arguments and attributes (``wl``, ``eigenspectra``, ``flux_mean``, ``flux_std``,
  ``grid_points``, ``param_names``, ``min_params``/``max_params``, ``ncomps``, ``v11``, ``w_hat``,
  ``lambda_xi``/``variances``/``lengthscales`` hyper-parameter properties) — Starfish/emulator/emulator.py:69-183
* ``__call__(params, full_cov=True, reinterpret_batch=False) -> (mu, cov)`` — :330-394, including the
  ``ValueError`` for out-of-grid queries (:377-378)
* ``bulk_fluxes`` (:396-402), ``norm_factor`` (:427-442), ``log_likelihood`` (:602-619),
  ``get/set_param_dict``, ``get/set_param_vector`` (:543-600)

SURVEY §8 row f1: the reference re-solves the (M·G)² system ``v11`` twice per call with ``dgesv``
(:387-388).  Here the LU factors of ``v11`` and ``v11⁻¹·ŵ`` are cached and refreshed only when the
hyper-parameters change, and ``predict_batch`` evaluates many walkers in one call.  HDF5 load/save,
PCA ``from_grid`` and training I/O stay with the reference (out of scope).
"""
from __future__ import annotations

import warnings
from typing import Optional, Sequence

import numpy as np
from scipy.interpolate import LinearNDInterpolator
from scipy.linalg import block_diag, cho_factor, cho_solve, lu_factor, lu_solve
from scipy.spatial.distance import cdist

from .utils import calculate_dv


def rbf_kernel(X, Z, variance, lengthscale):
    """σ²·exp(−½ (x−z)ᵀ Λ⁻¹ (x−z)) for all pairs (Starfish/emulator/kernels.py:5-26)."""
    return variance * np.exp(-0.5 * cdist(X / lengthscale, Z / lengthscale, "sqeuclidean"))


def batch_kernel(X, Z, variances, lengthscales):
    """Block-diagonal stack of one RBF block per PCA component (Starfish/emulator/kernels.py:29-49)."""
    return block_diag(*[rbf_kernel(X, Z, v, l) for v, l in zip(variances, lengthscales)])


def phi_squared(eigenspectra, n_grid):
    """ΦᵀΦ for Φ = eigenspectra ⊗ I_G without forming Φ (Starfish/emulator/_utils.py:28-48)."""
    eig = np.asarray(eigenspectra)
    return np.kron(eig @ eig.T, np.eye(n_grid))


class Emulator:
    def __init__(
        self,
        grid_points,
        param_names: Sequence[str],
        wavelength,
        weights,
        eigenspectra,
        w_hat,
        flux_mean,
        flux_std,
        factors,
        lambda_xi: float = 1.0,
        variances=None,
        lengthscales=None,
        name: Optional[str] = None,
    ):
        self.grid_points = np.asarray(grid_points, dtype=np.float64)
        self.param_names = param_names
        self.wl = np.asarray(wavelength, dtype=np.float64)
        self.weights = weights
        self.eigenspectra = np.asarray(eigenspectra, dtype=np.float64)
        self.flux_mean = np.asarray(flux_mean, dtype=np.float64)
        self.flux_std = np.asarray(flux_std, dtype=np.float64)
        self.factors = np.asarray(factors, dtype=np.float64)
        self.factor_interpolator = LinearNDInterpolator(self.grid_points, self.factors, rescale=True)
        self.dv = calculate_dv(self.wl)
        self.ncomps = self.eigenspectra.shape[0]
        self.name = name
        self.hyperparams = {}
        self.lambda_xi = lambda_xi
        self.variances = variances if variances is not None else 1e4 * np.ones(self.ncomps)
        axes = [np.unique(col) for col in self.grid_points.T]
        self._grid_sep = np.array([np.diff(ax).max() for ax in axes])
        if lengthscales is None:
            lengthscales = np.tile(3 * self._grid_sep, (self.ncomps, 1))
        self.lengthscales = lengthscales
        self.min_params = self.grid_points.min(axis=0)
        self.max_params = self.grid_points.max(axis=0)
        self.iPhiPhi = np.linalg.inv(phi_squared(self.eigenspectra, self.grid_points.shape[0]))
        self.w_hat = np.asarray(w_hat, dtype=np.float64)
        self._trained = False
        self._refresh_v11()

    # -- hyper-parameters (stored as logs, exactly like the reference) -----------------------
    @property
    def lambda_xi(self) -> float:
        return np.exp(self.hyperparams["log_lambda_xi"])

    @lambda_xi.setter
    def lambda_xi(self, value):
        self.hyperparams["log_lambda_xi"] = np.log(value)

    @property
    def variances(self):
        return np.exp([v for k, v in self.hyperparams.items() if k.startswith("log_variance:")])

    @variances.setter
    def variances(self, values):
        for i, v in enumerate(values):
            self.hyperparams[f"log_variance:{i}"] = np.log(v)

    @property
    def lengthscales(self):
        vals = [v for k, v in self.hyperparams.items() if k.startswith("log_lengthscale:")]
        return np.exp(vals).reshape(self.ncomps, -1)

    @lengthscales.setter
    def lengthscales(self, values):
        for i, row in enumerate(values):
            for j, ls in enumerate(row):
                self.hyperparams[f"log_lengthscale:{i}:{j}"] = np.log(ls)

    def __getitem__(self, key):
        return self.hyperparams[key]

    def get_param_dict(self):
        return self.hyperparams

    def set_param_dict(self, params):
        for key, val in params.items():
            if key in self.hyperparams:
                self.hyperparams[key] = val
        self._refresh_v11()

    def get_param_vector(self):
        return np.array(list(self.hyperparams.values()))

    def set_param_vector(self, params):
        if len(params) != len(self.hyperparams):
            raise ValueError("params must match length of parameters (get_param_vector())")
        self.set_param_dict(dict(zip(self.hyperparams.keys(), params)))

    # -- cached factorisation (row f1) --------------------------------------------------------
    def _refresh_v11(self):
        self.v11 = self.iPhiPhi / self.lambda_xi + batch_kernel(
            self.grid_points, self.grid_points, self.variances, self.lengthscales
        )
        self._v11_lu = None
        self._v11_src = None

    def _factor(self):
        # users may assign self.v11 / self.w_hat directly; key the cache on object identity
        key = (id(self.v11), id(self.w_hat))
        if self._v11_lu is None or self._v11_src != key:
            self._v11_lu = lu_factor(self.v11)
            self._alpha = lu_solve(self._v11_lu, self.w_hat)
            self._v11_src = key
        return self._v11_lu

    # -- GP predictive -------------------------------------------------------------------------
    def __call__(self, params, full_cov: bool = True, reinterpret_batch: bool = False):
        params = np.atleast_2d(params)
        if full_cov and reinterpret_batch:
            raise ValueError("Cannot reshape the full_covariance matrix for many parameters.")
        if not self._trained:
            warnings.warn("This emulator has not been trained and therefore is not reliable. "
                          "call emulator.train() to train.")
        if np.any(params < self.min_params) or np.any(params > self.max_params):
            raise ValueError("Querying emulator outside of original parameter range.")
        lu = self._factor()
        v12 = batch_kernel(self.grid_points, params, self.variances, self.lengthscales)
        v22 = batch_kernel(params, params, self.variances, self.lengthscales)
        mu = v12.T @ self._alpha
        cov = v22 - v12.T @ lu_solve(lu, v12)
        if not full_cov:
            cov = np.diag(cov)
        if reinterpret_batch:
            mu = mu.reshape(-1, self.ncomps, order="F").squeeze()
            cov = cov.reshape(-1, self.ncomps, order="F").squeeze()
        return mu, cov

    def predict_batch(self, params):
        """(mu[B,M], cov[B,M,M]) for B parameter rows; rows outside the grid give NaN."""
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        B, M = params.shape[0], self.ncomps
        mu = np.full((B, M), np.nan)
        cov = np.full((B, M, M), np.nan)
        ok = ~(np.any(params < self.min_params, axis=1) | np.any(params > self.max_params, axis=1))
        with warnings.catch_warnings():
            warnings.simplefilter("ignore")
            for b in np.flatnonzero(ok):
                mu[b], cov[b] = self(params[b])
        return mu, cov

    @property
    def bulk_fluxes(self):
        return np.vstack([self.eigenspectra, self.flux_mean, self.flux_std])

    def norm_factor(self, params):
        return self.factor_interpolator(np.asarray(params))

    def log_likelihood(self, device: Optional[int] = None) -> float:
        """−½(log det v11 + ŵᵀ v11⁻¹ ŵ) (Starfish/emulator/emulator.py:602-619).

        ``device=None`` evaluates on the host like the reference; an integer runs the (M·G)² Cholesky and
        the forward solve on that GPU with the same kernels as the spectrum likelihood (``sfb_potrf`` /
        ``sfb_solve_lower``) — worthwhile for emulators with thousands of grid points × components.
        Raises ``numpy.linalg.LinAlgError`` when v11 is not positive definite, as scipy does."""
        if device is None:
            L, low = cho_factor(self.v11)
            logdet = 2 * np.sum(np.log(np.diag(L)))
            return -(logdet + self.w_hat @ cho_solve((L, low), self.w_hat)) / 2
        import torch

        from .engine import LikelihoodEngine

        n = self.v11.shape[0]
        eng = getattr(self, "_gpu_engine", None)
        if eng is None or eng.N != n or eng.device.index != device:
            if eng is not None:
                eng.close()
            self._gpu_engine = eng = LikelihoodEngine(n, 0, 1, 1, device=device, workspace_walkers=2)
        Cm = torch.from_numpy(np.ascontiguousarray(self.v11, dtype=np.float64)[None]).to(eng.device)
        Cm, info, logdet = eng.cho_factor(Cm, return_logdet=True)
        if int(info.cpu()[0]) != 0:
            raise np.linalg.LinAlgError(f"{int(info.cpu()[0])}-th leading minor of the array is not positive definite")
        z = eng.solve_lower(Cm, np.asarray(self.w_hat, dtype=np.float64)[None])
        return float(-(logdet[0] + (z * z).sum()).cpu() / 2)

    def train(self, device: Optional[int] = None, **opt_kwargs):
        """Optimise the GP hyper-parameters with ``scipy.optimize.minimize`` (Nelder-Mead, maxiter 10000 by
        default), mirroring Starfish/emulator/emulator.py:484-524; ``device`` is passed to ``log_likelihood``."""
        from scipy.optimize import minimize

        def nll(P):
            if np.any(~np.isfinite(P)):
                return np.inf
            self.set_param_vector(P)
            if np.any(self.lengthscales < 2 * self._grid_sep):
                return np.inf
            try:
                return -self.log_likelihood(device=device)
            except np.linalg.LinAlgError:
                return np.inf

        kwargs = {"method": "Nelder-Mead", "options": {"maxiter": 10000}}
        kwargs.update(opt_kwargs)
        P0 = self.get_param_vector()
        soln = minimize(nll, P0, **kwargs)
        if soln.success:
            self.set_param_vector(soln.x)
            self._trained = True
        else:
            self.set_param_vector(P0)
        return soln
