"""Drop-in ``SpectrumModel`` whose covariance assembly, Cholesky and solve run on the B200.

Mirrors the object API of the reference's ``Starfish.models.SpectrumModel``
(Starfish/models/spectrum_model.py): constructor :126-181, ``grid_params``/``cheb``/``labels`` :183-222,
item access :224-275, ``__call__`` :277-365, ``log_likelihood`` :367-407, parameter dict/vector
get/set :409-493, ``freeze``/``thaw`` :495-590, ``save``/``load`` :592-633, ``train`` :635-696,
``__repr__`` :789-822 — same names, argument meaning, parameter-vector order and error behaviour
(``KeyError`` for unknown parameters, ``ValueError`` for multi-order data / wrong vector length,
``numpy.linalg.LinAlgError`` when the covariance is not positive definite).

What differs is where the work happens.  Everything from the rank-M emulator term onwards
(:334-363, :399-405) is one call into libsfb200 (``LikelihoodEngine``); the dense N×N covariance only
comes back to the host when ``model()`` is asked for it.  ``log_likelihood_batch(P)`` is new: it evaluates
a whole ensemble of parameter vectors in one GPU pass and plugs into ``emcee.EnsembleSampler(...,
vectorize=True)``.  The spectral transforms and the emulator's GP predictive upstream of the path
(:287-332) run on the device as well (csrc/upstream.cu); there is no host path.
``solver="structured"`` evaluates the same likelihood through the banded-plus-low-rank structure of the
covariance (csrc/band.cu) instead of the dense N×N Cholesky.
"""
from __future__ import annotations

import logging
from collections import deque
from typing import Optional, Sequence

import numpy as np
from scipy.optimize import minimize

from .constants import JITTER
from .paramtree import ParamTree
from .transforms import resample
from .utils import calculate_dv, create_log_lam_grid

_TRANSFORM_KEYS = ("vz", "vsini", "Av", "Rv", "log_scale", "global_cov", "local_cov", "cheb")
_GLOBAL_KEYS = ("log_amp", "log_ls")
_LOCAL_KEYS = ("mu", "log_amp", "log_sigma")


class _CachedKernel:
    """What ``model._glob_cov`` / ``model._loc_cov`` hold: the (exponentiated) hyper-parameters the cached
    kernel was built from.  Behaves like the reference's cached N×N array where it is looked at
    (``.shape``, ``numpy.asarray``), materialising the matrix on the GPU only on demand."""

    def __init__(self, model, glob=None, loc=None):
        self._model = model
        self.glob = glob            # (amp, ls) or None
        self.loc = loc              # array [K,3] or None
        n = len(model.data.wave)
        self.shape = (n, n)

    def __array__(self, dtype=None, copy=None):
        # the kernel-only matrix, built directly (zero noise term) — bit-identical to what
        # kernels.global_covariance_matrix / local_covariance_matrix return, as the reference's cache is
        from .kernels import _build

        out = _build(np.asarray(self._model.data.wave, dtype=np.float64),
                     None if self.glob is None else np.array([self.glob], dtype=np.float64),
                     None if self.loc is None else np.asarray(self.loc, dtype=np.float64)[None],
                     False, self._model.device)
        return out if dtype is None else out.astype(dtype)


class SpectrumModel:
    """A single-order spectrum model evaluated on the GPU (see module docstring)."""

    _PARAMS = list(_TRANSFORM_KEYS)
    _GLOBAL_PARAMS = list(_GLOBAL_KEYS)
    _LOCAL_PARAMS = list(_LOCAL_KEYS)

    def __init__(self, emulator, data, grid_params: Sequence[float], max_deque_len: int = 100, norm=False,
                 name: str = "SpectrumModel", device: int = 0, emulator_term: str = "reference",
                 solver: str = "dense", **params):
        if isinstance(emulator, str) or isinstance(data, str):
            raise NotImplementedError("loading from HDF5 paths needs h5py; pass in-memory Emulator/Spectrum")
        if len(data) > 1:
            raise ValueError("Multiple orders detected in data, please use EchelleModel")
        if emulator_term not in ("reference", "paper"):
            raise ValueError("emulator_term must be 'reference' (XᵀΣ_w⁻¹X, as coded) or 'paper' (XᵀΣ_wX)")
        if solver not in ("dense", "dense_i8", "structured"):
            raise ValueError("solver must be 'dense', 'dense_i8' or 'structured'")
        self.solver = solver
        self.emulator = emulator
        self.data_name = data.name
        self.data = data[0]
        self.device = device
        self.emulator_term = emulator_term

        # internal fine log-λ grid on which broadening happens (power-of-two length)
        dv = calculate_dv(self.data.wave)
        self.min_dv_wave = create_log_lam_grid(dv, self.emulator.wl.min(), self.emulator.wl.max())["wl"]
        self.bulk_fluxes = resample(self.emulator.wl, self.emulator.bulk_fluxes, self.min_dv_wave)
        self.residuals = deque(maxlen=max_deque_len)

        params = dict(params)
        if "cheb" in params:  # c0 is pinned to 1; stored coefficients are c1, c2, ... (re-inserted last,
            # which fixes the position of cheb:* in `labels` exactly as the reference does)
            coeffs = params.pop("cheb")
            params["cheb"] = {str(i + 1): c for i, c in enumerate(coeffs)}
        self.params = ParamTree(params)
        self.frozen = []
        self.name = name
        self.norm = norm
        self.n_grid_params = len(grid_params)
        self.grid_params = grid_params

        self._lnprob = None
        self._glob_cov = None
        self._loc_cov = None
        self._log_scale = params.get("log_scale", None)
        self._engine = None
        self._static_sig = None
        self._model_sig = None
        self.log = logging.getLogger(self.__class__.__name__)

    # ------------------------------------------------------------------------------------------------
    # parameter bookkeeping
    # ------------------------------------------------------------------------------------------------
    @property
    def grid_params(self):
        return np.array([self.params[k] for k in self.emulator.param_names])

    @grid_params.setter
    def grid_params(self, values):
        for key, val in zip(self.emulator.param_names, values):
            if key not in self.frozen:
                self.params[key] = val

    @property
    def cheb(self):
        return np.array(self.params["cheb"].values())

    @cheb.setter
    def cheb(self, values):
        if "cheb" in self.frozen:
            return
        for key, val in zip(self.params["cheb"], values):
            if key not in self.frozen:
                self.params["cheb"][key] = val

    @property
    def labels(self):
        return tuple(self.get_param_dict(flat=True).keys())

    def __getitem__(self, key):
        if key == "cheb":
            return list(self.params[key].values())
        return self.params[key]

    def __setitem__(self, key, value):
        if ":" not in key:
            if key == "cheb":
                self.params[key] = {str(i + 1): c for i, c in enumerate(value)}
            elif key in self._PARAMS or key in self.emulator.param_names:
                self.params[key] = value
            else:
                raise KeyError(f"{key} not recognized")
            return
        group, rest = key.split(":", 1)
        leaf = rest.rsplit(":", 1)[-1]
        if group == "global_cov" and leaf in self._GLOBAL_PARAMS:
            self.params[key] = value
        elif group == "local_cov" and leaf in self._LOCAL_PARAMS:
            self.params[key] = value
        elif group == "cheb":
            idx = int(rest)
            if idx == 0:
                raise KeyError("cannot change constant Chebyshev term")
            if "cheb" in self.params:  # fill skipped orders with zeros
                for i in range(len(self.params["cheb"]) + 1, idx + 1):
                    self.params[f"cheb:{i}"] = 0
            self.params[key] = value
        else:
            raise KeyError(f"{key} not recognized")

    def __delitem__(self, key):
        if key not in self.params:
            raise KeyError(f"{key} not in params")
        if key in ("global_cov", "local_cov"):
            if key == "global_cov":
                self._glob_cov = None
            else:
                self._loc_cov = None
            self.frozen = [k for k in self.frozen if not k.startswith(key)]
        del self.params[key]
        if key in self.frozen:
            self.frozen.remove(key)

    def get_param_dict(self, flat: bool = False):
        thawed = ParamTree()
        for key, val in self.params.items():
            if key not in self.frozen:
                thawed[key] = val
        if not flat:
            return thawed.as_dict()
        return thawed

    def set_param_dict(self, params):
        for key, val in ParamTree(params).items():
            if key not in self.frozen:
                self.params[key] = val

    def get_param_vector(self):
        return np.array(self.get_param_dict(flat=True).values())

    def set_param_vector(self, params):
        labels = self.labels
        if len(params) != len(labels):
            raise ValueError("Param Vector does not match length of thawed parameters")
        self.set_param_dict(dict(zip(labels, params)))

    def _group_flat_keys(self, name):
        sub = self.params.as_dict()[name]
        if name == "local_cov":
            return [f"local_cov:{i}:{k}" for i, kern in enumerate(sub) for k in kern]
        return [f"{name}:{k}" for k in sub]

    def freeze(self, names):
        names = [str(n) for n in np.atleast_1d(names)]
        if names[0] == "all":
            for key in self.labels:
                if key not in self.frozen:
                    self.frozen.append(key)
            for group in ("global_cov", "local_cov", "cheb"):
                if group in self.params:
                    self.frozen.append(group)
            return
        for name in names:
            if name in ("global_cov", "local_cov", "cheb"):
                self.frozen.append(name)
                if name == "global_cov":
                    self._glob_cov = None
                elif name == "local_cov":
                    self._loc_cov = None
                for flat in self._group_flat_keys(name):
                    if flat not in self.frozen:
                        self.frozen.append(flat)
            elif name not in self.frozen and name in self.params:
                self.frozen.append(name)

    def thaw(self, names):
        names = [str(n) for n in np.atleast_1d(names)]
        if names[0] == "all":
            self.frozen = []
            return
        for name in names:
            if name in ("global_cov", "local_cov", "cheb"):
                self.frozen.remove(name)
                for flat in self._group_flat_keys(name):
                    self.frozen.remove(flat)
            elif name in self.frozen:
                self.frozen.remove(name)

    # ------------------------------------------------------------------------------------------------
    # kernel hyper-parameters (host bookkeeping only)
    # ------------------------------------------------------------------------------------------------
    def _kernel_hyper(self):
        """(glob (amp, ls) | None, loc [K,3] | None) honouring the frozen-group cache semantics of
        spectrum_model.py:341-363: a group is re-read from the parameters on every call unless it is
        frozen and already cached."""
        if "global_cov" in self.params:
            if "global_cov" not in self.frozen or self._glob_cov is None:
                self._glob_cov = _CachedKernel(self, glob=(float(np.exp(self.params["global_cov:log_amp"])),
                                                           float(np.exp(self.params["global_cov:log_ls"]))))
        if "local_cov" in self.params:
            if "local_cov" not in self.frozen or self._loc_cov is None:
                rows = [(np.exp(k["log_amp"]), k["mu"], np.exp(k["log_sigma"]))
                        for k in self.params.as_dict()["local_cov"]]
                self._loc_cov = _CachedKernel(self, loc=np.array(rows, dtype=np.float64).reshape(-1, 3))
        glob = self._glob_cov.glob if self._glob_cov is not None else None
        loc = self._loc_cov.loc if self._loc_cov is not None else None
        return glob, loc

    # ------------------------------------------------------------------------------------------------
    # GPU plumbing
    # ------------------------------------------------------------------------------------------------
    def _n_local(self):
        return len(self.params.as_dict()["local_cov"]) if "local_cov" in self.params else 0

    def _n_cheb(self):
        return len(self.params["cheb"].keys()) if "cheb" in self.params else 0

    def _model_flags(self):
        from . import _lib

        flags = 0
        if "vsini" in self.params:
            flags |= _lib.MODEL_VSINI
        if "vz" in self.params:
            flags |= _lib.MODEL_VZ
        if "log_scale" in self.params:
            flags |= _lib.MODEL_LOG_SCALE
        if self.norm:
            flags |= _lib.MODEL_NORM
        if self.emulator_term == "paper":
            flags |= _lib.MODEL_PAPER_TERM
        if "Av" in self.params:
            flags |= _lib.MODEL_AV
        return flags

    def _get_engine(self, n_walkers, n_local=None):
        from .engine import LikelihoodEngine

        n = len(self.data.wave)
        m = self.emulator.ncomps
        k = max(1, n_local if n_local is not None else self._n_local())
        eng = self._engine
        if eng is None or eng.N != n or eng.M != m or eng.K < k or eng.B < n_walkers:
            if eng is not None:
                k = max(k, eng.K)
                n_walkers = max(n_walkers, eng.B)
                eng.close()
            self._engine = eng = LikelihoodEngine(n, m, k, n_walkers, device=self.device)
            self._static_sig = None
            self._model_sig = None
        if eng.solver != self.solver:
            eng.set_solver(self.solver)
        return eng

    def _sync_static(self, eng):
        """Re-upload wave/σ/flux when the user swapped or edited them (tests overwrite ``data._flux``)."""
        cur = (np.array(self.data.wave, dtype=np.float64), np.array(self.data.sigma, dtype=np.float64),
               np.array(self.data.flux, dtype=np.float64))
        old = self._static_sig
        if old is None or any(a.shape != b.shape or not np.array_equal(a, b) for a, b in zip(cur, old)):
            eng.set_data(*cur)
            self._static_sig = cur

    def _sync_model(self, eng):
        """Upload the static model tables (fine-grid bulk fluxes, emulator GP) when they changed: new
        engine, re-trained emulator hyper-parameters, a parameter group added or removed."""
        emu = self.emulator
        # content, not identity: in-place edits (emu.w_hat[:] = ..., model.bulk_fluxes[...] = ...) must be seen, as the
        # reference re-reads these arrays on every call (crc32 of ~1 MB: 0.3 ms per call)
        import zlib

        def crc(a):
            return zlib.crc32(np.ascontiguousarray(a, dtype=np.float64).view(np.uint8))

        sig = (id(eng), crc(self.bulk_fluxes), crc(self.min_dv_wave), crc(emu.v11), crc(emu.w_hat), crc(emu.grid_points),
               emu.get_param_vector().tobytes(), self._model_flags(), self._n_cheb())
        if getattr(self, "_model_sig", None) != sig:
            eng.set_model(self.min_dv_wave, self.bulk_fluxes, emu.grid_points, emu.variances, emu.lengthscales,
                          emu.v11, emu.w_hat, ncheb_max=self._n_cheb(), flags=self._model_flags())
            self._model_sig = sig

    @staticmethod
    def _check_finite(*arrays):
        for a in arrays:
            if not np.all(np.isfinite(a)):
                raise ValueError("array must not contain infs or NaNs")

    def _columns(self, P=None):
        """{flat parameter key: value array [B]} — thawed parameters from the columns of ``P`` (in
        ``labels`` order), frozen ones from the model; ``P=None`` means the model's current state (B=1)."""
        if P is None:
            return 1, {k: np.array([v], dtype=np.float64) for k, v in self.params.items()}
        labels = self.labels
        B = P.shape[0]
        cols = {k: np.full(B, v, dtype=np.float64) for k, v in self.params.items()}
        for j, key in enumerate(labels):
            cols[key] = np.ascontiguousarray(P[:, j])
        return B, cols

    def _theta(self, B, cols):
        """Pack the per-walker parameters the device upstream stage reads (include/sfb200.h, sfb_upstream):
        [grid params | vsini | vz | log_scale | norm | c1..c_ncheb | Av]."""
        D, nc = len(self.emulator.param_names), self._n_cheb()
        th = np.zeros((B, D + 4 + nc + (1 if "Av" in self.params else 0)))
        if "Av" in self.params:
            th[:, D + 4 + nc] = cols["Av"]
        for d, name in enumerate(self.emulator.param_names):
            th[:, d] = cols[name]
        if "vsini" in self.params:
            th[:, D] = cols["vsini"]
        if "vz" in self.params:
            th[:, D + 1] = cols["vz"]
        if "log_scale" in self.params:
            th[:, D + 2] = cols["log_scale"]
        th[:, D + 3] = self.emulator.norm_factor(th[:, :D]) if self.norm else 1.0
        for i, key in enumerate(self.params["cheb"].keys() if nc else []):
            th[:, D + 4 + i] = cols[f"cheb:{key}"]
        return th

    def _hyper_rows(self, B, cols):
        """-> (glob [Bh,2], nloc [Bh] int32, loc [Bh,K,3], shared) with the frozen-group cache semantics of
        spectrum_model.py:341-363: a frozen group keeps the values it was first evaluated with."""
        g_cached, l_cached = self._kernel_hyper()
        K = max(1, self._n_local())
        g_thawed = "global_cov" in self.params and "global_cov" not in self.frozen
        l_thawed = "local_cov" in self.params and "local_cov" not in self.frozen
        shared = not (g_thawed or l_thawed)
        Bh = 1 if shared else B
        glob = np.zeros((Bh, 2))
        glob[:, 1] = 1.0
        if g_thawed and not shared:
            glob[:, 0] = np.exp(cols["global_cov:log_amp"])
            glob[:, 1] = np.exp(cols["global_cov:log_ls"])
        elif g_cached is not None:
            glob[:] = g_cached
        loc = np.zeros((Bh, K, 3))
        nloc = np.zeros(Bh, dtype=np.int32)
        if l_thawed and not shared:
            nk = self._n_local()
            for k in range(nk):
                loc[:, k, 0] = np.exp(cols[f"local_cov:{k}:log_amp"])
                loc[:, k, 1] = cols[f"local_cov:{k}:mu"]
                loc[:, k, 2] = np.exp(cols[f"local_cov:{k}:log_sigma"])
            nloc[:] = nk
        elif l_cached is not None and len(l_cached):
            loc[:, :len(l_cached)] = l_cached
            nloc[:] = len(l_cached)
        return glob, nloc, loc, shared

    @staticmethod
    def _pad_loc(eng, loc):
        if loc.shape[1] == eng.K:
            return loc
        pad = np.zeros((loc.shape[0], eng.K, 3))
        pad[:, :loc.shape[1]] = loc
        return pad

    def _check_transforms(self, cols):
        if "vsini" in self.params and np.any(cols["vsini"] <= 0):
            raise ValueError("vsini must be positive")

    # ------------------------------------------------------------------------------------------------
    # evaluation
    # ------------------------------------------------------------------------------------------------
    def __call__(self):
        """-> (flux[N], cov[N,N]) like the reference; transforms, emulator and covariance all run on the GPU."""
        glob, loc = self._kernel_hyper()
        eng = self._get_engine(1)
        self._sync_static(eng)
        B, cols = self._columns()
        self._check_transforms(cols)
        gp = self.grid_params
        if np.any(gp < self.emulator.min_params) or np.any(gp > self.emulator.max_params):
            raise ValueError("Querying emulator outside of original parameter range.")
        self._sync_model(eng)
        up = eng.upstream(self._theta(B, cols), self._n_cheb())
        if int(up["status"].cpu().numpy()[0]) != 0:
            raise np.linalg.LinAlgError("emulator weights covariance is not positive definite")
        flux = up["flux"][0].cpu().numpy()
        self._log_scale = float(up["log_scale"].cpu().numpy()[0])
        C = eng.build_covariance(up["X"], up["A"], glob=None if glob is None else np.array([glob]),
                                 loc=None if loc is None or len(loc) == 0 else loc[None], n_walkers=1)
        return flux, C[0].cpu().numpy()

    def _prior_rows(self, B, cols, priors):
        """Σ prior.logpdf per row (spectrum_model.py:387-395); priors on 'cheb' see the coefficient list."""
        lp = np.zeros(B)
        if priors is None:
            return lp
        for key, prior in priors.items():
            if key not in self.params:
                continue
            if key == "cheb":
                vals = np.stack([cols[f"cheb:{k}"] for k in self.params["cheb"].keys()], axis=1)
                lp += np.array([np.sum(prior.logpdf(list(v))) for v in vals])
                continue
            try:
                term = np.asarray(prior.logpdf(cols[key]), dtype=np.float64)
                if term.shape != (B,):
                    raise ValueError
            except Exception:
                term = np.array([prior.logpdf(v) for v in cols[key]], dtype=np.float64)
            lp += term
        return lp

    def log_likelihood(self, priors: Optional[dict] = None) -> float:
        prior_lp = 0
        if priors is not None:
            for key, prior in priors.items():
                if key in self.params:
                    prior_lp += prior.logpdf(self[key])
        if not np.isfinite(prior_lp):
            return -np.inf
        B, cols = self._columns()
        self._check_transforms(cols)
        gp = self.grid_params
        if np.any(gp < self.emulator.min_params) or np.any(gp > self.emulator.max_params):
            raise ValueError("Querying emulator outside of original parameter range.")
        theta = self._theta(B, cols)
        glob, nloc, loc, shared = self._hyper_rows(B, cols)
        self._check_finite(theta, glob, loc)
        eng = self._get_engine(1)
        self._sync_static(eng)
        self._sync_model(eng)
        loc = self._pad_loc(eng, loc)
        lnL, info = np.empty(1), np.empty(1, dtype=np.int32)
        resid, lsc = np.empty((1, eng.N)), np.empty(1)
        eng.log_likelihood_params_host(theta, self._n_cheb(), glob, nloc, loc, lnL, info, shared_hyper=shared,
                                       resid_out=resid, log_scale_out=lsc)
        self._log_scale = float(lsc[0])
        self.residuals.append(resid[0])
        if info[0] < 0:
            raise np.linalg.LinAlgError("emulator weights covariance is not positive definite")
        if info[0] > 0:
            raise np.linalg.LinAlgError(f"{int(info[0])}-th leading minor of the array is not positive definite")
        self._lnprob = float(lnL[0])
        return self._lnprob + prior_lp

    def log_likelihood_batch(self, P, priors: Optional[dict] = None, on_not_pd: str = "-inf",
                             store_residual: bool = False):
        """Log-probability of B parameter vectors (rows of ``P``, columns in ``self.labels`` order) in ONE
        GPU pass: B×ndim numbers go up, B log-likelihoods come back; transforms, emulator, covariance,
        Cholesky and solve all run on the device.  Plugs into ``emcee.EnsembleSampler(..., vectorize=True)``.

        Rows whose priors are non-finite, whose parameters are non-finite or whose grid parameters fall
        outside the emulator grid get ``-inf`` without any GPU work (mirrors spectrum_model.py:394-395 and
        the priors' usual role); rows whose covariance is not positive definite get ``-inf``
        (``on_not_pd='-inf'``) or raise ``LinAlgError`` (``on_not_pd='raise'``).  The model's own parameters
        are left unchanged; ``store_residual`` appends the residual of the last evaluated row to
        ``self.residuals``.
        """
        P = np.atleast_2d(np.asarray(P, dtype=np.float64))
        if P.shape[1] != len(self.labels):
            raise ValueError("Param Vector does not match length of thawed parameters")
        B, cols = self._columns(P)
        out = np.full(B, -np.inf)
        if B == 0:
            return out
        lp = self._prior_rows(B, cols, priors)
        theta = self._theta(B, cols)
        D = len(self.emulator.param_names)
        ok = np.isfinite(lp) & np.all(np.isfinite(theta), axis=1)
        ok &= ~(np.any(theta[:, :D] < self.emulator.min_params, axis=1) |
                np.any(theta[:, :D] > self.emulator.max_params, axis=1))
        saved_cache = (self._glob_cov, self._loc_cov)
        try:
            glob, nloc, loc, shared = self._hyper_rows(B, cols)
        finally:
            self._glob_cov, self._loc_cov = saved_cache
        if not shared:
            ok &= np.all(np.isfinite(glob), axis=1) & np.all(np.isfinite(loc.reshape(B, -1)), axis=1)
        rows = np.flatnonzero(ok)
        if rows.size == 0:
            return out
        self._check_transforms({k: v[rows] for k, v in cols.items()})
        theta = np.ascontiguousarray(theta[rows])
        if not shared:
            glob, nloc, loc = (np.ascontiguousarray(a[rows]) for a in (glob, nloc, loc))
        nb = rows.size
        eng = self._get_engine(nb)
        self._sync_static(eng)
        self._sync_model(eng)
        loc = self._pad_loc(eng, loc)
        lnL, info = np.empty(nb), np.empty(nb, dtype=np.int32)
        resid = np.empty((nb, eng.N)) if store_residual else None
        eng.log_likelihood_params_host(theta, self._n_cheb(), glob, nloc, loc, lnL, info, shared_hyper=shared,
                                       resid_out=resid)
        if store_residual:
            self.residuals.append(resid[-1])
        if on_not_pd == "raise" and (info != 0).any():
            bad = int(np.flatnonzero(info != 0)[0])
            raise np.linalg.LinAlgError(f"walker {rows[bad]}: covariance not positive definite (info={int(info[bad])})")
        good = info == 0
        out[rows[good]] = lnL[good] + lp[rows][good]
        return out

    # ------------------------------------------------------------------------------------------------
    # persistence, optimisation, display
    # ------------------------------------------------------------------------------------------------
    def save(self, filename, metadata=None):
        import toml

        meta = {"name": self.name, "data": self.data_name}
        if self.emulator.name is not None:
            meta["emulator"] = self.emulator.name
        if metadata is not None:
            meta.update(metadata)
        doc = {"parameters": self.params.as_dict(), "frozen": self.frozen, "metadata": meta}
        with open(filename, "w") as fh:
            toml.dump(doc, fh, encoder=toml.TomlNumpyEncoder(doc.__class__))
        self.log.info(f"Saved current state at {filename}")

    def load(self, filename):
        import toml

        with open(filename, "r") as fh:
            doc = toml.load(fh)
        self.params = ParamTree(doc["parameters"])
        self.frozen = doc["frozen"]

    def train(self, priors: Optional[dict] = None, **kwargs):
        """MAP estimate with ``scipy.optimize.minimize`` (Nelder-Mead by default), as the reference."""
        priors = {} if priors is None else priors
        for key, val in priors.items():
            if key not in self.params and not key.startswith("cheb"):
                raise ValueError(f"Invalid priors. {key} not a vlid key.")
            if not callable(getattr(val, "logpdf", None)):
                raise ValueError(f"Invalid priors. {key} does not have a `logpdf` method")
            lp = val.logpdf(self[key])
            if not np.isfinite(lp):
                raise RuntimeError(f"{key}'s logpdf evaluated to {lp}")

        def nll(P):
            self.set_param_vector(P)
            return -self.log_likelihood(priors)

        opts = {"method": "Nelder-Mead"}
        opts.update(kwargs)
        soln = minimize(nll, self.get_param_vector(), **opts)
        if soln.success:
            self.set_param_vector(soln.x)
        return soln

    def __repr__(self):
        lines = [self.name, "-" * len(self.name), f"Data: {self.data_name}", f"Emulator: {self.emulator.name}",
                 f"Log Likelihood: {self._lnprob}", "", "Parameters"]
        thawed = self.get_param_dict(flat=True)
        top = []
        for key in thawed.keys():
            head = key.split(":", 1)[0]
            if head not in top:
                top.append(head)
        nested = self.get_param_dict()
        for key in top:
            value = nested[key]
            if key == "global_cov":
                lines.append("  global_cov:")
                lines += [f"    {k}: {v}" for k, v in value.items()]
            elif key == "local_cov":
                lines.append("  local_cov:")
                kerns = value.values() if isinstance(value, dict) else value
                for i, kern in enumerate(kerns):
                    lines.append(f"    {i}: " + ", ".join(f"{k}: {v}" for k, v in kern.items()))
            elif key == "cheb":
                lines.append(f"  cheb: {list(value.values())}")
            else:
                lines.append(f"  {key}: {value}")
        if "log_scale" not in self.params:
            lines.append(f"  log_scale: {self._log_scale} (fit)")
        shown = [k for k in self.frozen if k not in ("global_cov", "local_cov")]
        if self.frozen:
            lines += ["", "Frozen Parameters"] + [f"  {k}: {self[k]}" for k in shown]
        return "\n".join(lines)
