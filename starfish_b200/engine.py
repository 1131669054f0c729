"""`LikelihoodEngine` — Python owner of one libsfb200 handle (one GPU, one process).

This is the stage boundary of SURVEY §8d: per-walker ``X[M,N]``, ``A[M,M]``, ``model_flux[N]`` and
kernel hyper-parameters in, ``lnL[B]`` / ``info[B]`` out.  torch tensors are used purely as device
memory containers (allocation, ``data_ptr()``, current stream); there are no torch ops on the path.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence

import numpy as np

from . import _lib
from .constants import JITTER


def _torch():
    import torch

    return torch


class LikelihoodEngine:
    """Batched covariance build + Cholesky + solve on one B200.

    Parameters
    ----------
    n_pix, n_comp, max_local, max_walkers : sizes (N, M, Kmax, Bmax) the handle is created for
    device : CUDA device index
    workspace_walkers : number of N×N factorisation slots (0 = automatic)
    """

    def __init__(self, n_pix: int, n_comp: int, max_local: int, max_walkers: int, device: int = 0,
                 workspace_walkers: int = 0):
        torch = _torch()
        if not torch.cuda.is_available():
            raise RuntimeError("starfish_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self._lib = _lib.lib()
        self.N, self.M, self.K, self.B = int(n_pix), int(n_comp), max(int(max_local), 1), int(max_walkers)
        self.device = torch.device("cuda", device)
        torch.cuda.init()
        with torch.cuda.device(self.device):
            torch.empty(1, device=self.device)  # make sure the primary context exists
        h = C.c_void_p()
        rc = self._lib.sfb_create(device, self.N, self.M, self.K, self.B, int(workspace_walkers), C.byref(h))
        if rc != 0:
            raise _lib.SfbError(f"sfb_create failed with status {rc} (N={n_pix}, M={n_comp}, "
                                f"Kmax={max_local}, Bmax={max_walkers})")
        self._h = h
        self._static = None
        self.D, self.model_flags, self.ncheb_max = 0, 0, 0   # set by set_model

    # -- plumbing ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_h", None):
            self._lib.sfb_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, what):
        if rc != 0:
            msg = self._lib.sfb_last_error(self._h)
            raise _lib.SfbError(f"{what} failed ({rc}): {msg.decode() if msg else ''}")

    def _stream(self):
        return C.c_void_p(_torch().cuda.current_stream(self.device).cuda_stream)

    def _dev(self, a, dtype=None, shape=None):
        """numpy / torch (any device) -> contiguous device tensor of the wanted dtype."""
        torch = _torch()
        dtype = dtype or torch.float64
        if a is None:
            return None
        if isinstance(a, torch.Tensor):
            t = a.to(device=self.device, dtype=dtype).contiguous()
        else:
            np_dtype = np.float64 if dtype == torch.float64 else np.int32
            t = torch.from_numpy(np.ascontiguousarray(a, dtype=np_dtype)).to(self.device)
        if shape is not None and tuple(t.shape) != tuple(shape):
            raise ValueError(f"expected shape {tuple(shape)}, got {tuple(t.shape)}")
        return t

    @staticmethod
    def _ptr(t):
        return C.c_void_p(t.data_ptr()) if t is not None else C.c_void_p(0)

    @property
    def workspace_walkers(self) -> int:
        return self._lib.sfb_workspace_walkers(self._h)

    @property
    def padded_n(self) -> int:
        return self._lib.sfb_padded_n(self._h)

    @property
    def launch_count(self) -> int:
        return int(self._lib.sfb_launch_count(self._h))

    def synchronize(self):
        self._check(self._lib.sfb_sync(self._h), "sfb_sync")
        _torch().cuda.synchronize(self.device)

    # -- static data --------------------------------------------------------------------------------
    def set_data(self, wave, sigma, data_flux):
        """Upload the walker-independent arrays (wavelengths, σ, observed flux), each of length N."""
        w = self._dev(wave, shape=(self.N,))
        s = self._dev(sigma, shape=(self.N,))
        f = self._dev(data_flux, shape=(self.N,))
        self._check(self._lib.sfb_set_static(self._h, self._ptr(w), self._ptr(s), self._ptr(f), self._stream()),
                    "sfb_set_static")
        self._static = (w, s, f)  # keep alive until the async copies have run

    # -- hyper-parameter packing ----------------------------------------------------------------------
    def pack_hyper(self, B, glob, nloc, loc, shared):
        """-> device tensors glob[Bh,2], nloc[Bh] (int32), loc[Bh,Kmax,3];  Bh = 1 if shared else B."""
        torch = _torch()
        Bh = 1 if shared else B
        g = np.zeros((Bh, 2)) if glob is None else glob
        g = self._dev(g).reshape(Bh, 2)
        if loc is None:
            l = torch.zeros((Bh, self.K, 3), dtype=torch.float64, device=self.device)
            n = torch.zeros((Bh,), dtype=torch.int32, device=self.device)
        else:
            l = self._dev(loc)
            if l.dim() == 2:
                l = l.unsqueeze(0)
            if l.shape[0] != Bh or l.shape[2] != 3 or l.shape[1] > self.K:
                raise ValueError(f"loc must be [{Bh}, <= {self.K}, 3], got {tuple(l.shape)}")
            kk = l.shape[1]
            if kk < self.K:
                pad = torch.zeros((Bh, self.K, 3), dtype=torch.float64, device=self.device)
                pad[:, :kk] = l
                l = pad
            n = (torch.full((Bh,), kk, dtype=torch.int32, device=self.device) if nloc is None
                 else self._dev(nloc, dtype=torch.int32).reshape(Bh))
        return g.contiguous(), n.contiguous(), l.contiguous()

    def _xa(self, B, X, A):
        if X is None or self.M == 0:
            return None, None
        X = self._dev(X, shape=(B, self.M, self.N))
        A = self._dev(A, shape=(B, self.M, self.M))
        return X, A

    # -- entry points -------------------------------------------------------------------------------
    def build_covariance(self, X, A, glob=None, nloc=None, loc=None, shared_hyper=False,
                         jitter: float = 0.0, n_walkers: Optional[int] = None):
        """C[b] = XᵀAX + diag(σ²+jitter) + K_global + ΣK_local  → device tensor [B,N,N] (both triangles)."""
        torch = _torch()
        B = int(n_walkers if n_walkers is not None else (X.shape[0] if X is not None else np.shape(glob)[0]))
        X, A = self._xa(B, X, A)
        g, n, l = self.pack_hyper(B, glob, nloc, loc, shared_hyper)
        Cm = torch.empty((B, self.N, self.N), dtype=torch.float64, device=self.device)
        self._check(self._lib.sfb_build_cov(self._h, B, self._ptr(X), self._ptr(A), self._ptr(g), self._ptr(n),
                                            self._ptr(l), int(shared_hyper), float(jitter), self._ptr(Cm),
                                            self._stream()), "sfb_build_cov")
        self._keep = (X, A, g, n, l)
        return Cm

    def log_likelihood(self, X, A, model_flux, glob=None, nloc=None, loc=None, shared_hyper=False,
                       return_residuals=False):
        """Device path: inputs numpy or torch; returns device tensors (lnL[B], info[B][, resid[B,N]])."""
        torch = _torch()
        F = self._dev(model_flux)
        if F.dim() == 1:
            F = F.unsqueeze(0)
        B = F.shape[0]
        if F.shape[1] != self.N:
            raise ValueError(f"model_flux must be [B,{self.N}]")
        if B > self.B:
            raise _lib.SfbError(f"batch of {B} walkers exceeds the handle's Bmax={self.B}")
        if B == 0:  # nothing to do (and empty tensors have no device pointer to hand over)
            empty = torch.empty((0,), dtype=torch.float64, device=self.device)
            out = (empty, torch.empty((0,), dtype=torch.int32, device=self.device))
            return out + (torch.empty((0, self.N), dtype=torch.float64, device=self.device),) \
                if return_residuals else out
        X, A = self._xa(B, X, A)
        g, n, l = self.pack_hyper(B, glob, nloc, loc, shared_hyper)
        lnL = torch.empty((B,), dtype=torch.float64, device=self.device)
        info = torch.empty((B,), dtype=torch.int32, device=self.device)
        resid = torch.empty((B, self.N), dtype=torch.float64, device=self.device) if return_residuals else None
        self._check(self._lib.sfb_loglike(self._h, B, self._ptr(X), self._ptr(A), self._ptr(F), self._ptr(g),
                                          self._ptr(n), self._ptr(l), int(shared_hyper), self._ptr(lnL),
                                          self._ptr(info), self._ptr(resid), self._stream()), "sfb_loglike")
        self._keep = (X, A, F, g, n, l)
        return (lnL, info, resid) if return_residuals else (lnL, info)

    def log_likelihood_resident(self, B, X, A, F, g, n, l, lnL, info, shared_hyper=False):
        """Zero-overhead variant for benchmarks: every argument is already a packed device tensor."""
        self._check(self._lib.sfb_loglike(self._h, B, self._ptr(X), self._ptr(A), self._ptr(F), self._ptr(g),
                                          self._ptr(n), self._ptr(l), int(shared_hyper), self._ptr(lnL),
                                          self._ptr(info), C.c_void_p(0), self._stream()), "sfb_loglike")

    def log_likelihood_host(self, X, A, model_flux, glob, nloc, loc, lnL_out, info_out, shared_hyper=False,
                            resid_out=None):
        """End-to-end path on HOST buffers (numpy arrays, ideally views of pinned memory).

        Host→device copies of the inputs and the device→host read of lnL/info happen inside the call,
        pipelined chunk by chunk behind the factorisation.  Arrays must be C-contiguous fp64 / int32 with
        ``loc`` already padded to [Bh, Kmax, 3].
        """
        def hp(a, dt):
            if a is None:
                return C.c_void_p(0)
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
                raise ValueError("host buffers must be C-contiguous numpy arrays of the documented dtype")
            return C.c_void_p(a.ctypes.data)

        B = model_flux.shape[0]
        self._check(self._lib.sfb_loglike_host(self._h, B, hp(X, np.float64), hp(A, np.float64),
                                               hp(model_flux, np.float64), hp(glob, np.float64),
                                               hp(nloc, np.int32), hp(loc, np.float64), int(shared_hyper),
                                               hp(lnL_out, np.float64), hp(info_out, np.int32),
                                               hp(resid_out, np.float64)), "sfb_loglike_host")
        return lnL_out, info_out

    # -- solver choice (row f4) -----------------------------------------------------------------------
    def set_solver(self, solver: str):
        """'dense' (default: blocked fp64 Cholesky of the N×N covariance) or 'structured' (banded Cholesky of
        diag + kernels, rank-M capacitance system; walkers whose band does not fit take the dense path)."""
        code = {"dense": _lib.SOLVER_DENSE, "structured": _lib.SOLVER_STRUCTURED,
                "dense_i8": _lib.SOLVER_DENSE_I8}.get(solver)
        if code is None:
            raise ValueError("solver must be 'dense', 'dense_i8' or 'structured'")
        self._check(self._lib.sfb_set_solver(self._h, code), "sfb_set_solver")

    @property
    def solver(self) -> str:
        return {_lib.SOLVER_STRUCTURED: "structured", _lib.SOLVER_DENSE_I8: "dense_i8"}.get(
            self._lib.sfb_get_solver(self._h), "dense")

    def set_shared_factor(self, on: bool):
        """Frozen-kernel calls (``shared_hyper=True``): factorise the shared S once per call and solve all walkers'
        right-hand sides against it (default), or factorise every walker's full covariance (``on=False``)."""
        self._check(self._lib.sfb_set_shared_factor(self._h, int(bool(on))), "sfb_set_shared_factor")

    @property
    def shared_factor_calls(self) -> int:
        return int(self._lib.sfb_shared_factor_calls(self._h))

    def i8_mma_counts(self):
        """(issued, dense) int8 MMA counts of the dense_i8 trailing update since the last call: products with an
        all-zero digit slab are skipped, `dense` is what a dense digit pattern would have needed."""
        a, b = C.c_ulonglong(0), C.c_ulonglong(0)
        self._check(self._lib.sfb_i8_mma_counts(self._h, C.byref(a), C.byref(b)), "sfb_i8_mma_counts")
        return int(a.value), int(b.value)

    def band_classes(self):
        """{window width: walkers routed to it since creation}; key 0 is the dense fallback."""
        w = (C.c_int * 8)()
        n = (C.c_longlong * 8)()
        k = self._lib.sfb_band_classes(self._h, w, n, 8)
        if k < 0:
            self._check(k, "sfb_band_classes")
        return {int(w[i]): int(n[i]) for i in range(k)}

    # -- multi-GPU: the one exchange per ensemble step --------------------------------------
    @staticmethod
    def comm_unique_id() -> bytes:
        """128-byte NCCL unique id (call on rank 0, hand it to every rank by any host-side means)."""
        buf = C.create_string_buffer(128)
        rc = _lib.lib().sfb_comm_unique_id(buf)
        if rc != 0:
            raise _lib.SfbError(f"sfb_comm_unique_id failed ({rc}): is libnccl.so.2 loadable?")
        return buf.raw

    def comm_init(self, rank: int, world: int, unique_id: bytes):
        """Collective: create this handle's NCCL communicator (ncclCommInitRank inside the library)."""
        if len(unique_id) != 128:
            raise ValueError("unique_id must be the 128 bytes returned by comm_unique_id()")
        self._check(self._lib.sfb_comm_init(self._h, int(rank), int(world), C.c_char_p(unique_id)), "sfb_comm_init")
        self.comm_rank, self.comm_world = int(rank), int(world)

    def allgather_lnl(self, local, out=None):
        """``out[r·n + i] = local[i]`` of rank r — one ncclAllGather of the lnL shard, enqueued on the current
        stream by the library itself (no torch collective on the path).  Shards must have equal length."""
        torch = _torch()
        n = local.numel()
        if out is None:
            out = torch.empty(n * self.comm_world, dtype=torch.float64, device=self.device)
        self._check(self._lib.sfb_allgather_lnL(self._h, self._ptr(local), n, self._ptr(out), self._stream()),
                    "sfb_allgather_lnL")
        return out

    # -- upstream of the covariance (rows f1/f2/f3): parameters in, log-likelihood out --------------
    def set_model(self, fine_wave, bulk_fluxes, grid_points, variances, lengthscales, v11, w_hat,
                  ncheb_max: int = 0, flags: int = 0):
        """Upload the static model data (SpectrumModel.min_dv_wave / bulk_fluxes and the emulator's GP
        tables).  The library derives what it needs from them once: the spectrum of the bulk fluxes, the
        banded inverse of the spline collocation matrix, chol(v11) and L⁻¹ŵ."""
        def h(a, shape):
            a = np.ascontiguousarray(a, dtype=np.float64)
            if a.shape != tuple(shape):
                raise ValueError(f"expected shape {tuple(shape)}, got {a.shape}")
            return a

        fine_wave = np.ascontiguousarray(fine_wave, dtype=np.float64)
        nf = fine_wave.shape[0]
        grid_points = np.ascontiguousarray(grid_points, dtype=np.float64)
        G, D = grid_points.shape
        bulk = h(bulk_fluxes, (self.M + 2, nf))
        var = h(variances, (self.M,))
        ls = h(lengthscales, (self.M, D))
        v11 = h(v11, (self.M * G, self.M * G))
        what = h(w_hat, (self.M * G,))
        self._check(self._lib.sfb_set_model_host(
            self._h, nf, fine_wave.ctypes.data, bulk.ctypes.data, G, D, grid_points.ctypes.data, var.ctypes.data,
            ls.ctypes.data, v11.ctypes.data, what.ctypes.data, int(ncheb_max), int(flags)), "sfb_set_model_host")
        self.D, self.model_flags, self.ncheb_max = D, int(flags), int(ncheb_max)

    def theta_width(self, ncheb):
        return self.D + 4 + int(ncheb) + (1 if self.model_flags & _lib.MODEL_AV else 0)

    def upstream(self, theta, ncheb: int):
        """Device tensors of everything SpectrumModel.__call__ computes before the rank-M term:
        dict(X[B,M,N], A[B,M,M], flux[B,N], log_scale[B], status[B], weights[B,M], weights_cov[B,M,M])."""
        torch = _torch()
        th = self._dev(theta)
        B = th.shape[0]
        if th.dim() != 2 or th.shape[1] != self.theta_width(ncheb):
            raise ValueError(f"theta must be [B,{self.theta_width(ncheb)}]")
        f64 = dict(dtype=torch.float64, device=self.device)
        out = dict(X=torch.empty((B, self.M, self.N), **f64), A=torch.empty((B, self.M, self.M), **f64),
                   flux=torch.empty((B, self.N), **f64), log_scale=torch.empty((B,), **f64),
                   status=torch.empty((B,), dtype=torch.int32, device=self.device),
                   weights=torch.empty((B, self.M), **f64), weights_cov=torch.empty((B, self.M, self.M), **f64))
        if B:
            self._check(self._lib.sfb_upstream(self._h, B, self._ptr(th), int(ncheb), self._ptr(out["X"]),
                                               self._ptr(out["A"]), self._ptr(out["flux"]),
                                               self._ptr(out["log_scale"]), self._ptr(out["status"]),
                                               self._ptr(out["weights"]), self._ptr(out["weights_cov"]),
                                               self._stream()), "sfb_upstream")
        self._keep = (th,)
        return out

    def log_likelihood_params_resident(self, B, theta, ncheb, g, n, l, lnL, info, shared_hyper=False):
        """Benchmark variant: every argument is a packed device tensor."""
        self._check(self._lib.sfb_loglike_params(self._h, B, self._ptr(theta), int(ncheb), self._ptr(g),
                                                 self._ptr(n), self._ptr(l), int(shared_hyper), self._ptr(lnL),
                                                 self._ptr(info), C.c_void_p(0), C.c_void_p(0), self._stream()),
                    "sfb_loglike_params")

    def log_likelihood_params_host(self, theta, ncheb, glob, nloc, loc, lnL_out, info_out, shared_hyper=False,
                                   resid_out=None, log_scale_out=None):
        """The ensemble step on HOST buffers: B×ntheta parameter values and the kernel hyper-parameters go
        up, lnL/info (and optionally the fitted log_scale and the residuals) come back."""
        def hp(a, dt):
            if a is None:
                return C.c_void_p(0)
            if not (isinstance(a, np.ndarray) and a.dtype == dt and a.flags["C_CONTIGUOUS"]):
                raise ValueError("host buffers must be C-contiguous numpy arrays of the documented dtype")
            return C.c_void_p(a.ctypes.data)

        B = theta.shape[0]
        if theta.ndim != 2 or theta.shape[1] != self.theta_width(ncheb):
            raise ValueError(f"theta must be [B,{self.theta_width(ncheb)}]")
        self._check(self._lib.sfb_loglike_params_host(
            self._h, B, hp(theta, np.float64), int(ncheb), hp(glob, np.float64), hp(nloc, np.int32),
            hp(loc, np.float64), int(shared_hyper), hp(lnL_out, np.float64), hp(info_out, np.int32),
            hp(resid_out, np.float64), hp(log_scale_out, np.float64)), "sfb_loglike_params_host")
        return lnL_out, info_out

    def cho_factor(self, Cmat, return_logdet=False):
        """Batched in-place lower Cholesky of device tensor [B,N,N] (row-major; lower triangle read and
        overwritten with L; strict upper untouched).  Returns (Cmat, info[, logdet])."""
        torch = _torch()
        if not (isinstance(Cmat, torch.Tensor) and Cmat.is_cuda and Cmat.dtype == torch.float64
                and Cmat.is_contiguous()):
            raise ValueError("cho_factor needs a contiguous fp64 CUDA tensor [B,N,N]")
        if Cmat.dim() == 2:
            Cmat = Cmat.unsqueeze(0)
        B = Cmat.shape[0]
        info = torch.empty((B,), dtype=torch.int32, device=self.device)
        logdet = torch.empty((B,), dtype=torch.float64, device=self.device)
        self._check(self._lib.sfb_potrf(self._h, B, self._ptr(Cmat), self._ptr(info), self._ptr(logdet),
                                        self._stream()), "sfb_potrf")
        return (Cmat, info, logdet) if return_logdet else (Cmat, info)

    def solve_lower(self, L, r):
        """z = L⁻¹ r for device tensors L[B,N,N] (lower, row-major) and r[B,N]."""
        torch = _torch()
        L = self._dev(L)
        r = self._dev(r)
        if L.dim() == 2:
            L, r = L.unsqueeze(0), r.reshape(1, -1)
        z = torch.empty_like(r)
        self._check(self._lib.sfb_solve_lower(self._h, L.shape[0], self._ptr(L), self._ptr(r), self._ptr(z),
                                              self._stream()), "sfb_solve_lower")
        self._keep = (L, r)
        return z

    # -- profiling ----------------------------------------------------------------------------------
    def profile(self, on: bool):
        self._check(self._lib.sfb_profile_enable(self._h, int(on)), "sfb_profile_enable")

    def profile_read(self):
        """{class: dict(launches, ms, work)} accumulated since the last read (work: bytes for 'build',
        FLOPs otherwise)."""
        buf = (C.c_double * (3 * len(_lib.KERNEL_CLASSES)))()
        self._check(self._lib.sfb_profile_read(self._h, buf, len(buf)), "sfb_profile_read")
        return {name: dict(launches=int(buf[3 * i]), ms=buf[3 * i + 1], work=buf[3 * i + 2])
                for i, name in enumerate(_lib.KERNEL_CLASSES)}


JITTER_DEFAULT = JITTER
