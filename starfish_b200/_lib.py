"""ctypes binding of libsfb200.so — the thin layer between Python and the sm_100a kernels.

The signatures here are a 1:1 transcription of include/sfb200.h.  There is NO fallback: if the library
is missing the import fails loudly (build it with ``python -m starfish_b200.build``).
"""
import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("SFB200_LIB", os.path.join(_HERE, "libsfb200.so"))  # override: dev experiments only

EXPORTS = [
    "sfb_abi_version", "sfb_create", "sfb_destroy", "sfb_set_static", "sfb_set_static_host",
    "sfb_build_cov", "sfb_potrf", "sfb_solve_lower", "sfb_loglike", "sfb_loglike_host", "sfb_sync",
    "sfb_profile_enable", "sfb_profile_read", "sfb_workspace_walkers", "sfb_padded_n",
    "sfb_launch_count", "sfb_last_error",
    "sfb_set_model_host", "sfb_upstream", "sfb_loglike_params", "sfb_loglike_params_host",
    "sfb_host_rfft", "sfb_host_spline_inverse_band", "sfb_host_cholesky_lower", "sfb_spline_halfwidth",
    "sfb_set_solver", "sfb_get_solver", "sfb_band_classes",
    "sfb_comm_unique_id", "sfb_comm_init", "sfb_allgather_lnL", "sfb_comm_destroy",
    "sfb_set_shared_factor", "sfb_shared_factor_calls", "sfb_i8_mma_counts",
]

SOLVER_DENSE, SOLVER_STRUCTURED, SOLVER_DENSE_I8 = 0, 1, 2

ABI_VERSION = 5

# sfb_model_flags of include/sfb200.h
MODEL_VSINI, MODEL_VZ, MODEL_LOG_SCALE, MODEL_NORM, MODEL_PAPER_TERM, MODEL_AV = 1, 2, 4, 8, 16, 32

KERNEL_CLASSES = ("build", "potrf_diag", "trsm", "syrk", "upstream", "band_build", "band_chol", "oz_slice", "fwd_rows")

_p = C.c_void_p
_i = C.c_int
_d = C.c_double


class SfbError(RuntimeError):
    pass


def load():
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} not found: the CUDA extension is required (no CPU fallback). "
            "Build it with `python -m starfish_b200.build`."
        )
    lib = C.CDLL(LIB_PATH)
    lib.sfb_abi_version.restype = _i
    lib.sfb_create.argtypes = [_i, _i, _i, _i, _i, _i, C.POINTER(_p)]
    lib.sfb_destroy.argtypes = [_p]
    lib.sfb_set_static.argtypes = [_p, _p, _p, _p, _p]
    lib.sfb_set_static_host.argtypes = [_p, _p, _p, _p]
    lib.sfb_build_cov.argtypes = [_p, _i, _p, _p, _p, _p, _p, _i, _d, _p, _p]
    lib.sfb_potrf.argtypes = [_p, _i, _p, _p, _p, _p]
    lib.sfb_solve_lower.argtypes = [_p, _i, _p, _p, _p, _p]
    lib.sfb_loglike.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p, _p]
    lib.sfb_loglike_host.argtypes = [_p, _i, _p, _p, _p, _p, _p, _p, _i, _p, _p, _p]
    lib.sfb_set_model_host.argtypes = [_p, _i, _p, _p, _i, _i, _p, _p, _p, _p, _p, _i, _i]
    lib.sfb_upstream.argtypes = [_p, _i, _p, _i, _p, _p, _p, _p, _p, _p, _p, _p]
    lib.sfb_loglike_params.argtypes = [_p, _i, _p, _i, _p, _p, _p, _i, _p, _p, _p, _p, _p]
    lib.sfb_loglike_params_host.argtypes = [_p, _i, _p, _i, _p, _p, _p, _i, _p, _p, _p, _p]
    lib.sfb_host_rfft.argtypes = [_i, _p, _p]
    lib.sfb_host_spline_inverse_band.argtypes = [_i, _p, _i, _p]
    lib.sfb_host_cholesky_lower.argtypes = [_i, _p]
    lib.sfb_spline_halfwidth.argtypes = []
    lib.sfb_set_solver.argtypes = [_p, _i]
    lib.sfb_get_solver.argtypes = [_p]
    lib.sfb_band_classes.argtypes = [_p, _p, _p, _i]
    lib.sfb_comm_unique_id.argtypes = [_p]
    lib.sfb_comm_init.argtypes = [_p, _i, _i, _p]
    lib.sfb_allgather_lnL.argtypes = [_p, _p, _i, _p, _p]
    lib.sfb_comm_destroy.argtypes = [_p]
    lib.sfb_set_shared_factor.argtypes = [_p, _i]
    lib.sfb_shared_factor_calls.argtypes = [_p]
    lib.sfb_shared_factor_calls.restype = C.c_longlong
    lib.sfb_i8_mma_counts.argtypes = [_p, _p, _p]
    lib.sfb_sync.argtypes = [_p]
    lib.sfb_profile_enable.argtypes = [_p, _i]
    lib.sfb_profile_read.argtypes = [_p, _p, _i]
    lib.sfb_workspace_walkers.argtypes = [_p]
    lib.sfb_padded_n.argtypes = [_p]
    lib.sfb_launch_count.argtypes = [_p]
    lib.sfb_launch_count.restype = C.c_longlong
    lib.sfb_last_error.argtypes = [_p]
    lib.sfb_last_error.restype = C.c_char_p
    for name in EXPORTS:
        fn = getattr(lib, name)
        if fn.restype is None:
            fn.restype = _i
    return lib


_LIB = None


def lib():
    global _LIB
    if _LIB is None:
        _LIB = load()
    return _LIB
