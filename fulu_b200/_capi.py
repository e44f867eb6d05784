"""ctypes binding of include/fulu_gp.h.  Fails loudly when the CUDA library is missing: there is no CPU path."""
from __future__ import annotations

import ctypes
import os

from .build import LIB_PATH

MAX_PARAMS = 6
ABI_VERSION = 5   # FULU_GP_ABI_VERSION of include/fulu_gp.h this binding was written against
c_double_p = ctypes.POINTER(ctypes.c_double)


class Batch(ctypes.Structure):
    _fields_ = [
        ("n_curves", ctypes.c_int64), ("n_obs_total", ctypes.c_int64), ("n_max", ctypes.c_int32), ("n_bands", ctypes.c_int32),
        ("offsets", ctypes.c_void_p), ("t", ctypes.c_void_p), ("flux", ctypes.c_void_p), ("flux_err", ctypes.c_void_p),
        ("band", ctypes.c_void_p), ("loglam", ctypes.c_void_p),
        ("kernel", ctypes.c_int32), ("use_err", ctypes.c_int32), ("alpha0", ctypes.c_double),
        ("order", ctypes.c_void_p), ("n_order", ctypes.c_int64),
    ]


class FitOptions(ctypes.Structure):
    _fields_ = [
        ("n_params", ctypes.c_int32), ("n_starts", ctypes.c_int32), ("m", ctypes.c_int32), ("maxiter", ctypes.c_int32),
        ("maxfun", ctypes.c_int32), ("maxls", ctypes.c_int32), ("window", ctypes.c_int32), ("max_rounds", ctypes.c_int32),
        ("starts_per_curve", ctypes.c_int32), ("profile_counts_only", ctypes.c_int32),
        ("factr", ctypes.c_double), ("pgtol", ctypes.c_double),
        ("lower", ctypes.c_double * MAX_PARAMS), ("upper", ctypes.c_double * MAX_PARAMS),
        ("profile", ctypes.POINTER(ctypes.c_double)),
    ]


EXPORTS = (
    "fulu_gp_abi_version", "fulu_gp_strerror", "fulu_gp_workspace_bytes", "fulu_gp_lml_batch", "fulu_gp_fit_batch",
    "fulu_gp_predict_grid_batch", "fulu_gp_predict_points_batch", "fulu_gp_fp64_probe", "fulu_gp_debug_phase_cycles",
    "fulu_gp_peak_batch", "fulu_gp_ingest_workspace_bytes", "fulu_gp_ingest_table",
)

_lib = None


class FuluGpError(RuntimeError):
    pass


def load():
    """dlopen libfulu_gp.so (built in-tree by fulu_b200.build).  Raises if it is absent."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("FULU_GP_LIB", LIB_PATH)   # override: A/B runs of differently tuned builds of the same library
    if not os.path.exists(path):
        raise FuluGpError(f"fulu_b200: {path} is missing. Build it with `python -m fulu_b200.build` (needs nvcc); "
                          "this package has no CPU fallback.")
    lib = ctypes.CDLL(path)
    for name in EXPORTS:
        if not hasattr(lib, name):
            raise FuluGpError(f"fulu_b200: {path} does not export {name}")
    vp, i32, i64, dbl = ctypes.c_void_p, ctypes.c_int32, ctypes.c_int64, ctypes.c_double
    lib.fulu_gp_abi_version.restype = ctypes.c_int
    lib.fulu_gp_strerror.restype = ctypes.c_char_p
    lib.fulu_gp_strerror.argtypes = [ctypes.c_int]
    lib.fulu_gp_workspace_bytes.argtypes = [ctypes.POINTER(Batch), i32, i32, i32, ctypes.POINTER(ctypes.c_size_t)]
    lib.fulu_gp_lml_batch.argtypes = [ctypes.POINTER(Batch), i32, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.fulu_gp_fit_batch.argtypes = [ctypes.POINTER(Batch), ctypes.POINTER(FitOptions), vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.fulu_gp_predict_grid_batch.argtypes = [ctypes.POINTER(Batch), i32, vp, vp, vp, i32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.fulu_gp_predict_points_batch.argtypes = [ctypes.POINTER(Batch), i32, vp, vp, vp, vp, i32, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.fulu_gp_fp64_probe.argtypes = [i32, i32, vp, ctypes.c_size_t, c_double_p, vp]
    lib.fulu_gp_peak_batch.argtypes = [ctypes.POINTER(Batch), vp, vp, vp, i32, vp, vp, vp, vp, vp]
    lib.fulu_gp_ingest_workspace_bytes.argtypes = [i64, ctypes.POINTER(ctypes.c_size_t)]
    lib.fulu_gp_ingest_table.argtypes = [i64, vp, vp, vp, i32, i64, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    lib.fulu_gp_debug_phase_cycles.argtypes = [ctypes.POINTER(Batch), i32, vp, vp, vp, vp, vp, ctypes.c_size_t, vp]
    for name in EXPORTS:
        if name not in ("fulu_gp_strerror",):
            getattr(lib, name).restype = ctypes.c_int
    if lib.fulu_gp_abi_version() != ABI_VERSION:
        raise FuluGpError("fulu_b200: ABI version mismatch between the Python binding and libfulu_gp.so")
    _lib = lib
    return lib


def check(rc):
    if rc != 0:
        raise FuluGpError(f"libfulu_gp call failed ({rc}): {load().fulu_gp_strerror(rc).decode()}")
