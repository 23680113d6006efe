"""ctypes binding of the C ABI declared in include/mgicp.h (libmgicp.so, built in-tree by build.py).

There is no fallback: if the CUDA library is missing or cannot be loaded, importing the engine fails
loudly.  Nothing in this package touches oracle/.
"""
from __future__ import annotations

import ctypes as C
import os

from . import build as _build

STATUS = {0: "OK", 1: "INVALID", 2: "CUDA", 3: "NOMEM", 4: "RANGE", 5: "OVERFLOW", 6: "STATE"}
F32, F64 = 0, 1
LOSS = {"l2": 0, "l1": 1, "huber": 2, "cauchy": 3, "gm": 4, "tukey": 5}
(STAGE_DOWNSAMPLED, STAGE_GRID_POINTS, STAGE_SOR_AVG, STAGE_SOR_KEEP, STAGE_POINTS, STAGE_NORMALS, STAGE_KNN_SOR,
 STAGE_KNN_NORMAL, STAGE_BOUNDS, STAGE_ICP_POINTS, STAGE_ICP_NORMALS) = range(11)

EXPORTS = ["mgicp_default_opts", "mgicp_create", "mgicp_destroy", "mgicp_last_error", "mgicp_version",
           "mgicp_kernel_launches", "mgicp_cloud_bounds", "mgicp_preprocess", "mgicp_register_batch", "mgicp_run_batch",
           "mgicp_evaluate_batch", "mgicp_evaluate_clouds", "mgicp_fpfh_clouds", "mgicp_fgr_pairs", "mgicp_get_stage", "mgicp_check",
           "mgicp_job_errors", "mgicp_set_timing", "mgicp_get_timing", "mgicp_get_correspondences", "mgicp_auto_icp_cell_factor"]


class Opts(C.Structure):
    _fields_ = [("sor_k", C.c_int32), ("sor_std", C.c_double), ("normal_k", C.c_int32), ("epsilon", C.c_double),
                ("loss", C.c_int32), ("loss_k", C.c_double), ("rel_fitness", C.c_double), ("rel_rmse", C.c_double),
                ("cell_factor", C.c_double), ("icp_cell_factor", C.c_double), ("ctas_per_pair", C.c_int32), ("debug", C.c_int32)]


class FgrOpts(C.Structure):
    """mgicp_fgr_opts: Open3D's FastGlobalRegistrationOption"""
    _fields_ = [("division_factor", C.c_double), ("use_absolute_scale", C.c_int32), ("decrease_mu", C.c_int32),
                ("maximum_correspondence_distance", C.c_double), ("iteration_number", C.c_int32), ("tuple_scale", C.c_double),
                ("maximum_tuple_count", C.c_int32)]


_lib = None


def lib_path() -> str:
    return _build.LIB


def load():
    """Load libmgicp.so (building it first if the sources are newer).  Raises if that is impossible."""
    global _lib
    if _lib is not None:
        return _lib
    path = os.environ.get("MGICP_LIB") or _build.LIB      # MGICP_LIB: A/B experiments against another build of the same ABI
    if path == _build.LIB and _build.needs_build():
        try:
            _build.build()
        except Exception as e:  # no nvcc on this machine: a prebuilt library must already be there
            if not os.path.exists(path):
                raise RuntimeError(f"libmgicp.so is missing and could not be built: {e}") from e
    L = C.CDLL(path)
    vp, i32, i64, dbl = C.c_void_p, C.c_int32, C.c_int64, C.c_double
    P = C.POINTER
    L.mgicp_default_opts.argtypes = [P(Opts)]
    L.mgicp_default_opts.restype = None
    L.mgicp_create.argtypes = [C.c_int, P(vp)]
    L.mgicp_destroy.argtypes = [vp]
    L.mgicp_last_error.argtypes = [vp]
    L.mgicp_last_error.restype = C.c_char_p
    L.mgicp_version.restype = C.c_char_p
    L.mgicp_kernel_launches.argtypes = [vp]
    L.mgicp_kernel_launches.restype = i64
    L.mgicp_cloud_bounds.argtypes = [vp, vp, i32, vp, P(i64), i32, vp]
    L.mgicp_preprocess.argtypes = [vp, vp, i32, vp, P(i64), i32, i32, P(dbl), P(Opts)]
    L.mgicp_register_batch.argtypes = [vp, vp, i32, P(i32), P(i32), P(dbl), P(i32), P(Opts), vp, vp, vp, vp, vp, vp, vp]
    L.mgicp_run_batch.argtypes = [vp, vp, i32, vp, P(i64), i32, i32, P(dbl), i32, P(i32), P(i32), P(dbl), P(i32), P(Opts),
                                  vp, vp, vp, vp, vp, vp, vp]
    L.mgicp_evaluate_batch.argtypes = [vp, vp, i32, i32, P(i32), P(i32), P(dbl), P(Opts), vp, vp]
    L.mgicp_evaluate_clouds.argtypes = [vp, vp, i32, vp, P(i64), i32, i32, P(i32), P(i32), P(dbl), P(dbl), vp, vp]
    L.mgicp_fpfh_clouds.argtypes = [vp, vp, i32, vp, P(i64), i32, dbl, i32, dbl, i32, vp, vp]
    L.mgicp_fgr_pairs.argtypes = [vp, vp, i32, vp, P(i64), i32, vp, i32, P(i32), P(i32), P(FgrOpts), P(i32), P(C.c_uint64), vp, vp]
    L.mgicp_get_stage.argtypes = [vp, i32, i32, i32, vp, i64, P(i64)]
    L.mgicp_check.argtypes = [vp]
    L.mgicp_job_errors.argtypes = [vp, vp, vp]
    L.mgicp_set_timing.argtypes = [vp, i32]
    L.mgicp_get_timing.argtypes = [vp, P(dbl)]
    L.mgicp_get_correspondences.argtypes = [vp, i32, vp, i64, P(i64)]
    L.mgicp_auto_icp_cell_factor.argtypes = [i32, P(dbl), i32, P(dbl)]
    L.mgicp_auto_icp_cell_factor.restype = dbl
    for name in ("mgicp_create", "mgicp_destroy", "mgicp_cloud_bounds", "mgicp_preprocess", "mgicp_register_batch",
                 "mgicp_run_batch", "mgicp_evaluate_batch", "mgicp_evaluate_clouds", "mgicp_fpfh_clouds", "mgicp_fgr_pairs", "mgicp_get_stage", "mgicp_check",
                 "mgicp_job_errors", "mgicp_set_timing", "mgicp_get_timing", "mgicp_get_correspondences"):
        getattr(L, name).restype = C.c_int
    _lib = L
    return L
