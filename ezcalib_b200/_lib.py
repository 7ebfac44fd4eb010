"""ctypes binding of libezcalib_b200.so (the C ABI declared in include/ezcalib_b200.h).

There is no CPU fallback: if the shared library is missing, or no CUDA device is present,
every compute entry point raises.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libezcalib_b200.so")


class EzcError(RuntimeError):
    pass


class ImageBatchC(C.Structure):
    _fields_ = [("data", C.c_void_p), ("n_images", C.c_int32), ("width", C.c_int32), ("height", C.c_int32),
                ("row_stride", C.c_int64), ("image_stride", C.c_int64)]


class LmProblemC(C.Structure):
    _fields_ = [("model", C.c_int32), ("use_mono_focal", C.c_int32), ("n_blocks", C.c_int32),
                ("obs_per_block", C.c_int32), ("pts_period", C.c_int32), ("cams_per_block", C.c_int32),
                ("img_w", C.c_double), ("img_h", C.c_double), ("pts", C.c_void_p), ("obs", C.c_void_p),
                ("n_blocks_global", C.c_int32), ("collective", C.c_int32)]


class LmOptionsC(C.Structure):
    _fields_ = [("opt_focal", C.c_int32), ("opt_pp", C.c_int32), ("opt_dist", C.c_int32),
                ("max_num_iterations", C.c_int32), ("function_tolerance", C.c_double),
                ("gradient_tolerance", C.c_double), ("parameter_tolerance", C.c_double),
                ("initial_trust_region_radius", C.c_double), ("max_trust_region_radius", C.c_double),
                ("min_trust_region_radius", C.c_double), ("min_relative_decrease", C.c_double),
                ("min_lm_diagonal", C.c_double), ("max_lm_diagonal", C.c_double),
                ("max_num_consecutive_invalid_steps", C.c_int32), ("trace_capacity", C.c_int32),
                ("trace_cost", C.c_void_p), ("trace_radius", C.c_void_p), ("trace_flag", C.c_void_p)]


class LmSummaryC(C.Structure):
    _fields_ = [("iterations", C.c_int32), ("num_successful", C.c_int32), ("termination", C.c_int32),
                ("n_residual_blocks", C.c_int32), ("initial_cost", C.c_double), ("final_cost", C.c_double),
                ("jacobian_evaluations", C.c_int32), ("linear_solves", C.c_int32)]

    def as_dict(self):
        return {k: getattr(self, k) for k, _ in self._fields_}


ALLREDUCE_FN = C.CFUNCTYPE(C.c_int32, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p)

_lib = None


def lib():
    """Load the shared library (built in-tree by ezcalib_b200/csrc/Makefile)."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise EzcError(f"{LIB_PATH} is missing: build it with `make -C ezcalib_b200/csrc` "
                           "(or __graft_entry__.build()); there is no CPU fallback")
        L = C.CDLL(LIB_PATH)
        L.ezc_last_error.restype = C.c_char_p
        L.ezc_kernel_launches.restype = C.c_uint64
        L.ezc_kernel_launches.argtypes = [C.c_void_p]
        L.ezc_create.argtypes = [C.c_int32, C.POINTER(C.c_void_p)]
        L.ezc_destroy.argtypes = [C.c_void_p]
        L.ezc_set_stream.argtypes = [C.c_void_p, C.c_void_p]
        L.ezc_set_allreduce.argtypes = [C.c_void_p, ALLREDUCE_FN, C.c_void_p, C.c_int32, C.c_int32]
        L.ezc_synchronize.argtypes = [C.c_void_p]
        L.ezc_measure_fp64_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.ezc_measure_dmma_peak.argtypes = [C.c_void_p, C.POINTER(C.c_double)]
        L.ezc_images_upload.argtypes = [C.c_void_p, C.POINTER(ImageBatchC), C.POINTER(C.c_void_p)]
        L.ezc_images_upload_async.argtypes = [C.c_void_p, C.POINTER(ImageBatchC), C.POINTER(C.c_void_p)]
        L.ezc_images_wrap_device.argtypes = [C.c_void_p, C.POINTER(ImageBatchC), C.POINTER(C.c_void_p)]
        L.ezc_images_destroy.argtypes = [C.c_void_p]
        L.ezc_corner_subpix.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_double]
        L.ezc_detect_aprilgrid_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32]
        L.ezc_detect_chessboard_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.ezc_find_chessboard_corners_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_void_p, C.c_int32]
        L.ezc_apriltag_debug_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p, C.c_void_p, C.c_void_p]
        L.ezc_apriltag_debug_count.argtypes = [C.c_void_p, C.c_int32]
        L.ezc_apriltag_debug_get.argtypes = [C.c_void_p] * 8
        L.ezc_apriltag_debug_free.argtypes = [C.c_void_p]
        L.ezc_images_download.argtypes = [C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.ezc_render_targets.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32, C.c_int32, C.c_int32,
                                         C.c_int32, C.c_void_p, C.c_void_p, C.c_double, C.c_uint64, C.POINTER(C.c_void_p)]
        L.ezc_lm_create.argtypes = [C.c_void_p, C.POINTER(LmProblemC), C.POINTER(C.c_void_p)]
        L.ezc_lm_create_device.argtypes = [C.c_void_p, C.POINTER(LmProblemC), C.POINTER(C.c_void_p)]
        L.ezc_lm_destroy.argtypes = [C.c_void_p]
        L.ezc_lm_set_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ezc_lm_get_params.argtypes = [C.c_void_p] + [C.c_void_p] * 4
        L.ezc_lm_solve.argtypes = [C.c_void_p, C.POINTER(LmOptionsC), C.POINTER(LmSummaryC)]
        L.ezc_lm_frame_rmse.argtypes = [C.c_void_p, C.c_void_p]
        L.ezc_lm_covariance.argtypes = [C.c_void_p, C.c_void_p]
        L.ezc_lm_mono.argtypes = [C.c_void_p, C.POINTER(LmProblemC)] + [C.c_void_p] * 5 + [C.POINTER(LmSummaryC)]
        L.ezc_lm_residual_jacobian.argtypes = [C.c_void_p, C.c_int32, C.c_int32] + [C.c_void_p] * 5 + [C.c_double, C.c_double, C.c_void_p, C.c_void_p]
        L.ezc_lm_normal_equations.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32] + [C.c_void_p] * 6
        L.ezc_lm_accumulate_only.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_int32]
        L.ezc_lm_default_options.argtypes = [C.POINTER(LmOptionsC)]
        L.ezc_trim_cache.argtypes = [C.c_int32]
        L.ezc_debug_libm.argtypes = [C.c_void_p, C.c_int32, C.c_void_p, C.c_void_p, C.c_int64, C.c_void_p]
        L.ezc_comm_mailbox.argtypes = [C.c_void_p, C.c_int32, C.POINTER(C.c_void_p), C.c_void_p]
        L.ezc_comm_init.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.c_void_p]
        L.ezc_comm_attach.argtypes = [C.c_void_p, C.c_int32, C.c_int32, C.POINTER(C.c_void_p)]
        L.ezc_lm_multicam.argtypes = [C.c_void_p, C.POINTER(LmProblemC)] + [C.c_void_p] * 4 + [C.POINTER(LmSummaryC)]
        L.ezc_p3p_ransac_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_void_p, C.c_int32,
                                           C.c_void_p, C.c_int32, C.c_int32, C.c_int32, C.c_double, C.c_double, C.c_int32] + [C.c_void_p] * 6
        _lib = L
    return _lib


def check(status: int, what: str = ""):
    if status != 0:
        msg = lib().ezc_last_error().decode("utf-8", "replace")
        raise EzcError(f"{what} failed (status {status}): {msg}")


# Every symbol include/ezcalib_b200.h declares (checked by tests/test_abi.py).
EXPORTED_SYMBOLS = [
    "ezc_create", "ezc_destroy", "ezc_last_error", "ezc_set_stream", "ezc_set_allreduce", "ezc_kernel_launches",
    "ezc_synchronize", "ezc_images_upload", "ezc_images_upload_async", "ezc_images_wrap_device", "ezc_images_destroy", "ezc_corner_subpix", "ezc_images_download", "ezc_render_targets", "ezc_detect_aprilgrid_batch", "ezc_detect_chessboard_batch", "ezc_find_chessboard_corners_batch", "ezc_apriltag_debug_run", "ezc_apriltag_debug_count",
    "ezc_apriltag_debug_get", "ezc_apriltag_debug_free", "ezc_measure_fp64_peak", "ezc_measure_dmma_peak", "ezc_num_dist", "ezc_lm_default_options", "ezc_lm_create", "ezc_lm_create_device",
    "ezc_lm_destroy", "ezc_lm_set_params", "ezc_lm_get_params", "ezc_lm_solve", "ezc_lm_frame_rmse",
    "ezc_lm_covariance", "ezc_lm_mono", "ezc_lm_residual_jacobian", "ezc_lm_normal_equations", "ezc_lm_accumulate_only",
    "ezc_trim_cache", "ezc_p3p_ransac_batch", "ezc_lm_multicam", "ezc_comm_mailbox", "ezc_comm_init", "ezc_comm_attach", "ezc_debug_libm",
]
