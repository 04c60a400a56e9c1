"""ctypes binding of libgfs_b200.so (the C ABI declared in include/gfs_b200.h).

There is no CPU fallback: if the library cannot be built/loaded, or no CUDA device is present
when a compute entry point is called, the call raises.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libgfs_b200.so")
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])

GFS_OK, GFS_ERR_INVALID, GFS_ERR_CUDA, GFS_ERR_CAPACITY, GFS_ERR_EMPTY, GFS_ERR_NODEVICE = 0, -1, -2, -3, -4, -5


class GfsError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__("libgfs_b200 error %d: %s" % (code, msg))
        self.code = code


def _sig(L, name, argtypes, restype=C.c_int):
    f = getattr(L, name)
    f.argtypes = argtypes
    f.restype = restype


vp, ci, cf, sz = C.c_void_p, C.c_int, C.c_float, C.c_size_t

# every symbol include/gfs_b200.h declares (tests/test_cabi.py checks the library exports them all)
SIGNATURES = {
    "gfs_last_error": ([], C.c_char_p),
    "gfs_device_check": ([], ci),
    "gfs_version": ([], C.c_char_p),
    "gfs_orb_create": ([ci, cf, ci, ci, ci, ci, ci, ci, C.POINTER(vp)], ci),
    "gfs_orb_destroy": ([vp], ci),
    "gfs_orb_max_keypoints": ([vp], ci),
    "gfs_orb_tables": ([vp, vp, vp], ci),
    "gfs_orb_level_size": ([vp, ci, ci, ci, vp, vp], ci),
    "gfs_orb_extract_batch_device": ([vp, vp, vp, ci, ci, ci, ci, sz, ci, ci, vp, vp, vp, vp], ci),
    "gfs_orb_extract_batch": ([vp, vp, vp, ci, ci, ci, ci, sz, ci, ci, vp, vp, vp, vp], ci),
    "gfs_orb_extract": ([vp, vp, vp, ci, ci, ci, ci, ci, vp, vp, vp, vp], ci),
    "gfs_orb_get_level": ([vp, vp, ci, ci, ci, vp], ci),
    "gfs_orb_get_candidates": ([vp, vp, ci, ci, vp, ci, vp], ci),
    "gfs_orb_launches_per_call": ([vp, ci, ci], ci),
    "gfs_orb_set_profiling": ([vp, ci], ci),
    "gfs_orb_get_profile": ([vp, vp], ci),
    "gfs_search_by_projection": ([vp, ci, cf, ci, vp, ci, vp, vp, vp, vp, ci, cf, cf, cf, cf, vp, vp], ci),
    "gfs_search_by_projection_batch_device": ([vp, ci, cf, ci, vp, vp, ci, vp, vp, vp, vp, vp, ci, ci, cf, cf, cf, cf, vp, vp, vp], ci),
    "gfs_depth_to_cloud": ([vp, vp, ci, ci, ci, cf, cf, cf, cf, vp, ci, vp], ci),
    "gfs_depth_to_cloud_batch_device": ([vp, vp, ci, ci, ci, ci, sz, ci, cf, cf, cf, cf, vp, ci, vp], ci),
    "gfs_frontend_create": ([ci, cf, ci, ci, ci, ci, ci, ci, C.POINTER(vp)], ci),
    "gfs_frontend_destroy": ([vp], ci),
    "gfs_frontend_max_keypoints": ([vp], ci),
    "gfs_frontend_orb": ([vp], vp),
    "gfs_frontend_run_device": ([vp, vp, vp, ci, ci, ci, ci, sz, vp, vp, vp, vp, vp, vp, vp, vp], ci),
    "gfs_frontend_run": ([vp, vp, vp, ci, ci, ci, ci, sz, vp, vp, vp, vp, vp, vp, vp, vp], ci),
    "gfs_frontend_set_profiling": ([vp, ci], ci),
    "gfs_frontend_get_profile": ([vp, vp], ci),
    "gfs_frontend_launches_per_call": ([vp, ci], ci),
    "gfs_gicp_default_setting": ([vp], None),
    "gfs_gicp_create": ([vp, ci, ci, C.POINTER(vp)], ci),
    "gfs_gicp_destroy": ([vp], ci),
    "gfs_gicp_align": ([vp, vp, vp, ci, vp, ci, vp, vp], ci),
    "gfs_gicp_align_batch": ([vp, vp, vp, vp, vp, vp, ci, ci, vp, vp], ci),
    "gfs_gicp_align_batch_device": ([vp, vp, vp, vp, vp, vp, ci, ci, vp, vp], ci),
    "gfs_gicp_track_reset": ([vp], ci),
    "gfs_gicp_track_calls": ([vp], ci),
    "gfs_gicp_track_batch_device": ([vp, vp, vp, vp, ci, ci, vp, vp], ci),
    "gfs_gicp_track_batch": ([vp, vp, vp, vp, ci, ci, vp, vp], ci),
    "gfs_gicp_get_cloud": ([vp, vp, ci, vp, vp, ci, vp], ci),
    "gfs_gicp_last_launches": ([vp], ci),
    "gfs_gicp_set_profiling": ([vp, ci], ci),
    "gfs_gicp_get_profile": ([vp, vp, vp], ci),
    "gfs_gicp_get_knn_stats": ([vp, vp, ci, vp, vp], ci),
    "gfs_ba_create": ([ci, ci, ci, ci, ci, C.POINTER(vp)], ci),
    "gfs_ba_destroy": ([vp], ci),
    "gfs_ba_solve_batch": ([vp, vp, vp, vp, ci], ci),
    "gfs_ba_solve": ([vp, vp, vp, vp], ci),
    "gfs_ba_upload": ([vp, vp, vp, ci], ci),
    "gfs_ba_solve_uploaded": ([vp, vp], ci),
    "gfs_ba_download": ([vp, vp, vp, ci], ci),
    "gfs_ba_last_launches": ([vp], ci),
    "gfs_ba_set_partition": ([vp, ci, ci, vp, vp], ci),
    "gfs_nccl_unique_id": ([vp], ci),
    "gfs_ba_set_partition_nccl": ([vp, ci, ci, vp], ci),
    "gfs_ba_last_nccl_calls": ([vp], ci),
    "gfs_pose_create": ([ci, ci, C.POINTER(vp)], ci),
    "gfs_pose_destroy": ([vp], ci),
    "gfs_pose_optimize_batch": ([vp, vp, vp, ci, vp], ci),
    "gfs_pose_optimize": ([vp, vp, vp, vp], ci),
    "gfs_pose_last_launches": ([vp], ci),
    "gfs_pose_inertial_create": ([ci, ci, C.POINTER(vp)], ci),
    "gfs_pose_inertial_destroy": ([vp], ci),
    "gfs_pose_inertial_optimize_batch": ([vp, vp, vp, ci, vp], ci),
    "gfs_pose_inertial_optimize": ([vp, vp, vp, vp], ci),
    "gfs_pose_inertial_last_launches": ([vp], ci),
    "gfs_klt_create": ([ci, ci, ci, ci, ci, C.POINTER(vp)], ci),
    "gfs_klt_destroy": ([vp], ci),
    "gfs_klt_pyramid_bytes": ([vp, ci, ci], C.c_size_t),
    "gfs_klt_pyramid_layout": ([vp, ci, ci, vp, vp, vp, vp], ci),
    "gfs_klt_build_pyramid_batch_device": ([vp, vp, vp, ci, ci, ci, ci, C.c_size_t, vp], ci),
    "gfs_klt_fb_track_batch_device": ([vp, vp, vp, vp, ci, ci, ci, vp, vp, vp, ci, ci, ci, C.c_float, C.c_float, vp], ci),
    "gfs_klt_calc_batch_device": ([vp, vp, vp, vp, ci, ci, ci, vp, vp, vp, ci, ci, ci, ci, C.c_float, ci, vp, vp], ci),
    "gfs_klt_fb_track": ([vp, vp, vp, vp, ci, ci, ci, vp, vp, ci, ci, ci, C.c_float, C.c_float, vp], ci),
    "gfs_klt_last_launches": ([vp], ci),
    "gfs_clahe_apply_batch_device": ([vp, vp, ci, ci, ci, ci, C.c_size_t, C.c_double, ci, ci, vp, ci, C.c_size_t], ci),
    "gfs_clahe_apply": ([vp, vp, ci, ci, ci, C.c_double, ci, ci, vp], ci),
    "gfs_imu_preintegrate_batch": ([vp, vp, vp, vp, ci, C.c_float, C.c_float, C.c_float, C.c_float, vp], ci),
    "gfs_imu_preintegrate_batch_device": ([vp, vp, vp, vp, ci, C.c_float, C.c_float, C.c_float, C.c_float, vp], ci),
    "gfs_match_bf_hamming_batch_device": ([vp, vp, vp, vp, vp, ci, ci, vp, vp], ci),
    "gfs_match_bf_hamming": ([vp, vp, ci, vp, ci, vp, vp], ci),
    "gfs_gms_filter_batch_device": ([vp, vp, vp, vp, vp, vp, ci, ci, ci, ci, ci, ci, vp, vp], ci),
    "gfs_gms_filter": ([vp, vp, ci, ci, ci, vp, ci, ci, ci, vp, ci, vp, vp], ci),
}
DEBUG_SIGNATURES = {
    "gfs_debug_gcc_sort": ([vp, vp, vp, ci, vp], ci),
}


def lib():
    """Load (building if needed) the native library.  Raises if it cannot be had."""
    global _LIB
    if _LIB is None:
        if not os.path.exists(_SO):
            from . import build
            build.build_native()
        L = C.CDLL(_SO)
        for table in (SIGNATURES, DEBUG_SIGNATURES):
            for name, (args, res) in table.items():
                if hasattr(L, name):
                    _sig(L, name, args, res)
        _LIB = L
    return _LIB


def check(rc):
    if rc != 0:
        raise GfsError(rc, lib().gfs_last_error().decode("utf-8", "replace"))


def require_device():
    check(lib().gfs_device_check())


def ptr(a):
    """void* of a numpy array, torch tensor, int address or None."""
    if a is None:
        return None
    if isinstance(a, int):
        return C.c_void_p(a)
    if isinstance(a, np.ndarray):
        return a.ctypes.data_as(C.c_void_p)
    if hasattr(a, "data_ptr"):
        return C.c_void_p(a.data_ptr())
    raise TypeError(type(a))
