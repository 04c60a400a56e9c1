"""Host-side mirror of RegistrationGICP (reference include/RegistrationGICP.h:15-33,
src/RegistrationGICP.cc:5-20) over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


class GicpSetting(C.Structure):
    _fields_ = [("downsampling_resolution", C.c_double), ("max_correspondence_distance", C.c_double),
                ("rotation_eps", C.c_double), ("translation_eps", C.c_double), ("num_neighbors", C.c_int),
                ("max_iterations", C.c_int), ("num_threads", C.c_int)]


class GicpResult(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("H", C.c_double * 36), ("b", C.c_double * 6), ("error", C.c_double),
                ("iterations", C.c_int), ("num_inliers", C.c_int), ("converged", C.c_int), ("n_target", C.c_int),
                ("n_source", C.c_int), ("inner_evals", C.c_int)]


RESULT_DTYPE = np.dtype([("T", "<f8", (4, 4)), ("H", "<f8", (6, 6)), ("b", "<f8", (6,)), ("error", "<f8"),
                         ("iterations", "<i4"), ("num_inliers", "<i4"), ("converged", "<i4"), ("n_target", "<i4"),
                         ("n_source", "<i4"), ("inner_evals", "<i4")], align=True)
assert RESULT_DTYPE.itemsize == C.sizeof(GicpResult)


def _res_to_dict(r):
    return dict(T=np.array(r["T"]), H=np.array(r["H"]), b=np.array(r["b"]), error=float(r["error"]),
                iterations=int(r["iterations"]), num_inliers=int(r["num_inliers"]), converged=bool(r["converged"]),
                n_target=int(r["n_target"]), n_source=int(r["n_source"]), inner_evals=int(r["inner_evals"]))


class RegistrationGICP:
    """RegistrationGICP().RegisterPointClouds(target_points, source_points, init_T_target_source).
    The setting defaults are the ones the reference hard-codes (voxel 0.02, max dist 0.1, GICP)."""

    def __init__(self, max_points=65536, max_pairs=1, **setting):
        self._L = _lib.lib()
        _lib.require_device()
        s = GicpSetting()
        self._L.gfs_gicp_default_setting(C.byref(s))
        for k, v in setting.items():
            setattr(s, k, v)
        self.setting = s
        h = C.c_void_p()
        check(self._L.gfs_gicp_create(C.byref(s), int(max_points), int(max_pairs), C.byref(h)))
        self._h = h
        self.max_points, self.max_pairs = int(max_points), int(max_pairs)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.gfs_gicp_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def RegisterPointClouds(self, target_points, source_points, init_T_target_source=None, stream=None):
        t = np.ascontiguousarray(target_points, np.float32).reshape(-1, 4)
        s = np.ascontiguousarray(source_points, np.float32).reshape(-1, 4)
        T0 = np.ascontiguousarray(np.eye(4) if init_T_target_source is None else init_T_target_source, np.float64)
        r = np.zeros(1, RESULT_DTYPE)
        check(self._L.gfs_gicp_align(self._h, stream, ptr(t), len(t), ptr(s), len(s), ptr(T0), ptr(r)))
        return _res_to_dict(r[0])

    def align_batch(self, targets, nt, sources, ns, T0, stream=None):
        """targets/sources: (P, stride, 4) float32 host arrays; nt/ns: (P,) int32; T0: (P,4,4)."""
        targets = np.ascontiguousarray(targets, np.float32); sources = np.ascontiguousarray(sources, np.float32)
        P, stride = targets.shape[0], targets.shape[1]
        assert sources.shape[:2] == (P, stride)
        nt = np.ascontiguousarray(nt, np.int32); ns = np.ascontiguousarray(ns, np.int32)
        T0 = np.ascontiguousarray(T0, np.float64).reshape(P, 16)
        r = np.zeros(P, RESULT_DTYPE)
        check(self._L.gfs_gicp_align_batch(self._h, stream, ptr(targets), ptr(nt), ptr(sources), ptr(ns), P, stride,
                                           ptr(T0), ptr(r)))
        return r

    def align_batch_device(self, d_targets, d_nt, d_sources, d_ns, pairs, stride, d_T0, d_out, stream=None):
        check(self._L.gfs_gicp_align_batch_device(self._h, stream, ptr(d_targets), ptr(d_nt), ptr(d_sources), ptr(d_ns),
                                                  int(pairs), int(stride), ptr(d_T0), ptr(d_out)))

    # -- tracking mode (Tracking::PredictStateICP's call pattern): every cloud is preprocessed once and kept on the device
    def track_reset(self):
        check(self._L.gfs_gicp_track_reset(self._h))

    def track_batch_device(self, d_cloud, d_n, seqs, stride, d_T0, d_out, stream=None):
        """New cloud of each of `seqs` sequences (device arrays); registers it against the previous call's.  The first
        call after a reset only stores the clouds."""
        check(self._L.gfs_gicp_track_batch_device(self._h, stream, ptr(d_cloud), ptr(d_n), int(seqs), int(stride), ptr(d_T0),
                                                  ptr(d_out)))

    def track_batch(self, clouds, n, T0=None, stream=None):
        """clouds: (S, stride, 4) float32 host array, n: (S,) int32, T0: (S,4,4) init_T_target_source (previous <- new).
        Returns the (S,) result array, or None for the first call after a reset."""
        clouds = np.ascontiguousarray(clouds, np.float32)
        S, stride = clouds.shape[0], clouds.shape[1]
        n = np.ascontiguousarray(n, np.int32)
        T0 = np.ascontiguousarray(np.tile(np.eye(4), (S, 1, 1)) if T0 is None else T0, np.float64).reshape(S, 16)
        r = np.zeros(S, RESULT_DTYPE)
        first = self._L.gfs_gicp_track_calls(self._h) == 0
        check(self._L.gfs_gicp_track_batch(self._h, stream, ptr(clouds), ptr(n), S, stride, ptr(T0), ptr(r)))
        return None if first else r

    def last_launches(self):
        return self._L.gfs_gicp_last_launches(self._h)

    STAGES = ("group_voxel_pack", "knn_cov", "nn_corr", "linearize", "lm")

    def set_profiling(self, on=True):
        check(self._L.gfs_gicp_set_profiling(self._h, int(on)))

    def profile(self):
        """{stage: (ms, launches)} of the last align / track call (CUDA events on the caller's stream)."""
        ms = np.zeros(8, np.float32); ln = np.zeros(8, np.int32)
        check(self._L.gfs_gicp_get_profile(self._h, ptr(ms), ptr(ln)))
        return {k: (float(ms[i]), int(ln[i])) for i, k in enumerate(self.STAGES)}

    def knn_stats(self, index, stream=None):
        """(occupied grid cells, queries finished by the per-query kernel) of cloud `index` of the last batch."""
        a, b = C.c_int(), C.c_int()
        check(self._L.gfs_gicp_get_knn_stats(self._h, stream, index, C.byref(a), C.byref(b)))
        return a.value, b.value

    def cloud(self, index, stream=None):
        """(downsampled xyz (M,3), covariances (M,6)) of cloud `index` of the last batch."""
        xyz = np.zeros((self.max_points, 3), np.float64); cov = np.zeros((self.max_points, 6), np.float64)
        n = C.c_int()
        check(self._L.gfs_gicp_get_cloud(self._h, stream, index, ptr(xyz), ptr(cov), self.max_points, C.byref(n)))
        return xyz[:n.value].copy(), cov[:n.value].copy()


def predict_state_icp(register, Tcw_last, Tcw_cur, last_points, cur_points, min_points=10, min_inliers=200):
    """Tracking::PredictStateICP (reference src/Tracking.cc:3364-3413), the tracking thread's caller of
    RegistrationGICP::RegisterPointClouds, on plain arrays.  `register(target, source, init_T_target_source)` is
    RegisterPointClouds (the CUDA path: `RegistrationGICP(...).RegisterPointClouds`) and returns its result dict.

    The last frame's cloud is the target, the current frame's the source, the initial guess is Tc1c2 = Tcw_last * Tcw_cur^-1
    (composed in float32, widened to double).  The result is accepted iff `converged && num_inliers > 200`; then the
    current pose becomes T_target_source^-1 * Tcw_last (in float32).  Returns dict(ok, Tcw (4,4) float32 -- unchanged
    if not ok --, delta (4,4) float64 = T_target_source, pos_error = |translation of delta^-1 * Tc1c2|, result)."""
    f32 = np.float32
    last_points = np.asarray(last_points, f32).reshape(-1, 4); cur_points = np.asarray(cur_points, f32).reshape(-1, 4)
    Tl = np.asarray(Tcw_last, f32).reshape(4, 4); Tc = np.asarray(Tcw_cur, f32).reshape(4, 4)
    if len(cur_points) < min_points or len(last_points) < min_points:                        # :3366-3370
        return dict(ok=False, Tcw=Tc.copy(), delta=np.eye(4), pos_error=None, result=None)
    Tc1c2 = (Tl @ np.linalg.inv(Tc)).astype(f32)                                              # :3374-3375
    res = register(last_points, cur_points, Tc1c2.astype(np.float64))                         # :3377-3381
    delta = np.asarray(res["T"], np.float64).reshape(4, 4)
    err = np.linalg.inv(delta) @ Tc1c2.astype(np.float64)                                     # :3387-3391
    pos_error = float(np.linalg.norm(err[:3, 3]))
    if res["converged"] and res["num_inliers"] > min_inliers:                                 # :3392
        Tcw = (np.linalg.inv(delta.astype(f32)) @ Tl).astype(f32)                             # :3393-3395
        return dict(ok=True, Tcw=Tcw, delta=delta, pos_error=pos_error, result=res)
    return dict(ok=False, Tcw=Tc.copy(), delta=delta, pos_error=pos_error, result=res)


def local_ba_icp_edges(register, kf_Tcw, kf_prev, kf_matches_inliers, kf_points, min_inliers=400, max_mean_error=0.01,
                       max_delta_xy=0.1, max_matches=75):
    """The EdgeICP block of Optimizer::LocalInertialBA (reference src/Optimizer.cc:3262-3318), the local-mapping thread's
    caller of RegisterPointClouds, on plain arrays: for every optimizable keyframe i with a predecessor and at most 75
    tracked inliers, register its cloud (source) against the predecessor's (target) starting from
    Tcjci = Tcw[prev] * Tcw[i]^-1 and keep the result as an edge iff it converged, has more than 400 inliers, a mean
    error below 0.01 and moved the initial guess by less than 0.1 m in x-y.

    kf_Tcw (n,4,4); kf_prev[i] = index of the predecessor in the same arrays or -1; kf_points[i] = (m,4) float32 cloud.
    Returns the arrays of the BA problem: dict(n_icp, icp_kf1 (predecessor), icp_kf2 (keyframe), icp_Rt (n_icp,12) =
    [R row-major | t] of T_target_source), plus `results` (one RegisterPointClouds result per attempted keyframe)."""
    T = np.asarray(kf_Tcw, np.float64).reshape(-1, 4, 4)
    kf1, kf2, Rt, results = [], [], [], []
    for i in range(len(T)):
        if kf_matches_inliers[i] > max_matches:                                               # :3266
            continue
        j = int(kf_prev[i])
        if j < 0:                                                                             # :3282
            continue
        Tcjci = T[j] @ np.linalg.inv(T[i])                                                    # :3267-3275 (Sim3 with s = 1)
        res = register(np.asarray(kf_points[j], np.float32), np.asarray(kf_points[i], np.float32), Tcjci)   # :3291-3292
        results.append((i, res))
        rel = np.asarray(res["T"], np.float64).reshape(4, 4)
        delta = rel @ np.linalg.inv(Tcjci)                                                    # :3294
        delta_dist = float(np.float32(np.sqrt(delta[0, 3] * delta[0, 3] + delta[1, 3] * delta[1, 3])))    # :3295-3298
        if (res["converged"] and res["num_inliers"] > min_inliers and res["error"] / res["num_inliers"] < max_mean_error
                and delta_dist < max_delta_xy):                                               # :3299-3300
            kf1.append(j); kf2.append(i)
            Rt.append(np.concatenate([rel[:3, :3].ravel(), rel[:3, 3]]))
    return dict(n_icp=len(kf1), icp_kf1=np.asarray(kf1, np.int32), icp_kf2=np.asarray(kf2, np.int32),
                icp_Rt=np.asarray(Rt, np.float64).reshape(-1, 12), results=results)
