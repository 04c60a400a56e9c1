"""Host-side mirror of ORB_SLAM3::Optimizer::PoseInertialOptimizationLastKeyFrame / LastFrame (reference
include/Optimizer.h:80-95, src/Optimizer.cc:5899-6284, 6762-7172) over the C ABI: the visual-inertial
pose-only optimisation TrackLocalMap runs on every frame (Tracking.cc:3777,3792), batched over frames
(one CUDA block per frame, every round, the classification and the Hessian hand-over in one launch).
The Frame is passed flattened (GfsPoseInertialProblem, include/gfs_b200.h).  Returns what the reference
leaves behind: the frame's body pose / velocity / biases, mvbOutlier, the average reprojection error, the
inlier count and the 15x15 prior (ConstraintPoseImu::H) for the next frame."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

dp, fp, bp_ = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8)
LAST_KEYFRAME, LAST_FRAME = 0, 1
PRE_STRIDE = 292

_D = C.c_double


class PoseInertialProblem(C.Structure):  # GfsPoseInertialProblem
    _fields_ = [("mode", C.c_int), ("n_obs", C.c_int), ("n_rounds", C.c_int), ("rec_init", C.c_int),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float),
                ("Rcb", _D * 9), ("tcb", _D * 3), ("tbc", _D * 3),
                ("Rwb", _D * 9), ("twb", _D * 3), ("Rcw", _D * 9), ("tcw", _D * 3), ("vel", _D * 3), ("bg", _D * 3), ("ba", _D * 3),
                ("p_Rwb", _D * 9), ("p_twb", _D * 3), ("p_vel", _D * 3), ("p_bg", _D * 3), ("p_ba", _D * 3),
                ("pre", fp), ("rw_Cg", C.c_float * 9), ("rw_Ca", C.c_float * 9),
                ("c_Rwb", _D * 9), ("c_twb", _D * 3), ("c_vwb", _D * 3), ("c_bg", _D * 3), ("c_ba", _D * 3), ("c_H", _D * 225),
                ("Xw", dp), ("uvr", fp), ("inv_sigma2", fp), ("close", bp_)]


class PoseInertialResult(C.Structure):  # GfsPoseInertialResult
    _fields_ = [("n_inliers", C.c_int), ("n_bad", C.c_int), ("n_inliers_last", C.c_int), ("avg_reproj_error", C.c_float),
                ("rounds_done", C.c_int), ("gn_iterations", C.c_int * 4),
                ("Rwb", _D * 9), ("twb", _D * 3), ("vel", _D * 3), ("bg", _D * 3), ("ba", _D * 3), ("H", _D * 225),
                ("outlier", bp_), ("chi2", fp)]


_VEC = ("Rcb", "tcb", "tbc", "Rwb", "twb", "Rcw", "tcw", "vel", "bg", "ba", "p_Rwb", "p_twb", "p_vel", "p_bg", "p_ba",
        "c_Rwb", "c_twb", "c_vwb", "c_bg", "c_ba", "c_H")


def pack_problem(prob, P=None):
    """dict (geoflowslam_b200.synth.pose_inertial_problem layout) -> (struct, keep-alive list)"""
    P = P if P is not None else PoseInertialProblem()
    n = int(prob["n_obs"])
    P.mode, P.n_obs, P.n_rounds, P.rec_init = int(prob["mode"]), n, int(prob["n_rounds"]), int(prob["rec_init"])
    for k in ("fx", "fy", "cx", "cy", "bf"):
        setattr(P, k, float(np.float32(prob[k])))
    for k in _VEC:
        a = np.asarray(prob[k], np.float64).ravel()
        fld = getattr(P, k)
        if len(a) != len(fld):
            raise ValueError("%s: expected %d values" % (k, len(fld)))
        setattr(P, k, type(fld)(*a))
    for k in ("rw_Cg", "rw_Ca"):
        setattr(P, k, (C.c_float * 9)(*np.asarray(prob[k], np.float32).ravel()))
    pre = np.ascontiguousarray(prob["pre"], np.float32).ravel()
    if pre.size != PRE_STRIDE:
        raise ValueError("pre: expected %d floats" % PRE_STRIDE)
    Xw = np.ascontiguousarray(prob["Xw"], np.float64).reshape(-1, 3)
    uvr = np.ascontiguousarray(prob["uvr"], np.float32).reshape(-1, 3)
    is2 = np.ascontiguousarray(prob["inv_sigma2"], np.float32).reshape(-1)
    close = np.ascontiguousarray(prob["close"], np.uint8).reshape(-1)
    if min(len(Xw), len(uvr), len(is2), len(close)) < n:
        raise ValueError("observation arrays shorter than n_obs")
    P.pre = pre.ctypes.data_as(fp)
    P.Xw, P.uvr, P.inv_sigma2, P.close = Xw.ctypes.data_as(dp), uvr.ctypes.data_as(fp), is2.ctypes.data_as(fp), close.ctypes.data_as(bp_)
    return P, [pre, Xw, uvr, is2, close]


def alloc_result(n_obs, R=None):
    R = R if R is not None else PoseInertialResult()
    out = dict(outlier=np.zeros(max(n_obs, 1), np.uint8), chi2=np.zeros(max(n_obs, 1), np.float32))
    R.outlier, R.chi2 = out["outlier"].ctypes.data_as(bp_), out["chi2"].ctypes.data_as(fp)
    return R, out


def unpack_result(R, out, n_obs):
    return dict(n_inliers=R.n_inliers, n_bad=R.n_bad, n_inliers_last=R.n_inliers_last,
                avg_reproj_error=float(R.avg_reproj_error), rounds_done=R.rounds_done, gn_iterations=list(R.gn_iterations),
                Rwb=np.array(R.Rwb).reshape(3, 3), twb=np.array(R.twb), vel=np.array(R.vel), bg=np.array(R.bg), ba=np.array(R.ba),
                H=np.array(R.H).reshape(15, 15), outlier=out["outlier"][:n_obs].astype(bool), chi2=out["chi2"][:n_obs].copy())


class PoseInertialOptimizer:
    """`PoseInertialOptimizer(max_obs, max_batch).PoseInertialOptimizationLastFrame(frame_dict)` /
    `.PoseInertialOptimizationLastKeyFrame(frame_dict)` / `.optimize_batch([...])`"""

    def __init__(self, max_obs=2048, max_batch=1):
        self._L = _lib.lib()
        _lib.require_device()
        self._h = C.c_void_p()
        check(self._L.gfs_pose_inertial_create(int(max_obs), int(max_batch), C.byref(self._h)))
        self.max_obs, self.max_batch = int(max_obs), int(max_batch)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gfs_pose_inertial_destroy(self._h)
            self._h = None

    __del__ = close

    def optimize_batch(self, problems, stream=None):
        n = len(problems)
        Ps = (PoseInertialProblem * n)()
        Rs = (PoseInertialResult * n)()
        keep, outs = [], []
        for i, pr in enumerate(problems):
            keep.append(pack_problem(pr, Ps[i])[1])
            outs.append(alloc_result(int(pr["n_obs"]), Rs[i])[1])
        check(self._L.gfs_pose_inertial_optimize_batch(self._h, stream, Ps, n, Rs))
        return [unpack_result(Rs[i], outs[i], int(problems[i]["n_obs"])) for i in range(n)]

    def PoseInertialOptimizationLastKeyFrame(self, frame, stream=None):
        return self.optimize_batch([dict(frame, mode=LAST_KEYFRAME)], stream)[0]

    def PoseInertialOptimizationLastFrame(self, frame, stream=None):
        return self.optimize_batch([dict(frame, mode=LAST_FRAME)], stream)[0]

    def last_launches(self):
        return int(self._L.gfs_pose_inertial_last_launches(self._h))
