"""Host-side mirror of ORB_SLAM3::Optimizer::PoseOptimization (reference include/Optimizer.h:62-65,
src/Optimizer.cc:763-1099) over the C ABI: the motion-only bundle adjustment of the tracking thread,
batched over frames (one CUDA block per frame, the whole 4-round optimisation in one launch).  The
Frame is passed flattened (GfsPoseProblem, include/gfs_b200.h): its pose, pinhole parameters and, for
every feature with a MapPoint, the world point, the undistorted keypoint / right coordinate and the
level's inverse sigma^2.  Returns what the reference leaves behind: mvbOutlier, the average
reprojection error and the inlier count."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

dp, fp, bp_ = C.POINTER(C.c_double), C.POINTER(C.c_float), C.POINTER(C.c_uint8)


class PoseProblem(C.Structure):  # GfsPoseProblem
    _fields_ = [("n_obs", C.c_int), ("q_wxyz", C.c_float * 4), ("t", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("Xw", dp), ("uvr", fp), ("inv_sigma2", fp)]


class PoseResult(C.Structure):  # GfsPoseResult
    _fields_ = [("n_inliers", C.c_int), ("n_bad", C.c_int), ("n_good", C.c_int), ("avg_reproj_error", C.c_float),
                ("rounds_done", C.c_int), ("lm_iterations", C.c_int * 4), ("q_wxyz", C.c_double * 4), ("t", C.c_double * 3),
                ("outlier", bp_), ("chi2", fp)]


def pack_problem(prob, P=None):
    """dict (geoflowslam_b200.synth.pose_problem layout) -> (struct, keep-alive list)"""
    P = P if P is not None else PoseProblem()
    n = int(prob["n_obs"])
    P.n_obs = n
    P.q_wxyz = (C.c_float * 4)(*np.asarray(prob["q_wxyz"], np.float32))
    P.t = (C.c_float * 3)(*np.asarray(prob["t"], np.float32))
    for k in ("fx", "fy", "cx", "cy", "bf"):
        setattr(P, k, float(np.float32(prob[k])))
    Xw = np.ascontiguousarray(prob["Xw"], np.float64).reshape(-1, 3)
    uvr = np.ascontiguousarray(prob["uvr"], np.float32).reshape(-1, 3)
    is2 = np.ascontiguousarray(prob["inv_sigma2"], np.float32).reshape(-1)
    if len(Xw) < n or len(uvr) < n or len(is2) < n:
        raise ValueError("observation arrays shorter than n_obs")
    P.Xw, P.uvr, P.inv_sigma2 = Xw.ctypes.data_as(dp), uvr.ctypes.data_as(fp), is2.ctypes.data_as(fp)
    return P, [Xw, uvr, is2]


def alloc_result(n_obs, R=None):
    R = R if R is not None else PoseResult()
    out = dict(outlier=np.zeros(max(n_obs, 1), np.uint8), chi2=np.zeros(max(n_obs, 1), np.float32))
    R.outlier, R.chi2 = out["outlier"].ctypes.data_as(bp_), out["chi2"].ctypes.data_as(fp)
    return R, out


def unpack_result(R, out, n_obs):
    return dict(n_inliers=R.n_inliers, n_bad=R.n_bad, n_good=R.n_good, avg_reproj_error=float(R.avg_reproj_error),
                rounds_done=R.rounds_done, lm_iterations=list(R.lm_iterations), q_wxyz=np.array(R.q_wxyz), t=np.array(R.t),
                outlier=out["outlier"][:n_obs].astype(bool), chi2=out["chi2"][:n_obs].copy())


class PoseOptimizer:
    """`PoseOptimizer(max_obs, max_batch).PoseOptimization(frame_dict)` / `.optimize_batch([...])`"""

    def __init__(self, max_obs=2048, max_batch=1):
        self._L = _lib.lib()
        _lib.require_device()
        self._h = C.c_void_p()
        check(self._L.gfs_pose_create(int(max_obs), int(max_batch), C.byref(self._h)))
        self.max_obs, self.max_batch = int(max_obs), int(max_batch)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gfs_pose_destroy(self._h)
            self._h = None

    __del__ = close

    def optimize_batch(self, problems, stream=None):
        n = len(problems)
        Ps = (PoseProblem * n)()
        Rs = (PoseResult * n)()
        keep, outs = [], []
        for i, pr in enumerate(problems):
            keep.append(pack_problem(pr, Ps[i])[1])
            outs.append(alloc_result(int(pr["n_obs"]), Rs[i])[1])
        check(self._L.gfs_pose_optimize_batch(self._h, stream, Ps, n, Rs))
        return [unpack_result(Rs[i], outs[i], int(problems[i]["n_obs"])) for i in range(n)]

    def PoseOptimization(self, frame, stream=None):
        """-> dict; dict['n_inliers'] is the reference's return value"""
        return self.optimize_batch([frame], stream)[0]

    def last_launches(self):
        return int(self._L.gfs_pose_last_launches(self._h))
