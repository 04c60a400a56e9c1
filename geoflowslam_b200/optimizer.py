"""Host-side mirror of the numerical core of ORB_SLAM3::Optimizer::LocalInertialBA (reference
include/Optimizer.h:180-183, src/Optimizer.cc:3056-3702) over the C ABI.  The reference's KeyFrame /
MapPoint pointer graph is passed flattened (GfsBaProblem, include/gfs_b200.h); window selection,
outlier erasure and the write-back into the map stay with the caller."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check

dp, ip, fp, bp_ = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8)


class BaProblem(C.Structure):  # GfsBaProblem
    _fields_ = [("n_opt_kf", C.c_int), ("n_fixed_kf", C.c_int), ("n_points", C.c_int), ("n_obs", C.c_int),
                ("n_inertial", C.c_int), ("iterations", C.c_int), ("b_large", C.c_int), ("lambda_init", C.c_double),
                ("Rcb", C.c_double * 9), ("tcb", C.c_double * 3), ("Rbc", C.c_double * 9), ("tbc", C.c_double * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_double),
                ("kf_Rwb", dp), ("kf_twb", dp), ("kf_Rcw", dp), ("kf_tcw", dp), ("kf_vel", dp), ("kf_bg", dp),
                ("kf_ba", dp), ("kf_has_imu", bp_), ("pt_xyz", dp), ("pt_close", bp_), ("obs_kf", ip), ("obs_pt", ip),
                ("obs_uvr", dp), ("obs_inv_sigma2", fp), ("in_kf1", ip), ("in_kf2", ip), ("in_pre", fp),
                ("in_downweight", bp_), ("n_icp", C.c_int), ("icp_kf1", ip), ("icp_kf2", ip), ("icp_Rt", dp),
                ("vertex_se3", C.c_int)]


class BaResult(C.Structure):  # GfsBaResult
    _fields_ = [("kf_Rwb", dp), ("kf_twb", dp), ("kf_Rcw", dp), ("kf_tcw", dp), ("kf_vel", dp), ("kf_bg", dp),
                ("kf_ba", dp), ("pt_xyz", dp), ("obs_chi2", dp), ("obs_depth_positive", bp_), ("obs_outlier", bp_),
                ("err", C.c_float), ("err_end", C.c_float), ("failed", C.c_int), ("iterations_done", C.c_int),
                ("lm_trials", C.c_int), ("lambda_final", C.c_double)]


_ARRAYS = [("kf_Rwb", np.float64), ("kf_twb", np.float64), ("kf_Rcw", np.float64), ("kf_tcw", np.float64),
           ("kf_vel", np.float64), ("kf_bg", np.float64), ("kf_ba", np.float64), ("kf_has_imu", np.uint8),
           ("pt_xyz", np.float64), ("pt_close", np.uint8), ("obs_kf", np.int32), ("obs_pt", np.int32),
           ("obs_uvr", np.float64), ("obs_inv_sigma2", np.float32), ("in_kf1", np.int32), ("in_kf2", np.int32),
           ("in_pre", np.float32), ("in_downweight", np.uint8)]
_BA_OPTIONAL = [("icp_kf1", np.int32), ("icp_kf2", np.int32), ("icp_Rt", np.float64)]


def pack_problem(prob, P=None):
    P = P if P is not None else BaProblem()
    keep = []
    for k in ("n_opt_kf", "n_fixed_kf", "n_points", "n_obs", "n_inertial", "iterations", "b_large"):
        setattr(P, k, int(prob[k]))
    P.lambda_init = float(prob["lambda_init"])
    for k, n in (("Rcb", 9), ("tcb", 3), ("Rbc", 9), ("tbc", 3)):
        setattr(P, k, (C.c_double * n)(*np.asarray(prob[k], np.float64).ravel()))
    for k in ("fx", "fy", "cx", "cy"):
        setattr(P, k, float(prob[k]))
    P.bf = float(prob["bf"])
    fields = dict(BaProblem._fields_)
    P.n_icp = int(prob.get("n_icp", 0))
    P.vertex_se3 = int(prob.get("vertex_se3", 0))
    for k, dt in _BA_OPTIONAL:
        a = np.ascontiguousarray(prob.get(k, np.zeros(0)), dt)
        keep.append(a)
        setattr(P, k, a.ctypes.data_as(fields[k]))
    for k, dt in _ARRAYS:
        a = np.ascontiguousarray(prob[k], dt)
        keep.append(a)
        setattr(P, k, a.ctypes.data_as(fields[k]))
    return P, keep


def alloc_result(prob, R=None):
    R = R if R is not None else BaResult()
    nk = prob["n_opt_kf"] + prob["n_fixed_kf"]
    no = max(prob["n_obs"], 1)
    out = dict(kf_Rwb=np.zeros((nk, 9)), kf_twb=np.zeros((nk, 3)), kf_Rcw=np.zeros((nk, 9)), kf_tcw=np.zeros((nk, 3)),
               kf_vel=np.zeros((nk, 3)), kf_bg=np.zeros((nk, 3)), kf_ba=np.zeros((nk, 3)),
               pt_xyz=np.zeros((max(prob["n_points"], 1), 3)), obs_chi2=np.zeros(no),
               obs_depth_positive=np.zeros(no, np.uint8), obs_outlier=np.zeros(no, np.uint8))
    fields = dict(BaResult._fields_)
    for k, a in out.items():
        setattr(R, k, a.ctypes.data_as(fields[k]))
    return R, out


def unpack_result(R, out, prob):
    res = {}
    for k, v in out.items():
        res[k] = v[:prob["n_obs"]] if k.startswith("obs_") else (v[:prob["n_points"]] if k == "pt_xyz" else v)
    res.update(err=R.err, err_end=R.err_end, failed=bool(R.failed), iterations_done=R.iterations_done,
               lm_trials=R.lm_trials, lambda_final=R.lambda_final)
    return res


class Optimizer:
    """Optimizer.LocalInertialBA(problem) / Optimizer.LocalBundleAdjustment(problem) -> result dict; batch variant for
    independent problems (inertial and non-inertial problems may share a batch)."""

    def __init__(self, max_kf=21, max_points=3000, max_obs=16384, max_inertial=20, max_batch=1):
        self._L = _lib.lib()
        _lib.require_device()
        h = C.c_void_p()
        check(self._L.gfs_ba_create(int(max_kf), int(max_points), int(max_obs), int(max_inertial), int(max_batch),
                                    C.byref(h)))
        self._h = h
        self.max_batch = int(max_batch)
        self._uploaded = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.gfs_ba_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def LocalInertialBA(self, problem, stream=None):
        return self.LocalInertialBA_batch([problem], stream)[0]

    def LocalBundleAdjustment(self, problem, stream=None):
        """Optimizer::LocalBundleAdjustment (reference src/Optimizer.cc:1588-2040): the same flattened problem with
        vertex_se3 = 1 (g2o::VertexSE3Expmap keyframes in kf_Rcw / kf_tcw, no inertial edges, 10 iterations)."""
        if not int(problem.get("vertex_se3", 0)):
            raise ValueError("LocalBundleAdjustment expects a vertex_se3 problem (synth.lba_problem layout)")
        return self.LocalInertialBA_batch([problem], stream)[0]

    def LocalInertialBA_batch(self, problems, stream=None):
        self.upload(problems, stream)
        self.solve_uploaded(stream)
        return self.download(stream)

    # resident-problem variant (bench: inputs already on the device)
    def upload(self, problems, stream=None):
        n = len(problems)
        Ps = (BaProblem * n)()
        keep = []
        for i, p in enumerate(problems):
            _, k = pack_problem(p, Ps[i])
            keep.append(k)
        check(self._L.gfs_ba_upload(self._h, stream, C.byref(Ps), n))
        self._uploaded = problems

    def solve_uploaded(self, stream=None):
        check(self._L.gfs_ba_solve_uploaded(self._h, stream))

    def download(self, stream=None):
        problems = self._uploaded
        n = len(problems)
        Rs = (BaResult * n)()
        outs = [alloc_result(p, Rs[i])[1] for i, p in enumerate(problems)]
        check(self._L.gfs_ba_download(self._h, stream, C.byref(Rs), n))
        return [unpack_result(Rs[i], outs[i], problems[i]) for i in range(n)]

    def last_launches(self):
        return self._L.gfs_ba_last_launches(self._h)

    def set_partition_nccl(self, rank, world, group=None):
        """Edge-partitioned mode with NCCL called directly from the solve (no Python in the loop): rank 0 creates the NCCL
        unique id, torch.distributed broadcasts its 128 bytes (any backend), every rank joins the communicator."""
        import numpy as np
        if world > 1:
            import torch
            import torch.distributed as dist
            uid = np.zeros(128, np.uint8)
            if rank == 0:
                check(self._L.gfs_nccl_unique_id(uid.ctypes.data_as(C.c_void_p)))
            dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
            t = torch.from_numpy(uid).to(dev)
            dist.broadcast(t, src=0, group=group)
            uid = t.cpu().numpy()
            check(self._L.gfs_ba_set_partition_nccl(self._h, int(rank), int(world), uid.ctypes.data_as(C.c_void_p)))
        else:
            check(self._L.gfs_ba_set_partition_nccl(self._h, 0, 1, None))

    def last_nccl_calls(self):
        return self._L.gfs_ba_last_nccl_calls(self._h)

    def set_partition(self, rank, world, group=None):
        """Edge-partitioned mode: this rank owns landmarks p % world == rank; the reduced pose system
        is all-reduced over `group` (torch.distributed, NCCL) once per LM trial."""
        from .parallel import ba_allreduce_callback
        self._cb = ba_allreduce_callback(group) if world > 1 else None
        check(self._L.gfs_ba_set_partition(self._h, int(rank), int(world), self._cb, None))
