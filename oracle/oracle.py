"""ctypes loader for the CPU oracle (oracle/libgfs_oracle.so).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (geoflowslam_b200) never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

KP_DTYPE = np.dtype([("x", "<f4"), ("y", "<f4"), ("size", "<f4"), ("angle", "<f4"),
                     ("response", "<f4"), ("octave", "<i4")])


def build(force=False):
    so = os.path.join(_HERE, "libgfs_oracle.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith(".cpp")]
    stale = (not os.path.exists(so)) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs)
    if force or stale:
        subprocess.check_call(["make", "-C", _HERE, "-B" if force else "-s"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is None:
        _LIB = C.CDLL(build())
        L = _LIB
        L.gfo_orb_create.restype = C.c_void_p
        L.gfo_orb_create.argtypes = [C.c_int, C.c_float, C.c_int, C.c_int, C.c_int]
        L.gfo_orb_destroy.argtypes = [C.c_void_p]
        L.gfo_orb_set_threads.argtypes = [C.c_void_p, C.c_int]
        L.gfo_orb_tables.argtypes = [C.c_void_p] + [C.c_void_p] * 3
        L.gfo_orb_level_size.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.gfo_orb_extract.restype = C.c_int
        L.gfo_orb_extract.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                      C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
        L.gfo_orb_get_level.restype = C.c_int
        L.gfo_orb_get_level.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.gfo_orb_get_candidates.restype = C.c_int
        L.gfo_orb_get_candidates.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int]
        L.gfo_resize_area.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int]
        L.gfo_blur7.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
        L.gfo_fast_atan2.restype = C.c_float
        L.gfo_fast_atan2.argtypes = [C.c_float, C.c_float]
        L.gfo_fast.restype = C.c_int
        L.gfo_fast.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int]
        L.gfo_distribute_octtree.restype = C.c_int
        L.gfo_distribute_octtree.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                             C.c_void_p, C.c_int]
        for name, setup in _LATE_BINDERS:
            if hasattr(L, name):
                setup(L)
    return _LIB


_LATE_BINDERS = []  # (symbol, binder) pairs registered by the other oracle modules


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


class OrbOracle:
    """Mirror of ORB_SLAM3::ORBextractor (include/ORBextractor.h:46-120) over the C oracle."""

    def __init__(self, nfeatures=1000, scale_factor=1.2, nlevels=8, ini_th=20, min_th=7, threads=1):
        self.L = lib()
        self.h = self.L.gfo_orb_create(nfeatures, scale_factor, nlevels, ini_th, min_th)
        self.nfeatures, self.nlevels = nfeatures, nlevels
        self.L.gfo_orb_set_threads(self.h, threads)
        self.cap = nfeatures + 4 * nlevels

    def __del__(self):
        try:
            self.L.gfo_orb_destroy(self.h)
        except Exception:
            pass

    def tables(self):
        sf = np.zeros(self.nlevels, np.float32)
        npl = np.zeros(self.nlevels, np.int32)
        um = np.zeros(16, np.int32)
        self.L.gfo_orb_tables(self.h, _p(sf), _p(npl), _p(um))
        return sf, npl, um

    def level_size(self, w, h, l):
        lw, lh = C.c_int(), C.c_int()
        self.L.gfo_orb_level_size(self.h, w, h, l, C.byref(lw), C.byref(lh))
        return lw.value, lh.value

    def extract(self, img, lapping=(0, 0)):
        """-> (keypoints structured array, descriptors (n,32) u8, monoIndex)"""
        img = np.ascontiguousarray(img, np.uint8)
        h, w = img.shape
        kps = np.zeros(self.cap, KP_DTYPE)
        desc = np.zeros((self.cap, 32), np.uint8)
        mono = C.c_int()
        n = self.L.gfo_orb_extract(self.h, _p(img), w, h, w, lapping[0], lapping[1], _p(kps), _p(desc), self.cap,
                                   C.byref(mono))
        if n < 0:
            return kps[:0], desc[:0], -1
        assert n <= self.cap
        self._wh = (w, h)
        return kps[:n].copy(), desc[:n].copy(), mono.value

    def level(self, l, blurred=False):
        lw, lh = self.level_size(self._wh[0], self._wh[1], l)
        out = np.zeros((lh, lw), np.uint8)
        ok = self.L.gfo_orb_get_level(self.h, l, int(blurred), _p(out))
        return out if ok else None

    def candidates(self, l):
        buf = np.zeros((200000, 3), np.float32)
        n = self.L.gfo_orb_get_candidates(self.h, l, _p(buf), buf.shape[0])
        return buf[:n].copy()


def resize_area(src, dw, dh):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros((dh, dw), np.uint8)
    lib().gfo_resize_area(_p(src), src.shape[1], src.shape[0], _p(dst), dw, dh)
    return dst


def blur7(src):
    src = np.ascontiguousarray(src, np.uint8)
    dst = np.zeros_like(src)
    lib().gfo_blur7(_p(src), src.shape[1], src.shape[0], _p(dst))
    return dst


def fast_atan2(y, x):
    return lib().gfo_fast_atan2(float(y), float(x))


def fast(img, thr):
    img = np.ascontiguousarray(img, np.uint8)
    buf = np.zeros((img.size // 4 + 16, 3), np.int32)
    n = lib().gfo_fast(_p(img), img.shape[1], img.shape[0], thr, _p(buf), buf.shape[0])
    return buf[:n].copy()


def distribute_octtree(xyr, minX, maxX, minY, maxY, N):
    xyr = np.ascontiguousarray(xyr, np.float32)
    out = np.zeros((max(N, 1) * 4 + 16, 3), np.float32)
    n = lib().gfo_distribute_octtree(_p(xyr), xyr.shape[0], minX, maxX, minY, maxY, N, _p(out), out.shape[0])
    return out[:n].copy()


# ------------------------------------------------------------------------------ matching
def _bind_match(L):
    L.gfo_descriptor_distance.restype = C.c_int
    L.gfo_descriptor_distance.argtypes = [C.c_void_p, C.c_void_p]
    L.gfo_bf_match.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_int]
    L.gfo_gms_filter.restype = C.c_int
    L.gfo_gms_filter.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                 C.c_void_p, C.c_int, C.c_void_p]


_LATE_BINDERS.append(("gfo_bf_match", _bind_match))


def descriptor_distance(a, b):
    a = np.ascontiguousarray(a, np.uint8); b = np.ascontiguousarray(b, np.uint8)
    return lib().gfo_descriptor_distance(_p(a), _p(b))


def bf_match(dq, dt, threads=1):
    """cv::BFMatcher(NORM_HAMMING).match(dq, dt) -> (trainIdx[nq], distance[nq])"""
    dq = np.ascontiguousarray(dq, np.uint8).reshape(-1, 32)
    dt = np.ascontiguousarray(dt, np.uint8).reshape(-1, 32)
    idx = np.zeros(len(dq), np.int32); dist = np.zeros(len(dq), np.int32)
    lib().gfo_bf_match(_p(dq), len(dq), _p(dt), len(dt), _p(idx), _p(dist), threads)
    return idx, dist


def gms_filter(pts1, size1, pts2, size2, matches):
    """gms_matcher(kp1, size1, kp2, size2, matches).GetInlierMask(mask, false, false)"""
    pts1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2)
    pts2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    matches = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    mask = np.zeros(max(len(matches), 1), np.uint8)
    n = lib().gfo_gms_filter(_p(pts1), len(pts1), size1[0], size1[1], _p(pts2), len(pts2), size2[0], size2[1],
                             _p(matches), len(matches), _p(mask))
    return mask[:len(matches)].astype(bool), n


# ------------------------------------------------------------------------------ GICP
class GicpResult(C.Structure):
    _fields_ = [("T", C.c_double * 16), ("H", C.c_double * 36), ("b", C.c_double * 6), ("error", C.c_double),
                ("iterations", C.c_int), ("num_inliers", C.c_int), ("converged", C.c_int), ("n_target", C.c_int),
                ("n_source", C.c_int), ("inner_evals", C.c_int)]


def _bind_gicp(L):
    L.gfo_gicp_align.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_double, C.c_double, C.c_int,
                                 C.c_int, C.c_double, C.c_double, C.c_int, C.POINTER(GicpResult)]
    L.gfo_voxelgrid.restype = C.c_int
    L.gfo_voxelgrid.argtypes = [C.c_void_p, C.c_int, C.c_double, C.c_void_p, C.c_int]
    L.gfo_kdtree_knn.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gfo_covariances.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p]
    L.gfo_eig3.argtypes = [C.c_void_p] * 3
    L.gfo_se3_exp.argtypes = [C.c_void_p] * 2


_LATE_BINDERS.append(("gfo_gicp_align", _bind_gicp))


def gicp_align(target, source, T0=None, voxel=0.02, max_dist=0.1, k=10, max_iter=20,
               rot_eps=0.1 * np.pi / 180.0, trans_eps=1e-3, threads=4):
    """RegistrationGICP::RegisterPointClouds -> dict(T, H, b, error, iterations, num_inliers, converged, ...)"""
    target = np.ascontiguousarray(target, np.float32).reshape(-1, 4)
    source = np.ascontiguousarray(source, np.float32).reshape(-1, 4)
    T0 = np.ascontiguousarray(np.eye(4) if T0 is None else T0, np.float64)
    r = GicpResult()
    lib().gfo_gicp_align(_p(target), len(target), _p(source), len(source), _p(T0), voxel, max_dist, k, max_iter,
                         rot_eps, trans_eps, threads, C.byref(r))
    return dict(T=np.array(r.T).reshape(4, 4), H=np.array(r.H).reshape(6, 6), b=np.array(r.b), error=r.error,
                iterations=r.iterations, num_inliers=r.num_inliers, converged=bool(r.converged), n_target=r.n_target,
                n_source=r.n_source, inner_evals=r.inner_evals)


def voxelgrid(pts4, leaf):
    pts4 = np.ascontiguousarray(pts4, np.float32).reshape(-1, 4)
    out = np.zeros((max(len(pts4), 1), 3), np.float64)
    n = lib().gfo_voxelgrid(_p(pts4), len(pts4), leaf, _p(out), len(out))
    return out[:n].copy()


def kdtree_knn(pts, queries, k):
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    q = np.ascontiguousarray(queries, np.float64).reshape(-1, 3)
    idx = np.zeros((len(q), k), np.int64); d2 = np.zeros((len(q), k), np.float64); f = np.zeros(len(q), np.int32)
    lib().gfo_kdtree_knn(_p(pts), len(pts), _p(q), len(q), k, _p(idx), _p(d2), _p(f))
    return idx, d2, f


def covariances(pts, k=10):
    pts = np.ascontiguousarray(pts, np.float64).reshape(-1, 3)
    out = np.zeros((len(pts), 6), np.float64)
    lib().gfo_covariances(_p(pts), len(pts), k, _p(out))
    return out


def eig3(A):
    A = np.ascontiguousarray(A, np.float64).reshape(3, 3)
    ev = np.zeros(3); V = np.zeros((3, 3))
    lib().gfo_eig3(_p(A), _p(ev), _p(V))
    return ev, V


def se3_exp(a):
    a = np.ascontiguousarray(a, np.float64).reshape(6)
    T = np.zeros((4, 4))
    lib().gfo_se3_exp(_p(a), _p(T))
    return T


# ------------------------------------------------------------------------------ Local inertial BA
dp, ip, fp, bp_ = C.POINTER(C.c_double), C.POINTER(C.c_int), C.POINTER(C.c_float), C.POINTER(C.c_uint8)


class BaProblem(C.Structure):  # mirrors GfsBaProblem (include/gfs_b200.h)
    _fields_ = [("n_opt_kf", C.c_int), ("n_fixed_kf", C.c_int), ("n_points", C.c_int), ("n_obs", C.c_int),
                ("n_inertial", C.c_int), ("iterations", C.c_int), ("b_large", C.c_int), ("lambda_init", C.c_double),
                ("Rcb", C.c_double * 9), ("tcb", C.c_double * 3), ("Rbc", C.c_double * 9), ("tbc", C.c_double * 3),
                ("fx", C.c_float), ("fy", C.c_float), ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_double),
                ("kf_Rwb", dp), ("kf_twb", dp), ("kf_Rcw", dp), ("kf_tcw", dp), ("kf_vel", dp), ("kf_bg", dp),
                ("kf_ba", dp), ("kf_has_imu", bp_), ("pt_xyz", dp), ("pt_close", bp_), ("obs_kf", ip), ("obs_pt", ip),
                ("obs_uvr", dp), ("obs_inv_sigma2", fp), ("in_kf1", ip), ("in_kf2", ip), ("in_pre", fp),
                ("in_downweight", bp_), ("n_icp", C.c_int), ("icp_kf1", ip), ("icp_kf2", ip), ("icp_Rt", dp),
                ("vertex_se3", C.c_int)]


class BaResult(C.Structure):  # mirrors GfsBaResult
    _fields_ = [("kf_Rwb", dp), ("kf_twb", dp), ("kf_Rcw", dp), ("kf_tcw", dp), ("kf_vel", dp), ("kf_bg", dp),
                ("kf_ba", dp), ("pt_xyz", dp), ("obs_chi2", dp), ("obs_depth_positive", bp_), ("obs_outlier", bp_),
                ("err", C.c_float), ("err_end", C.c_float), ("failed", C.c_int), ("iterations_done", C.c_int),
                ("lm_trials", C.c_int), ("lambda_final", C.c_double)]


_BA_ARRAYS = [("kf_Rwb", np.float64), ("kf_twb", np.float64), ("kf_Rcw", np.float64), ("kf_tcw", np.float64),
              ("kf_vel", np.float64), ("kf_bg", np.float64), ("kf_ba", np.float64), ("kf_has_imu", np.uint8),
              ("pt_xyz", np.float64), ("pt_close", np.uint8), ("obs_kf", np.int32), ("obs_pt", np.int32),
              ("obs_uvr", np.float64), ("obs_inv_sigma2", np.float32), ("in_kf1", np.int32), ("in_kf2", np.int32),
              ("in_pre", np.float32), ("in_downweight", np.uint8)]
_BA_OPTIONAL = [("icp_kf1", np.int32), ("icp_kf2", np.int32), ("icp_Rt", np.float64)]


def ba_pack(prob, struct_cls=BaProblem):
    """dict (geoflowslam_b200.synth.ba_problem layout) -> (ctypes struct, keep-alive list)"""
    P = struct_cls()
    keep = []
    for k in ("n_opt_kf", "n_fixed_kf", "n_points", "n_obs", "n_inertial", "iterations", "b_large"):
        setattr(P, k, int(prob[k]))
    P.lambda_init = float(prob["lambda_init"])
    for k, n in (("Rcb", 9), ("tcb", 3), ("Rbc", 9), ("tbc", 3)):
        setattr(P, k, (C.c_double * n)(*np.asarray(prob[k], np.float64).ravel()))
    for k in ("fx", "fy", "cx", "cy"):
        setattr(P, k, float(prob[k]))
    P.bf = float(prob["bf"])
    P.n_icp = int(prob.get("n_icp", 0))
    P.vertex_se3 = int(prob.get("vertex_se3", 0))
    for k, dt in _BA_OPTIONAL:
        a = np.ascontiguousarray(prob.get(k, np.zeros(0)), dt)
        keep.append(a)
        setattr(P, k, a.ctypes.data_as(dict(type(P)._fields_)[k]))
    for k, dt in _BA_ARRAYS:
        a = np.ascontiguousarray(prob[k], dt)
        keep.append(a)
        setattr(P, k, a.ctypes.data_as(dict(P._fields_)[k]))
    return P, keep


def ba_alloc_result(prob, struct_cls=BaResult):
    nk = prob["n_opt_kf"] + prob["n_fixed_kf"]
    out = dict(kf_Rwb=np.zeros((nk, 9)), kf_twb=np.zeros((nk, 3)), kf_Rcw=np.zeros((nk, 9)), kf_tcw=np.zeros((nk, 3)),
               kf_vel=np.zeros((nk, 3)), kf_bg=np.zeros((nk, 3)), kf_ba=np.zeros((nk, 3)),
               pt_xyz=np.zeros((prob["n_points"], 3)), obs_chi2=np.zeros(max(prob["n_obs"], 1)),
               obs_depth_positive=np.zeros(max(prob["n_obs"], 1), np.uint8),
               obs_outlier=np.zeros(max(prob["n_obs"], 1), np.uint8))
    R = struct_cls()
    for k, a in out.items():
        setattr(R, k, a.ctypes.data_as(dict(R._fields_)[k]))
    return R, out


def ba_unpack_result(R, out, prob):
    res = {k: v[:prob["n_obs"]] if k.startswith("obs_") else v for k, v in out.items()}
    res.update(err=R.err, err_end=R.err_end, failed=bool(R.failed), iterations_done=R.iterations_done,
               lm_trials=R.lm_trials, lambda_final=R.lambda_final)
    return res


def _bind_ba(L):
    L.gfo_ba_solve.argtypes = [C.POINTER(BaProblem), C.POINTER(BaResult)]
    L.gfo_ba_chi2.restype = C.c_double
    L.gfo_ba_chi2.argtypes = [C.POINTER(BaProblem)]
    L.gfo_ba_inertial.argtypes = [C.POINTER(BaProblem), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p]
    L.gfo_ba_system.argtypes = [C.POINTER(BaProblem)] + [C.c_void_p] * 5
    L.gfo_ba_step.restype = C.c_int
    L.gfo_ba_step.argtypes = [C.POINTER(BaProblem), C.c_double, C.c_void_p]
    L.gfo_ba_chi2_at.restype = C.c_double
    L.gfo_ba_chi2_at.argtypes = [C.POINTER(BaProblem), C.c_void_p, C.c_int]
    L.gfo_so3.argtypes = [C.c_void_p] * 5


_LATE_BINDERS.append(("gfo_ba_solve", _bind_ba))


def ba_solve(prob):
    """Optimizer::LocalInertialBA numerical core -> dict of optimised states + per-edge outcomes"""
    P, keep = ba_pack(prob)
    R, out = ba_alloc_result(prob)
    lib().gfo_ba_solve(C.byref(P), C.byref(R))
    return ba_unpack_result(R, out, prob)


def ba_system(prob):
    P, keep = ba_pack(prob)
    n, m, o = 15 * prob["n_opt_kf"], prob["n_points"], prob["n_obs"]
    Hpp = np.zeros((n, n)); bpv = np.zeros(n); Hll = np.zeros((m, 3, 3)); bl = np.zeros((m, 3)); Hpl = np.zeros((o, 6, 3))
    lib().gfo_ba_system(C.byref(P), _p(Hpp), _p(bpv), _p(Hll), _p(bl), _p(Hpl))
    return Hpp, bpv, Hll, bl, Hpl


def ba_step(prob, lam):
    P, keep = ba_pack(prob)
    x = np.zeros(15 * prob["n_opt_kf"] + 3 * prob["n_points"])
    ok = lib().gfo_ba_step(C.byref(P), float(lam), _p(x))
    return bool(ok), x


def ba_chi2_at(prob, x, robust=True):
    P, keep = ba_pack(prob)
    x = np.ascontiguousarray(x, np.float64)
    return lib().gfo_ba_chi2_at(C.byref(P), _p(x), int(robust))


def ba_inertial(prob, e):
    P, keep = ba_pack(prob)
    err = np.zeros(9); J = np.zeros((9, 24)); info = np.zeros((9, 9))
    lib().gfo_ba_inertial(C.byref(P), e, _p(err), _p(J), _p(info))
    return err, J, info


def so3(w):
    w = np.ascontiguousarray(w, np.float64)
    R = np.zeros((3, 3)); lg = np.zeros(3); Jr = np.zeros((3, 3)); Ji = np.zeros((3, 3))
    lib().gfo_so3(_p(w), _p(R), _p(lg), _p(Jr), _p(Ji))
    return R, lg, Jr, Ji


# ------------------------------------------------------------------------------ projection search / depth
PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("angle", "<f4"),
                             ("min_level", "<i4"), ("max_level", "<i4"), ("blocks", "<i4"), ("desc", "u1", (32,))])


def _bind_proj(L):
    L.gfo_search_by_projection.restype = C.c_int
    L.gfo_search_by_projection.argtypes = [C.c_int, C.c_float, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                           C.c_void_p, C.c_void_p, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                           C.c_void_p]
    L.gfo_depth_to_cloud.restype = C.c_int
    L.gfo_depth_to_cloud.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_float, C.c_float, C.c_float, C.c_float,
                                     C.c_void_p, C.c_int]


_LATE_BINDERS.append(("gfo_search_by_projection", _bind_proj))


def search_by_projection(mode, queries, kps_un, u_right, desc, occupied, grid, nnratio=0.9, check_orientation=True):
    q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
    k = np.ascontiguousarray(kps_un, KP_DTYPE)
    ur = np.ascontiguousarray(u_right, np.float32)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    assign = np.full(max(len(k), 1), -1, np.int32)
    n = lib().gfo_search_by_projection(int(mode), float(nnratio), int(check_orientation), _p(q), len(q), _p(k), _p(ur), _p(d),
                                       None if oc is None else _p(oc), len(k), float(grid[0]), float(grid[1]),
                                       float(grid[2]), float(grid[3]), _p(assign))
    return assign[:len(k)], n


def depth_to_cloud(depth, stride, fx, fy, cx, cy):
    depth = np.ascontiguousarray(depth, np.float32)
    h, w = depth.shape
    cap = ((w + stride - 1) // stride) * ((h + stride - 1) // stride)
    out = np.zeros((cap, 4), np.float32)
    n = lib().gfo_depth_to_cloud(_p(depth), w, h, int(stride), fx, fy, cx, cy, _p(out), cap)
    return out[:n].copy()


# ------------------------------------------------------------------------------ PoseOptimization
class PoseProblem(C.Structure):  # mirrors GfsPoseProblem (include/gfs_b200.h)
    _fields_ = [("n_obs", C.c_int), ("q_wxyz", C.c_float * 4), ("t", C.c_float * 3), ("fx", C.c_float), ("fy", C.c_float),
                ("cx", C.c_float), ("cy", C.c_float), ("bf", C.c_float), ("Xw", dp), ("uvr", fp), ("inv_sigma2", fp)]


class PoseResult(C.Structure):  # mirrors GfsPoseResult
    _fields_ = [("n_inliers", C.c_int), ("n_bad", C.c_int), ("n_good", C.c_int), ("avg_reproj_error", C.c_float),
                ("rounds_done", C.c_int), ("lm_iterations", C.c_int * 4), ("q_wxyz", C.c_double * 4), ("t", C.c_double * 3),
                ("outlier", bp_), ("chi2", fp)]


def pose_pack(prob):
    P = PoseProblem()
    n = int(prob["n_obs"])
    P.n_obs = n
    P.q_wxyz = (C.c_float * 4)(*np.asarray(prob["q_wxyz"], np.float32))
    P.t = (C.c_float * 3)(*np.asarray(prob["t"], np.float32))
    for k in ("fx", "fy", "cx", "cy", "bf"):
        setattr(P, k, float(np.float32(prob[k])))
    keep = [np.ascontiguousarray(prob["Xw"], np.float64).reshape(-1, 3), np.ascontiguousarray(prob["uvr"], np.float32).reshape(-1, 3),
            np.ascontiguousarray(prob["inv_sigma2"], np.float32).reshape(-1)]
    P.Xw, P.uvr, P.inv_sigma2 = keep[0].ctypes.data_as(dp), keep[1].ctypes.data_as(fp), keep[2].ctypes.data_as(fp)
    return P, keep


def _bind_pose(L):
    L.gfo_pose_optimize.argtypes = [C.POINTER(PoseProblem), C.POINTER(PoseResult)]
    L.gfo_pose_edge.restype = C.c_int
    L.gfo_pose_edge.argtypes = [C.POINTER(PoseProblem), C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.gfo_pose_oplus.argtypes = [C.c_void_p] * 5
    L.gfo_eigen_ldlt_solve.restype = C.c_int
    L.gfo_eigen_ldlt_solve.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]


_LATE_BINDERS.append(("gfo_pose_optimize", _bind_pose))


def pose_optimize(prob):
    """Optimizer::PoseOptimization -> dict(n_inliers, outlier, chi2, avg_reproj_error, q_wxyz, t, ...)"""
    P, keep = pose_pack(prob)
    n = int(prob["n_obs"])
    R = PoseResult()
    outl, chi2 = np.zeros(max(n, 1), np.uint8), np.zeros(max(n, 1), np.float32)
    R.outlier, R.chi2 = outl.ctypes.data_as(bp_), chi2.ctypes.data_as(fp)
    lib().gfo_pose_optimize(C.byref(P), C.byref(R))
    return dict(n_inliers=R.n_inliers, n_bad=R.n_bad, n_good=R.n_good, avg_reproj_error=float(R.avg_reproj_error),
                rounds_done=R.rounds_done, lm_iterations=list(R.lm_iterations), q_wxyz=np.array(R.q_wxyz), t=np.array(R.t),
                outlier=outl[:n].astype(bool), chi2=chi2[:n].copy())


def pose_edge(prob, q_wxyz, t, e):
    P, keep = pose_pack(prob)
    q = np.ascontiguousarray(q_wxyz, np.float64); tt = np.ascontiguousarray(t, np.float64)
    err = np.zeros(3); J = np.zeros(18)
    d = lib().gfo_pose_edge(C.byref(P), _p(q), _p(tt), int(e), _p(err), _p(J))
    return err[:d].copy(), J[:6 * d].reshape(d, 6).copy()


def pose_oplus(q_wxyz, t, u6):
    q = np.ascontiguousarray(q_wxyz, np.float64); tt = np.ascontiguousarray(t, np.float64); u = np.ascontiguousarray(u6, np.float64)
    qo = np.zeros(4); to = np.zeros(3)
    lib().gfo_pose_oplus(_p(q), _p(tt), _p(u), _p(qo), _p(to))
    return qo, to


def eigen_ldlt_solve(H, b):
    H = np.ascontiguousarray(H, np.float64); b = np.ascontiguousarray(b, np.float64)
    x = np.zeros(len(b))
    ok = lib().gfo_eigen_ldlt_solve(_p(H), _p(b), len(b), _p(x))
    return bool(ok), x


# ---- PoseInertialOptimizationLastKeyFrame / LastFrame (ba_oracle.cpp, namespace pin)
def _bind_pin(L):
    from geoflowslam_b200.pose_inertial import PoseInertialProblem, PoseInertialResult
    PP = C.POINTER(PoseInertialProblem)
    L.gfo_pose_inertial_optimize.argtypes = [PP, C.POINTER(PoseInertialResult)]
    L.gfo_pin_vis_edge.restype = C.c_int
    L.gfo_pin_vis_edge.argtypes = [PP, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    L.gfo_pin_pose_update.argtypes = [PP] + [C.c_void_p] * 5
    L.gfo_pin_prior_edge.argtypes = [PP] + [C.c_void_p] * 7
    L.gfo_pin_marginalize.argtypes = [C.c_void_p, C.c_void_p]
    L.gfo_pin_clamp.argtypes = [C.c_void_p]


_LATE_BINDERS.append(("gfo_pose_inertial_optimize", _bind_pin))


def pose_inertial_optimize(prob):
    """Optimizer::PoseInertialOptimizationLast{KeyFrame,Frame} -> dict (same keys as the product's result)"""
    from geoflowslam_b200 import pose_inertial as pi
    P, keep = pi.pack_problem(prob)
    n = int(prob["n_obs"])
    R, out = pi.alloc_result(n)
    lib().gfo_pose_inertial_optimize(C.byref(P), C.byref(R))
    return pi.unpack_result(R, out, n)


def pin_vis_edge(prob, Rwb, twb, e):
    from geoflowslam_b200 import pose_inertial as pi
    P, keep = pi.pack_problem(prob)
    R = np.ascontiguousarray(Rwb, np.float64); t = np.ascontiguousarray(twb, np.float64)
    err = np.zeros(3); J = np.zeros(18)
    d = lib().gfo_pin_vis_edge(C.byref(P), _p(R), _p(t), int(e), _p(err), _p(J))
    return err[:d].copy(), J[:6 * d].reshape(d, 6).copy()


def pin_pose_update(prob, Rwb, twb, u6):
    from geoflowslam_b200 import pose_inertial as pi
    P, keep = pi.pack_problem(prob)
    R = np.ascontiguousarray(Rwb, np.float64); t = np.ascontiguousarray(twb, np.float64); u = np.ascontiguousarray(u6, np.float64)
    Ro = np.zeros(9); to = np.zeros(3)
    lib().gfo_pin_pose_update(C.byref(P), _p(R), _p(t), _p(u), _p(Ro), _p(to))
    return Ro.reshape(3, 3), to


def pin_prior_edge(prob, Rwb, twb, v, bg, ba):
    from geoflowslam_b200 import pose_inertial as pi
    P, keep = pi.pack_problem(prob)
    a = [np.ascontiguousarray(x, np.float64) for x in (Rwb, twb, v, bg, ba)]
    err = np.zeros(15); J = np.zeros(225)
    lib().gfo_pin_prior_edge(C.byref(P), *[_p(x) for x in a], _p(err), _p(J))
    return err, J.reshape(15, 15)


def pin_marginalize(H30):
    H = np.ascontiguousarray(H30, np.float64); out = np.zeros((15, 15))
    lib().gfo_pin_marginalize(_p(H), _p(out))
    return out


def pin_clamp(H15):
    H = np.array(H15, np.float64, order="C", copy=True)
    lib().gfo_pin_clamp(_p(H))
    return H


# ---- optical-flow front end (klt_oracle.cpp): buildOpticalFlowPyramid, calcOpticalFlowPyrLK, fbKltTracking
def _bind_klt(L):
    vp_, ci_, cf_ = C.c_void_p, C.c_int, C.c_float
    L.gfo_pyr_down.argtypes = [vp_, ci_, ci_, vp_]
    L.gfo_scharr.argtypes = [vp_, ci_, ci_, vp_]
    L.gfo_clahe.argtypes = [vp_, ci_, ci_, C.c_double, ci_, ci_, vp_]
    L.gfo_klt_pyramid_pixels.restype = ci_
    L.gfo_klt_pyramid_pixels.argtypes = [ci_, ci_, ci_]
    L.gfo_klt_build_pyramid.argtypes = [vp_, ci_, ci_, ci_, vp_, vp_]
    L.gfo_klt_calc.argtypes = [vp_, vp_, vp_, vp_, ci_, ci_, ci_, vp_, vp_, ci_, ci_, ci_, ci_, cf_, ci_, vp_, vp_]
    L.gfo_fb_klt.argtypes = [vp_, vp_, vp_, vp_, ci_, ci_, ci_, vp_, vp_, ci_, ci_, ci_, cf_, cf_, vp_]


_LATE_BINDERS.append(("gfo_fb_klt", _bind_klt))


def pyr_down(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros(((h + 1) // 2, (w + 1) // 2), np.uint8)
    lib().gfo_pyr_down(_p(img), w, h, _p(out))
    return out


def clahe(img, clip_limit=3.0, tiles=(8, 8)):
    """cv::createCLAHE(clip_limit, tiles)->apply(img) for an 8-bit gray image (reference src/Frame.cc:366-368)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros_like(img)
    lib().gfo_clahe(_p(img), w, h, float(clip_limit), int(tiles[0]), int(tiles[1]), _p(out))
    return out


def scharr_deriv(img):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    out = np.zeros((h, w, 2), np.int16)
    lib().gfo_scharr(_p(img), w, h, _p(out))
    return out


def klt_level_shapes(w, h, levels):
    out = []
    for _ in range(levels + 1):
        out.append((h, w))
        w, h = (w + 1) // 2, (h + 1) // 2
    return out


def klt_build_pyramid(img, levels=3):
    """cv::buildOpticalFlowPyramid(img, pyr, winSize, levels) -> (packed images u8, packed derivatives int16)"""
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    n = lib().gfo_klt_pyramid_pixels(w, h, levels)
    pi = np.zeros(n, np.uint8); pd = np.zeros(2 * n, np.int16)
    lib().gfo_klt_build_pyramid(_p(img), w, h, levels, _p(pi), _p(pd))
    return pi, pd


def klt_unpack(pyr, w, h, levels):
    pi, pd = pyr
    out, off = [], 0
    for (lh, lw) in klt_level_shapes(w, h, levels):
        out.append((pi[off:off + lh * lw].reshape(lh, lw), pd[2 * off:2 * (off + lh * lw)].reshape(lh, lw, 2)))
        off += lh * lw
    return out


def klt_calc(prev_pyr, cur_pyr, w, h, levels, pts, init=None, win=35, max_level=3, max_count=30, eps=0.01):
    """cv::calcOpticalFlowPyrLK(prevPyr, curPyr, ..., OPTFLOW_LK_GET_MIN_EIGENVALS [| USE_INITIAL_FLOW])
    -> (next points (n,2) f32, status (n,) u8, min eigenvalues (n,) f32)"""
    pts = np.ascontiguousarray(pts, np.float32).reshape(-1, 2)
    n = len(pts)
    nxt = np.ascontiguousarray(pts if init is None else init, np.float32).reshape(-1, 2).copy()
    st = np.zeros(max(n, 1), np.uint8); er = np.zeros(max(n, 1), np.float32)
    lib().gfo_klt_calc(_p(prev_pyr[0]), _p(prev_pyr[1]), _p(cur_pyr[0]), _p(cur_pyr[1]), w, h, levels, _p(pts), _p(nxt), n, win,
                       max_level, max_count, eps, 0 if init is None else 1, _p(st), _p(er))
    return nxt, st[:n], er[:n]


def fb_klt_tracking(prev_pyr, cur_pyr, w, h, levels, kps, priors, win=35, nbpyrlvl=3, ferr=15.0, max_fbklt_dist=0.5):
    """ORBmatcher::fbKltTracking -> (priors after tracking (n,2) f32, vkpstatus (n,) bool)"""
    kps = np.ascontiguousarray(kps, np.float32).reshape(-1, 2)
    pr = np.ascontiguousarray(priors, np.float32).reshape(-1, 2).copy()
    n = len(kps)
    st = np.zeros(max(n, 1), np.uint8)
    lib().gfo_fb_klt(_p(prev_pyr[0]), _p(prev_pyr[1]), _p(cur_pyr[0]), _p(cur_pyr[1]), w, h, levels, _p(kps), _p(pr), n, win, nbpyrlvl,
                     ferr, max_fbklt_dist, _p(st))
    return pr, st[:n].astype(bool)


# ---- IMU preintegration (imu_oracle.cpp)
def _bind_imu(L):
    L.gfo_imu_preintegrate.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_float, C.c_float, C.c_float, C.c_float, C.c_void_p]


_LATE_BINDERS.append(("gfo_imu_preintegrate", _bind_imu))


def imu_preintegrate(meas, bias, ng, na, ngw, naw):
    """IMU::Preintegrated(bias, calib) + IntegrateNewMeasurement over meas rows (ax ay az wx wy wz dt) -> record (292,) f32"""
    m = np.ascontiguousarray(meas, np.float32).reshape(-1, 7)
    b = np.ascontiguousarray(bias, np.float32).reshape(6)
    out = np.zeros(292, np.float32)
    lib().gfo_imu_preintegrate(_p(m), len(m), _p(b), float(ng), float(na), float(ngw), float(naw), _p(out))
    return out


# ---- ORBmatcher::SearchByProjectionWithOF (reference src/ORBmatcher.cc:2303-2497): host logic around
# fbKltTracking, restated statement by statement (test infrastructure: the product version is
# geoflowslam_b200.matcher.ORBmatcher.SearchByProjectionWithOF)
def search_by_projection_with_of(cur_keys, last_keys, last_mp_state, last_mp_world, Rcw, tcw, K, bounds, prev_img, cur_img, mask,
                                 winsize=35, F_THRESHOLD=1.0, DIST_THRESHOLD=10, tracker=None, fundamental=None):
    """last_mp_state[i]: 0 = no map point (nullptr), 1 = usable map point, 2 = bad map point or outlier (skipped, :2349).
    K = (fx, fy, cx, cy), bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY).  tracker(prev_img, cur_img, kps, priors, win, nbpyrlvl,
    ferr, max_dist) -> (priors, status) is fbKltTracking; fundamental(p1, p2, thr) -> uchar status or None is
    cv::findFundamentalMat(p1, p2, FM_RANSAC, thr, 0.99, status).  Returns (nbgood, tracked, mask) with tracked = the
    (index into the last frame, new position) pairs in the order Frame::AddPts receives them."""
    import cv2
    f32 = np.float32
    H, W = mask.shape
    fx, fy, cx, cy = (f32(v) for v in K)
    mnMinX, mnMaxX, mnMinY, mnMaxY = (f32(v) for v in bounds)
    if fundamental is None:
        def fundamental(p1, p2, thr):
            _, st = cv2.findFundamentalMat(np.asarray(p1, f32), np.asarray(p2, f32), cv2.FM_RANSAC, float(thr), 0.99)
            return None if st is None else st.ravel()
    for kp in np.asarray(cur_keys, f32).reshape(-1, 2):                                       # :2327-2334
        if kp[0] > 0 and kp[0] < W and kp[1] > 0 and kp[1] < H:
            mask[int(kp[1]), int(kp[0])] = 255
    last_keys = np.asarray(last_keys, f32).reshape(-1, 2)
    R = np.asarray(Rcw, f32).reshape(3, 3); t = np.asarray(tcw, f32).reshape(3)
    v3dkps, v3dpriors, v3dkpids, v2dkps, v2dpriors, v2dkpids = [], [], [], [], [], []
    for cnt in range(len(last_keys)):                                                         # :2336-2374
        u, v = last_keys[cnt]
        if last_mp_state[cnt] == 0:
            v2dkps.append((u, v)); v2dpriors.append((u, v)); v2dkpids.append(cnt)
            continue
        if last_mp_state[cnt] == 2:
            continue
        x = np.asarray(last_mp_world[cnt], f32)
        xc = f32(f32(f32(R[0, 0] * x[0]) + f32(R[0, 1] * x[1])) + f32(R[0, 2] * x[2])) + t[0]
        yc = f32(f32(f32(R[1, 0] * x[0]) + f32(R[1, 1] * x[1])) + f32(R[1, 2] * x[2])) + t[1]
        zc = f32(f32(f32(R[2, 0] * x[0]) + f32(R[2, 1] * x[1])) + f32(R[2, 2] * x[2])) + t[2]
        with np.errstate(divide="ignore"):
            invzc = f32(np.float64(1.0) / np.float64(zc))
        pu = f32(f32(f32(fx * xc) * invzc) + cx)
        pv = f32(f32(f32(fy * yc) * invzc) + cy)
        if invzc < 0 or pu < mnMinX or pu > mnMaxX or pv < mnMinY or pv > mnMaxY:
            v2dkps.append((u, v)); v2dpriors.append((u, v)); v2dkpids.append(cnt)
            continue
        v3dkps.append((u, v)); v3dpriors.append((pu, pv)); v3dkpids.append(cnt)
    nbgood, tracked = 0, []

    def near(pt):                                                                             # isPointNearby :2299-2301
        return mask[int(pt[1]), int(pt[0])] == 255

    def update(pt):                                                                           # updateMask :2295-2297
        cv2.circle(mask, (int(round(float(pt[0]))), int(round(float(pt[1])))), int(DIST_THRESHOLD), 255, cv2.FILLED)

    def fcheck(kps, priors, status, thr):                                                     # :2389-2407, :2453-2470
        index = [i for i in range(len(status)) if status[i]]
        if len(index) > 8:
            st = fundamental([kps[i] for i in index], [priors[i] for i in index], thr)
            if st is not None:
                for i in range(len(st)):
                    if not st[i]:
                        status[index[i]] = False

    if v3dpriors:                                                                             # :2380-2437
        pr, status = tracker(prev_img, cur_img, np.array(v3dkps, f32), np.array(v3dpriors, f32), winsize, 3, 15.0, 0.5)
        status = [bool(s) for s in status]
        fcheck(v3dkps, pr, status, F_THRESHOLD)
        for i in range(len(v3dkps)):
            if status[i]:
                if near(pr[i]):
                    continue
                tracked.append((v3dkpids[i], (float(pr[i][0]), float(pr[i][1]))))
                nbgood += 1
                update(pr[i])
            else:
                u, v = last_keys[v3dkpids[i]]
                v2dkps.append((u, v)); v2dpriors.append((u, v)); v2dkpids.append(v3dkpids[i])
    if v2dkps:                                                                                # :2440-2492
        pr, status = tracker(prev_img, cur_img, np.array(v2dkps, f32), np.array(v2dpriors, f32), winsize, 6, 15.0, 0.5)
        status = [bool(s) for s in status]
        fcheck(v2dkps, pr, status, F_THRESHOLD * 0.5)
        for i in range(len(v2dkps)):
            if status[i]:
                if near(pr[i]):
                    continue
                tracked.append((v2dkpids[i], (float(pr[i][0]), float(pr[i][1]))))
                update(pr[i])
                nbgood += 1
    return nbgood, tracked, mask


def fb_klt_tracking_images(prev_img, cur_img, kps, priors, win=35, nbpyrlvl=3, ferr=15.0, max_fbklt_dist=0.5, levels=3):
    """fbKltTracking on two gray images: the pyramids Frame::Frame builds (maxLevel 3, src/Frame.cc:370-373) + fb_klt_tracking."""
    h, w = prev_img.shape
    return fb_klt_tracking(klt_build_pyramid(prev_img, levels), klt_build_pyramid(cur_img, levels), w, h, levels, kps, priors, win,
                           nbpyrlvl, ferr, max_fbklt_dist)


def filter_outliers(keys, has_mp, mp_world, Rcw, tcw, K, F_THRESHOLD, fundamental=None):
    """ORBmatcher::FilterOutliers (reference src/ORBmatcher.cc:208-246), statement by statement.  Returns (inliers,
    mvbOutlier).  As in the reference, the status of the j-th point handed to findFundamentalMat is written to
    mvbOutlier[j] -- the position in the filtered list, not the keypoint's own index (:237-240)."""
    import cv2
    f32 = np.float32
    fx, fy, cx, cy = (f32(v) for v in K)
    R = np.asarray(Rcw, f32).reshape(3, 3); t = np.asarray(tcw, f32).reshape(3)
    keys = np.asarray(keys, f32).reshape(-1, 2)
    outlier = np.zeros(len(keys), bool)
    un_cur, un_forw = [], []
    for i in range(len(keys)):
        if not has_mp[i]:
            continue
        outlier[i] = False
        x = np.asarray(mp_world[i], f32)
        xc = f32(f32(f32(R[0, 0] * x[0]) + f32(R[0, 1] * x[1])) + f32(R[0, 2] * x[2])) + t[0]
        yc = f32(f32(f32(R[1, 0] * x[0]) + f32(R[1, 1] * x[1])) + f32(R[1, 2] * x[2])) + t[1]
        zc = f32(f32(f32(R[2, 0] * x[0]) + f32(R[2, 1] * x[1])) + f32(R[2, 2] * x[2])) + t[2]
        with np.errstate(divide="ignore"):
            invzc = f32(np.float64(1.0) / np.float64(zc))
        if invzc < 0:
            continue
        # Pinhole::project(const Eigen::Vector3f&) (CameraModels/Pinhole.cpp:43-49): fx * x / z + cx, left to right in float32
        un_forw.append((keys[i][0], keys[i][1]))
        un_cur.append((f32(f32(f32(fx * xc) / zc) + cx), f32(f32(f32(fy * yc) / zc) + cy)))
    inliers = 0
    if len(un_cur) > 8:
        if fundamental is None:
            _, st = cv2.findFundamentalMat(np.array(un_cur, f32), np.array(un_forw, f32), cv2.FM_RANSAC, float(F_THRESHOLD), 0.99)
            st = None if st is None else st.ravel()
        else:
            st = fundamental(np.array(un_cur, f32), np.array(un_forw, f32), F_THRESHOLD)
        if st is not None:
            for j in range(len(st)):
                if not st[j]:
                    outlier[j] = True
                else:
                    inliers += 1
    return inliers, outlier
