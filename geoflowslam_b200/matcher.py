"""Host-side mirror of the descriptor-matching half of ORB_SLAM3::ORBmatcher
(reference include/ORBmatcher.h:36-151, src/ORBmatcher.cc) over the C ABI."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr


class ORBmatcher:
    TH_LOW, TH_HIGH, HISTO_LENGTH = 50, 100, 30  # include/ORBmatcher.h:139-141

    def __init__(self, nnratio=0.6, checkOri=True):
        self.mfNNratio, self.mbCheckOrientation = float(nnratio), bool(checkOri)
        self._L = _lib.lib()

    @staticmethod
    def DescriptorDistance(a, b):
        """ORBmatcher::DescriptorDistance (src/ORBmatcher.cc:2536-2550), evaluated on the device."""
        a = np.ascontiguousarray(a, np.uint8).reshape(1, 32)
        b = np.ascontiguousarray(b, np.uint8).reshape(1, 32)
        _, dist = ORBmatcher.bf_match(a, b)
        return int(dist[0])

    @staticmethod
    def bf_match(dq, dt, stream=None):
        """cv::BFMatcher(NORM_HAMMING).match(dq, dt): (trainIdx[nq], distance[nq])."""
        L = _lib.lib()
        dq = np.ascontiguousarray(dq, np.uint8).reshape(-1, 32)
        dt = np.ascontiguousarray(dt, np.uint8).reshape(-1, 32)
        idx = np.full(len(dq), -1, np.int32)
        dist = np.full(len(dq), -1, np.int32)
        check(L.gfs_match_bf_hamming(stream, ptr(dq), len(dq), ptr(dt), len(dt), ptr(idx), ptr(dist)))
        return idx, dist

    @staticmethod
    def gms_filter(kp1, size1, kp2, size2, matches, stream=None):
        """gms_matcher(kp1, size1, kp2, size2, matches).GetInlierMask(mask, false, false)
        -> (mask[nm] bool, num_inliers).  kp*: KP_DTYPE arrays or (n,2) float arrays of pt."""
        L = _lib.lib()
        k1, k2 = _as_kp(kp1), _as_kp(kp2)
        m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
        mask = np.zeros(max(len(m), 1), np.uint8)
        cnt = C.c_int()
        check(L.gfs_gms_filter(stream, ptr(k1), len(k1), int(size1[0]), int(size1[1]), ptr(k2), len(k2),
                               int(size2[0]), int(size2[1]), ptr(m), len(m), ptr(mask), C.byref(cnt)))
        return mask[:len(m)].astype(bool), cnt.value

    def SearchWithGMS(self, kp1, d1, kp2, d2, frame_size, stream=None):
        """The descriptor part of ORBmatcher::SearchWithGMS (src/ORBmatcher.cc:744-778): BF match
        d1 -> d2, GMS mask.  Returns (matches (nq,2) [queryIdx, trainIdx], inlier mask, nmatches);
        the MapPoint transfer that follows stays with the caller's data model."""
        idx, _ = self.bf_match(d1, d2, stream)
        m = np.stack([np.arange(len(idx), dtype=np.int32), idx], 1)
        mask, n = self.gms_filter(kp1, frame_size, kp2, frame_size, m, stream)
        return m, mask, n


def _default_fundamental(p1, p2, thr):
    """cv::findFundamentalMat(p1, p2, FM_RANSAC, thr, 0.99, status) -- the OpenCV call the reference makes on the host
    (src/ORBmatcher.cc:2399,2463); its RANSAC draws from OpenCV's own RNG, so it stays OpenCV's."""
    import cv2
    _, st = cv2.findFundamentalMat(np.ascontiguousarray(p1, np.float32), np.ascontiguousarray(p2, np.float32), cv2.FM_RANSAC,
                                   float(thr), 0.99)
    return None if st is None else st.ravel().astype(bool)


def search_by_projection_with_of(tracker, cur_keys, last_keys, last_mp_state, last_mp_world, Rcw, tcw, K, bounds, prev_img, cur_img,
                                 mask, winsize=35, F_THRESHOLD=1.0, DIST_THRESHOLD=10, fundamental=None):
    """ORBmatcher::SearchByProjectionWithOF (src/ORBmatcher.cc:2303-2497), the caller of fbKltTracking in the dual-stream
    front end (Tracking.cc:3499,3623,3646), on plain arrays instead of Frame / MapPoint objects.

    tracker: `KltTracker.fbKltTracking`-compatible callable (prev_img, cur_img, kps, priors, win, nbpyrlvl, ferr, max_dist)
    -> (priors, status) -- the CUDA path.  last_mp_state[i]: 0 = no map point, 1 = usable map point, 2 = bad map point or
    outlier (skipped).  K = (fx, fy, cx, cy); bounds = (mnMinX, mnMaxX, mnMinY, mnMaxY); mask (h, w) u8 is updated in place.
    Returns (nbgood, tracked_ids (m,) int32 into the last frame, tracked_pts (m, 2) float32) in the order the reference hands
    them to Frame::AddPts: points with a map point that projects into the image are tracked from their projection over 3
    pyramid levels, everything else (and the failures of the first pass) from the old position over all levels; both
    passes are filtered by OpenCV's fundamental-matrix RANSAC and by the occupancy mask (DIST_THRESHOLD-px discs).
    Sophus applies Tcw through its unit quaternion; here R x + t in float32, so a prior can differ from the reference's in
    the last ulp."""
    import cv2
    f32 = np.float32
    fundamental = fundamental or _default_fundamental
    H, W = mask.shape
    ck = np.asarray(cur_keys, f32).reshape(-1, 2)
    ok = (ck[:, 0] > 0) & (ck[:, 0] < W) & (ck[:, 1] > 0) & (ck[:, 1] < H)
    mask[ck[ok, 1].astype(np.int64), ck[ok, 0].astype(np.int64)] = 255                        # :2327-2334 (truncating index)
    lk = np.asarray(last_keys, f32).reshape(-1, 2)
    state = np.asarray(last_mp_state).reshape(-1)
    R = np.asarray(Rcw, f32).reshape(3, 3); t = np.asarray(tcw, f32).reshape(3)
    X = np.asarray(last_mp_world, f32).reshape(-1, 3)
    fx, fy, cx, cy = (f32(v) for v in K)
    # float32 throughout, products and sums rounded one by one in Eigen's order (:2352-2360)
    pc = [(R[r, 0] * X[:, 0] + R[r, 1] * X[:, 1]) + R[r, 2] * X[:, 2] + t[r] for r in range(3)]
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = (1.0 / pc[2].astype(np.float64)).astype(f32)
        u = fx * pc[0] * invz + cx
        v = fy * pc[1] * invz + cy
    inside = ~((invz < 0) | (u < f32(bounds[0])) | (u > f32(bounds[1])) | (v < f32(bounds[2])) | (v > f32(bounds[3])))
    is3d = (state == 1) & inside
    is2d = (state == 0) | ((state == 1) & ~inside)
    ids3 = np.flatnonzero(is3d)
    ids2 = list(np.flatnonzero(is2d))
    out_ids, out_pts = [], []

    def accept(ids, pts, status):
        rejected = []
        for i, good in enumerate(status):
            if not good:
                rejected.append(int(ids[i]))
                continue
            x, y = float(pts[i, 0]), float(pts[i, 1])
            if mask[int(y), int(x)] == 255:                                                   # isPointNearby: truncation
                continue
            out_ids.append(int(ids[i])); out_pts.append((x, y))
            cv2.circle(mask, (int(round(x)), int(round(y))), int(DIST_THRESHOLD), 255, cv2.FILLED)  # Point2f -> Point rounds
        return rejected

    def fcheck(kps, pts, status, thr):
        idx = np.flatnonzero(status)
        if len(idx) > 8:
            st = fundamental(kps[idx], pts[idx], thr)
            if st is not None:
                status[idx[:len(st)][~np.asarray(st, bool)]] = False

    if len(ids3):                                                                             # :2380-2437
        kps = lk[ids3]
        pts, status = tracker(prev_img, cur_img, kps, np.stack([u[ids3], v[ids3]], 1).astype(f32), winsize, 3, 15.0, 0.5)
        status = np.asarray(status, bool).copy()
        fcheck(kps, pts, status, F_THRESHOLD)
        ids2 += accept(ids3, pts, status)                                                     # failures join the second pass
    if len(ids2):                                                                             # :2440-2492
        ids2 = np.asarray(ids2, np.int64)
        kps = lk[ids2]
        pts, status = tracker(prev_img, cur_img, kps, kps.copy(), winsize, 6, 15.0, 0.5)
        status = np.asarray(status, bool).copy()
        fcheck(kps, pts, status, F_THRESHOLD * 0.5)
        accept(ids2, pts, status)
    return len(out_ids), np.asarray(out_ids, np.int32), np.asarray(out_pts, f32).reshape(-1, 2)


def filter_outliers(keys, has_mp, mp_world, Rcw, tcw, K, F_THRESHOLD, fundamental=None):
    """ORBmatcher::FilterOutliers (src/ORBmatcher.cc:208-246) on plain arrays: keypoints with a map point in front of the
    camera are paired with the map point's projection and filtered by OpenCV's fundamental-matrix RANSAC.  Returns
    (inliers, mvbOutlier (n,) bool).  The reference writes the status of the j-th FILTERED pair to mvbOutlier[j] (:237-240),
    not to the keypoint the pair came from; that is kept."""
    f32 = np.float32
    fundamental = fundamental or _default_fundamental
    keys = np.asarray(keys, f32).reshape(-1, 2)
    has = np.asarray(has_mp, bool).reshape(-1)
    R = np.asarray(Rcw, f32).reshape(3, 3); t = np.asarray(tcw, f32).reshape(3)
    X = np.asarray(mp_world, f32).reshape(-1, 3)
    fx, fy, cx, cy = (f32(v) for v in K)
    pc = [(R[r, 0] * X[:, 0] + R[r, 1] * X[:, 1]) + R[r, 2] * X[:, 2] + t[r] for r in range(3)]
    with np.errstate(divide="ignore", invalid="ignore"):
        invz = (1.0 / pc[2].astype(np.float64)).astype(f32)
        u = fx * pc[0] / pc[2] + cx                                                            # Pinhole::project, float32
        v = fy * pc[1] / pc[2] + cy
    sel = np.flatnonzero(has & ~(invz < 0))
    outlier = np.zeros(len(keys), bool)
    inliers = 0
    if len(sel) > 8:
        st = fundamental(np.stack([u[sel], v[sel]], 1).astype(f32), keys[sel], F_THRESHOLD)
        if st is not None:
            st = np.asarray(st, bool)
            outlier[:len(st)] = ~st
            inliers = int(st.sum())
    return inliers, outlier


def _as_kp(k):
    if isinstance(k, np.ndarray) and k.dtype == KP_DTYPE:
        return np.ascontiguousarray(k)
    a = np.asarray(k, np.float32).reshape(-1, 2)
    out = np.zeros(len(a), KP_DTYPE)
    out["x"], out["y"] = a[:, 0], a[:, 1]
    return out


PROJ_QUERY_DTYPE = np.dtype([("u", "<f4"), ("v", "<f4"), ("radius", "<f4"), ("ur", "<f4"), ("angle", "<f4"),
                             ("min_level", "<i4"), ("max_level", "<i4"), ("blocks", "<i4"), ("desc", "u1", (32,))])
assert PROJ_QUERY_DTYPE.itemsize == 64


def search_by_projection(mode, queries, kps_un, u_right, desc, occupied, grid, nnratio=0.9, check_orientation=True,
                         stream=None):
    """ORBmatcher::SearchByProjection on flattened inputs (include/gfs_b200.h GfsProjQuery).
    mode 0 = (CurrentFrame, LastFrame, th) src/ORBmatcher.cc:1853; mode 1 = (Frame, vpMapPoints, th) :43.
    grid = (mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv).
    -> (assign[n] query index per keypoint or -1, nmatches)"""
    L = _lib.lib()
    q = np.ascontiguousarray(queries, PROJ_QUERY_DTYPE)
    k = np.ascontiguousarray(kps_un, KP_DTYPE)
    ur = np.ascontiguousarray(u_right, np.float32)
    d = np.ascontiguousarray(desc, np.uint8).reshape(-1, 32)
    oc = None if occupied is None else np.ascontiguousarray(occupied, np.uint8)
    assign = np.full(max(len(k), 1), -1, np.int32)
    nm = C.c_int()
    check(L.gfs_search_by_projection(stream, int(mode), float(nnratio), int(check_orientation), ptr(q), len(q), ptr(k),
                                     ptr(ur), ptr(d), ptr(oc), len(k), float(grid[0]), float(grid[1]), float(grid[2]),
                                     float(grid[3]), ptr(assign), C.byref(nm)))
    return assign[:len(k)], nm.value


def depth_to_cloud(depth, stride, fx, fy, cx, cy, stream=None):
    """Frame::ConvertDepthToPointCloud (src/Frame.cc:590-623) -> (n, 4) float32 points (x, y, z, 1)."""
    L = _lib.lib()
    depth = np.ascontiguousarray(depth, np.float32)
    h, w = depth.shape
    cap = ((w + stride - 1) // stride) * ((h + stride - 1) // stride)
    out = np.zeros((cap, 4), np.float32)
    n = C.c_int()
    check(L.gfs_depth_to_cloud(stream, ptr(depth), w, h, int(stride), float(fx), float(fy), float(cx), float(cy), ptr(out),
                               cap, C.byref(n)))
    return out[:n.value].copy()
