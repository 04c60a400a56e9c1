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
