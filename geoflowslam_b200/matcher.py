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
