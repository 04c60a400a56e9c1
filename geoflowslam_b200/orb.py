"""Host-side mirror of ORB_SLAM3::ORBextractor (reference include/ORBextractor.h:46-120) over the
C ABI.  Same constructor arguments, same getters, same operator() contract; the batched entry
points are this build's addition (independent frames are the data-parallel axis on B200)."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr


class ORBextractor:
    HARRIS_SCORE, FAST_SCORE = 0, 1

    def __init__(self, nfeatures, scaleFactor, nlevels, iniThFAST, minThFAST, max_size=(640, 480), max_batch=1):
        self._L = _lib.lib()
        _lib.require_device()
        h = C.c_void_p()
        check(self._L.gfs_orb_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST),
                                     int(minThFAST), int(max_size[0]), int(max_size[1]), int(max_batch), C.byref(h)))
        self._h = h
        self.nfeatures, self.nlevels = int(nfeatures), int(nlevels)
        self.scaleFactor = float(np.float32(scaleFactor))
        self.max_batch = int(max_batch)
        self.stride = self._L.gfs_orb_max_keypoints(self._h)
        sf = np.zeros(self.nlevels, np.float32)
        npl = np.zeros(self.nlevels, np.int32)
        check(self._L.gfs_orb_tables(self._h, ptr(sf), ptr(npl)))
        self.mvScaleFactor = sf
        self.mnFeaturesPerLevel = npl
        self.mvInvScaleFactor = (np.float32(1.0) / sf).astype(np.float32)
        self.mvLevelSigma2 = (sf * sf).astype(np.float32)
        self.mvInvLevelSigma2 = (np.float32(1.0) / self.mvLevelSigma2).astype(np.float32)
        self._wh = None

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.gfs_orb_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # --- getters of the reference class (include/ORBextractor.h:66-80)
    def GetLevels(self): return self.nlevels
    def GetScaleFactor(self): return self.scaleFactor
    def GetScaleFactors(self): return self.mvScaleFactor
    def GetInverseScaleFactors(self): return self.mvInvScaleFactor
    def GetScaleSigmaSquares(self): return self.mvLevelSigma2
    def GetInverseScaleSigmaSquares(self): return self.mvInvLevelSigma2

    def level_size(self, w, h, level):
        lw, lh = C.c_int(), C.c_int()
        check(self._L.gfs_orb_level_size(self._h, w, h, level, C.byref(lw), C.byref(lh)))
        return lw.value, lh.value

    def launches_per_call(self, lapping=(0, 0)):
        return self._L.gfs_orb_launches_per_call(self._h, int(lapping[0]), int(lapping[1]))

    # --- operator() (include/ORBextractor.h:61-64): returns (monoIndex, keypoints, descriptors)
    def __call__(self, image, mask=None, vLappingArea=(0, 0), stream=None):
        if image is None or getattr(image, "size", 0) == 0:
            return -1, np.zeros(0, KP_DTYPE), np.zeros((0, 32), np.uint8)
        image = np.ascontiguousarray(image, np.uint8)
        assert image.ndim == 2, "CV_8UC1 expected"
        h, w = image.shape
        kps = np.zeros(self.stride, KP_DTYPE)
        desc = np.zeros((self.stride, 32), np.uint8)
        n, mono = C.c_int(), C.c_int()
        check(self._L.gfs_orb_extract(self._h, stream, ptr(image), w, h, w, int(vLappingArea[0]),
                                      int(vLappingArea[1]), ptr(kps), ptr(desc), C.byref(n), C.byref(mono)))
        self._wh = (w, h)
        return mono.value, kps[:n.value].copy(), desc[:n.value].copy()

    def extract_batch(self, images, vLappingArea=(0, 0), stream=None, out=None):
        """images: (B, H, W) uint8 host array (pinned or pageable).  Returns (kps[B,stride],
        desc[B,stride,32], n[B], mono[B]) host arrays (pass `out` to reuse pinned buffers)."""
        assert images.ndim == 3 and images.dtype == np.uint8 and images.flags["C_CONTIGUOUS"]
        B, h, w = images.shape
        if out is None:
            out = (np.zeros((B, self.stride), KP_DTYPE), np.zeros((B, self.stride, 32), np.uint8),
                   np.zeros(B, np.int32), np.zeros(B, np.int32))
        kps, desc, n, mono = out
        check(self._L.gfs_orb_extract_batch(self._h, stream, ptr(images), B, w, h, w, w * h, int(vLappingArea[0]),
                                            int(vLappingArea[1]), ptr(kps), ptr(desc), ptr(n), ptr(mono)))
        self._wh = (w, h)
        return kps, desc, n, mono

    def extract_batch_device(self, d_imgs, batch, w, h, pitch, img_stride, d_kp, d_desc, d_n, d_mono,
                             vLappingArea=(0, 0), stream=None):
        """Device pointers (ints / torch tensors); asynchronous on `stream`."""
        check(self._L.gfs_orb_extract_batch_device(self._h, stream, ptr(d_imgs), batch, w, h, pitch, img_stride,
                                                   int(vLappingArea[0]), int(vLappingArea[1]), ptr(d_kp), ptr(d_desc),
                                                   ptr(d_n), ptr(d_mono)))
        self._wh = (w, h)

    # --- mvImagePyramid (public member, include/ORBextractor.h:82) of the last batch
    def image_pyramid_level(self, level, frame=0, blurred=False, stream=None):
        lw, lh = self.level_size(self._wh[0], self._wh[1], level)
        out = np.zeros((lh, lw), np.uint8)
        check(self._L.gfs_orb_get_level(self._h, stream, frame, level, int(blurred), ptr(out)))
        return out

    def fast_candidates(self, level, frame=0, stream=None):
        buf = np.zeros((1 << 18, 3), np.float32)
        n = C.c_int()
        check(self._L.gfs_orb_get_candidates(self._h, stream, frame, level, ptr(buf), buf.shape[0], C.byref(n)))
        return buf[:n.value].copy()
