"""Batched tracking front-end over the C ABI: Frame::ExtractORB (reference src/Frame.cc:768-777)
followed by ORBmatcher::SearchWithGMS (src/ORBmatcher.cc:744-778) from frame i to frame i+1, for a
batch of independent frames per call (BASELINE.json configs[1])."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import KP_DTYPE, check, ptr

STAGES = ("pyramid", "fast_cells", "octree", "blur", "orient_desc", "pack_lapping", "bf_hamming", "gms")


class TrackingFrontend:
    def __init__(self, nfeatures=1000, scaleFactor=1.2, nlevels=8, iniThFAST=25, minThFAST=7, max_size=(640, 480),
                 max_batch=1):
        self._L = _lib.lib()
        _lib.require_device()
        h = C.c_void_p()
        check(self._L.gfs_frontend_create(int(nfeatures), float(scaleFactor), int(nlevels), int(iniThFAST),
                                          int(minThFAST), int(max_size[0]), int(max_size[1]), int(max_batch),
                                          C.byref(h)))
        self._h = h
        self.max_batch = int(max_batch)
        self.stride = self._L.gfs_frontend_max_keypoints(self._h)

    def close(self):
        if getattr(self, "_h", None) is not None and self._h.value:
            self._L.gfs_frontend_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def launches_per_call(self, batch):
        return self._L.gfs_frontend_launches_per_call(self._h, int(batch))

    def set_profiling(self, on=True):
        check(self._L.gfs_frontend_set_profiling(self._h, int(on)))

    def profile(self):
        ms = np.zeros(8, np.float32)
        check(self._L.gfs_frontend_get_profile(self._h, ptr(ms)))
        return dict(zip(STAGES, ms.tolist()))

    def alloc_host_outputs(self, batch, pinned_alloc=None):
        """Output arrays for run(); pinned_alloc(shape, dtype) may return pinned numpy views."""
        mk = pinned_alloc or (lambda shape, dt: np.zeros(shape, dt))
        s, m = self.stride, max(batch - 1, 1)
        return dict(kp=mk((batch, s), KP_DTYPE), desc=mk((batch, s, 32), np.uint8), n=mk((batch,), np.int32),
                    mono=mk((batch,), np.int32), train_idx=mk((m, s), np.int32), dist=mk((m, s), np.int32),
                    inlier=mk((m, s), np.uint8), inlier_count=mk((m,), np.int32))

    def run(self, images, out=None, stream=None):
        """images: (B,H,W) uint8 host array.  H2D + kernels + D2H; returns the dict of host arrays."""
        assert images.ndim == 3 and images.dtype == np.uint8 and images.flags["C_CONTIGUOUS"]
        B, h, w = images.shape
        out = out or self.alloc_host_outputs(B)
        check(self._L.gfs_frontend_run(self._h, stream, ptr(images), B, w, h, w, w * h, ptr(out["kp"]),
                                       ptr(out["desc"]), ptr(out["n"]), ptr(out["mono"]), ptr(out["train_idx"]),
                                       ptr(out["dist"]), ptr(out["inlier"]), ptr(out["inlier_count"])))
        return out

    def run_device(self, d_imgs, batch, w, h, pitch, img_stride, d_out, stream=None):
        """d_imgs and d_out[...] are device buffers (torch tensors or int addresses); asynchronous."""
        check(self._L.gfs_frontend_run_device(self._h, stream, ptr(d_imgs), batch, w, h, pitch, img_stride,
                                              ptr(d_out["kp"]), ptr(d_out["desc"]), ptr(d_out["n"]),
                                              ptr(d_out["mono"]), ptr(d_out["train_idx"]), ptr(d_out["dist"]),
                                              ptr(d_out["inlier"]), ptr(d_out["inlier_count"])))
