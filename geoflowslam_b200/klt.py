"""Host-side mirror of the optical-flow front end over the C ABI: cv::buildOpticalFlowPyramid as Frame::Frame calls
it (reference src/Frame.cc:370-373) and ORBmatcher::fbKltTracking (include/ORBmatcher.h:56-60,
src/ORBmatcher.cc:2186-2293), batched over frames / frame pairs."""
import ctypes as C

import numpy as np

from . import _lib
from ._lib import check, ptr


class KltTracker:
    """`KltTracker(max_size, levels=3).fbKltTracking(prev_img, cur_img, kps, priors, nwinsize, nbpyrlvl, ferr, max_dist)`
    for one pair of host images; `build_pyramids_device` / `fb_track_device` / `calc_device` for resident batches."""

    def __init__(self, max_size=(640, 480), levels=3, max_points=2048, max_batch=1):
        self._L = _lib.lib()
        _lib.require_device()
        self._h = C.c_void_p()
        check(self._L.gfs_klt_create(int(max_size[0]), int(max_size[1]), int(levels), int(max_points), int(max_batch), C.byref(self._h)))
        self.levels, self.max_points, self.max_batch = int(levels), int(max_points), int(max_batch)

    def close(self):
        if getattr(self, "_h", None):
            self._L.gfs_klt_destroy(self._h)
            self._h = None

    __del__ = close

    def pyramid_bytes(self, w, h):
        return int(self._L.gfs_klt_pyramid_bytes(self._h, int(w), int(h)))

    def pyramid_layout(self, w, h):
        """-> (level widths, level heights, level pixel offsets, byte offset of the derivative block)"""
        n = self.levels + 1
        lw, lh, lo = np.zeros(n, np.int32), np.zeros(n, np.int32), np.zeros(n, np.int32)
        off = C.c_size_t()
        check(self._L.gfs_klt_pyramid_layout(self._h, int(w), int(h), ptr(lw), ptr(lh), ptr(lo), C.byref(off)))
        return lw, lh, lo, int(off.value)

    def unpack_pyramid(self, blob, w, h):
        """host copy of one frame's pyramid block -> [(image (h,w) u8, derivative (h,w,2) i16), ...]"""
        lw, lh, lo, doff = self.pyramid_layout(w, h)
        blob = np.ascontiguousarray(blob, np.uint8)
        der = blob[doff:].view(np.int16)
        return [(blob[o:o + a * b].reshape(b, a), der[2 * o:2 * (o + a * b)].reshape(b, a, 2)) for a, b, o in zip(lw, lh, lo)]

    def build_pyramids_device(self, d_imgs, batch, w, h, pitch, img_stride, d_pyr, stream=None):
        check(self._L.gfs_klt_build_pyramid_batch_device(self._h, stream, ptr(d_imgs), int(batch), int(w), int(h), int(pitch),
                                                         int(img_stride), ptr(d_pyr)))

    def fb_track_device(self, d_prev_pyr, d_cur_pyr, batch, w, h, d_kps, d_priors, d_n, stride, d_status, nwinsize=35, nbpyrlvl=3,
                        ferr=15.0, fmax_fbklt_dist=0.5, stream=None):
        check(self._L.gfs_klt_fb_track_batch_device(self._h, stream, ptr(d_prev_pyr), ptr(d_cur_pyr), int(batch), int(w), int(h),
                                                    ptr(d_kps), ptr(d_priors), ptr(d_n), int(stride), int(nwinsize), int(nbpyrlvl),
                                                    float(ferr), float(fmax_fbklt_dist), ptr(d_status)))

    def calc_device(self, d_prev_pyr, d_cur_pyr, batch, w, h, d_pts, d_next, d_n, stride, d_status, d_err=None, win=35, max_level=3,
                    max_count=30, eps=0.01, use_initial_flow=True, stream=None):
        check(self._L.gfs_klt_calc_batch_device(self._h, stream, ptr(d_prev_pyr), ptr(d_cur_pyr), int(batch), int(w), int(h), ptr(d_pts),
                                                ptr(d_next), ptr(d_n), int(stride), int(win), int(max_level), int(max_count), float(eps),
                                                int(bool(use_initial_flow)), ptr(d_status), ptr(d_err) if d_err is not None else None))

    def fbKltTracking(self, prev_img, cur_img, vkps, vpriorkps, nwinsize=35, nbpyrlvl=3, ferr=15.0, fmax_fbklt_dist=0.5, stream=None):
        """-> (vpriorkps after tracking (n,2) float32, vkpstatus (n,) bool)"""
        a = np.ascontiguousarray(prev_img, np.uint8); b = np.ascontiguousarray(cur_img, np.uint8)
        assert a.ndim == 2 and a.shape == b.shape
        kps = np.ascontiguousarray(vkps, np.float32).reshape(-1, 2)
        pr = np.ascontiguousarray(vpriorkps, np.float32).reshape(-1, 2).copy()
        assert len(pr) == len(kps)
        st = np.zeros(max(len(kps), 1), np.uint8)
        h, w = a.shape
        check(self._L.gfs_klt_fb_track(self._h, stream, ptr(a), ptr(b), w, h, w, ptr(kps), ptr(pr), len(kps), int(nwinsize), int(nbpyrlvl),
                                       float(ferr), float(fmax_fbklt_dist), ptr(st)))
        return pr, st[:len(kps)].astype(bool)

    def last_launches(self):
        return int(self._L.gfs_klt_last_launches(self._h))


def clahe_apply(img, clip_limit=3.0, tiles=(8, 8), stream=None):
    """cv::createCLAHE(clip_limit, tiles)->apply(img) on one host image (reference src/Frame.cc:366-368)"""
    L = _lib.lib()
    _lib.require_device()
    img = np.ascontiguousarray(img, np.uint8)
    assert img.ndim == 2
    h, w = img.shape
    out = np.zeros_like(img)
    check(L.gfs_clahe_apply(stream, ptr(img), w, h, w, float(clip_limit), int(tiles[0]), int(tiles[1]), ptr(out)))
    return out


def clahe_apply_device(d_src, batch, w, h, pitch, img_stride, d_dst, clip_limit=3.0, tiles=(8, 8), stream=None):
    """batched, device buffers; d_dst may be d_src (in place, as Frame::Frame does)"""
    check(_lib.lib().gfs_clahe_apply_batch_device(stream, ptr(d_src), int(batch), int(w), int(h), int(pitch), int(img_stride),
                                                  float(clip_limit), int(tiles[0]), int(tiles[1]), ptr(d_dst), int(pitch), int(img_stride)))
