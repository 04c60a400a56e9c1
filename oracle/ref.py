"""ctypes loader for oracle/_ref/libgfs_ref.so: the REFERENCE'S OWN code -- src/ORBextractor.cc and
Thirdparty/GMS/include/gms_matcher.h, compiled unmodified from /root/reference against the stand-in OpenCV headers in
oracle/ref_stubs (see ref_stubs/opencv2/opencv.hpp for what is stubbed and how it is pinned).

TEST INFRASTRUCTURE ONLY.  Built here by `make -C oracle ref` (needs /root/reference); on the GPU box only the prebuilt .so
exists.  available() says whether it can be used; tests that need it skip otherwise.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libgfs_ref.so")
REFERENCE = "/root/reference"
_LIB = None


def build(force=False):
    """(Re)build _ref/libgfs_ref.so when the reference tree is present; returns the path or None."""
    from . import oracle as O
    O.build()
    if os.path.isdir(os.path.join(REFERENCE, "src")):
        deps = [os.path.join(_HERE, "ref_stubs", f) for f in ("cv_stub.cpp", "ref_glue.cpp", "opencv2/opencv.hpp")]
        stale = (not os.path.exists(_SO)) or any(os.path.getmtime(d) > os.path.getmtime(_SO) for d in deps)
        if force or stale:
            subprocess.check_call(["make", "-C", _HERE, "-s", "ref"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return _SO if os.path.exists(_SO) else None


def available():
    try:
        return lib() is not None
    except OSError:
        return False


def lib():
    global _LIB
    if _LIB is None:
        so = build()
        if so is None:
            return None
        from . import oracle as O
        O.lib()                      # libgfs_oracle.so holds the cv2-pinned primitives the stubs forward to
        L = C.CDLL(so)
        vp, ci, cf = C.c_void_p, C.c_int, C.c_float
        L.ref_orb_extract.restype = ci
        L.ref_orb_extract.argtypes = [vp, ci, ci, ci, cf, ci, ci, ci, ci, ci, vp, vp, ci, vp]
        L.ref_orb_pyramid_level.restype = ci
        L.ref_orb_pyramid_level.argtypes = [vp, ci, ci, cf, ci, ci, vp, ci]
        L.ref_gms.restype = ci
        L.ref_gms.argtypes = [vp, ci, ci, ci, vp, ci, ci, ci, vp, ci, ci, ci, vp]
        _LIB = L
    return _LIB


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def orb_extract(img, nfeatures=1000, scale=1.2, nlevels=8, ini_th=20, min_th=7, lapping=(0, 0)):
    """ORB_SLAM3::ORBextractor(...)(img, noArray(), kps, desc, vLappingArea) -> (kps structured array, desc (n,32) u8, monoIndex)"""
    from .oracle import KP_DTYPE
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    cap = 4 * nfeatures + 64
    kp = np.zeros((cap, 6), np.float32); desc = np.zeros((cap, 32), np.uint8); mono = C.c_int()
    n = lib().ref_orb_extract(_p(img), w, h, nfeatures, scale, nlevels, ini_th, min_th, int(lapping[0]), int(lapping[1]), _p(kp), _p(desc),
                              cap, C.byref(mono))
    assert n <= cap
    out = np.zeros(n, KP_DTYPE)
    for j, f in enumerate(("x", "y", "size", "angle", "response")):
        out[f] = kp[:n, j]
    out["octave"] = kp[:n, 5].astype(np.int32)
    return out, desc[:n].copy(), mono.value


def pyramid_level(img, level, scale=1.2, nlevels=8):
    img = np.ascontiguousarray(img, np.uint8)
    h, w = img.shape
    buf = np.zeros(h * w, np.uint8)
    r = lib().ref_orb_pyramid_level(_p(img), w, h, scale, nlevels, level, _p(buf), buf.size)
    lw, lh = r & 0xffff, r >> 16
    return buf[:lw * lh].reshape(lh, lw).copy()


def gms(pts1, size1, pts2, size2, matches, with_scale=False, with_rotation=False):
    """gms_matcher(kp1, size1, kp2, size2, matches).GetInlierMask(mask, with_scale, with_rotation) -> (mask (nm,) bool, count)"""
    p1 = np.ascontiguousarray(pts1, np.float32).reshape(-1, 2); p2 = np.ascontiguousarray(pts2, np.float32).reshape(-1, 2)
    m = np.ascontiguousarray(matches, np.int32).reshape(-1, 2)
    mask = np.zeros(max(len(m), 1), np.uint8)
    n = lib().ref_gms(_p(p1), len(p1), int(size1[0]), int(size1[1]), _p(p2), len(p2), int(size2[0]), int(size2[1]), _p(m), len(m),
                      int(with_scale), int(with_rotation), _p(mask))
    return mask[:len(m)].astype(bool), n
