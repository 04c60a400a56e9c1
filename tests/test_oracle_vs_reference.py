"""Pins the CPU oracle to THE REFERENCE'S OWN CODE where that compiles here: oracle/_ref/libgfs_ref.so is
/root/reference/src/ORBextractor.cc (+ include/ORBextractor.h) and /root/reference/Thirdparty/GMS/include/gms_matcher.h compiled
unmodified against stand-in OpenCV headers (oracle/ref_stubs/opencv2/opencv.hpp); the five OpenCV image primitives behind
them are the restatements pinned bit-exact to cv2 4.13 by tests/test_oracle_orb.py and tests/golden.  So everything the
reference itself owns on the ORB / GMS rows of SURVEY.md 8(a) -- constructor tables, cell grid, ini->min threshold retry,
DistributeOctTree / DivideNode / the unstable std::sort, IC_Angle, computeOrbDescriptor with glibc's sinf / cosf, operator()'s
mono / lapping packing, and the whole GMS filter -- is compared statement-for-statement-compiled against the restatement,
bit for bit.  The built .so is git-ignored and travels to the GPU box; the tests skip only if it is absent AND cannot be built.
"""
import numpy as np
import pytest

from geoflowslam_b200 import synth
from oracle import oracle as O
from oracle import ref as R

pytestmark = pytest.mark.skipif(not R.available(), reason="oracle/_ref/libgfs_ref.so absent and /root/reference not mounted")

FIELDS = ("x", "y", "size", "angle", "response", "octave")


def _same(a, b):
    ka, da, ma = a
    kb, db, mb = b
    assert len(ka) == len(kb) and ma == mb
    for f in FIELDS:
        assert np.array_equal(ka[f], kb[f]), f
    return int(np.unpackbits(da ^ db).sum())


def test_extractor_equals_the_reference_on_configs1_frames():
    """ORB_SLAM3::ORBextractor (compiled from the reference) vs the oracle on configs[1] frames: keypoints (position, size,
    angle, response, octave), their ORDER, monoIndex and every descriptor bit.  The reference evaluates cos / sin of the
    float angle with glibc's cosf / sinf (src/ORBextractor.cc:101-102); the oracle rounds the exact value once.  The count
    below MEASURES how many descriptor bits that changes (DESIGN.md "rBRIEF trig pin")."""
    frames = synth.orb_frames(12, 640, 480, group=4, seed0=1000)
    orc = O.OrbOracle(1000, 1.2, 8, 25, 7)
    bits = kps = 0
    for img in frames:
        r = R.orb_extract(img, 1000, 1.2, 8, 25, 7)
        o = orc.extract(img)
        bits += _same(r, o)
        kps += len(r[0])
    assert kps > 11000
    assert bits == 0, "%d descriptor bits differ over %d keypoints (glibc sinf/cosf vs correctly rounded)" % (bits, kps)


@pytest.mark.parametrize("cfg", [dict(nfeatures=500, scale=1.2, nlevels=8, ini_th=20, min_th=7),
                                 dict(nfeatures=1500, scale=1.1, nlevels=5, ini_th=12, min_th=5),
                                 dict(nfeatures=300, scale=1.5, nlevels=4, ini_th=40, min_th=10)])
def test_extractor_equals_the_reference_other_settings(cfg):
    frames = synth.orb_frames(2, 640, 480, group=2, seed0=1100 + cfg["nfeatures"])
    orc = O.OrbOracle(cfg["nfeatures"], cfg["scale"], cfg["nlevels"], cfg["ini_th"], cfg["min_th"])
    for img in frames:
        assert _same(R.orb_extract(img, **cfg), orc.extract(img)) == 0


def test_extractor_edge_images_and_lapping_area():
    rng = np.random.default_rng(5)
    flat = np.full((480, 640), 128, np.uint8)                       # no corners anywhere: the minTh retry finds nothing either
    noise = rng.integers(0, 256, (480, 640), dtype=np.uint8)        # corners everywhere: the quadtree saturates
    small = synth.orb_frames(1, 640, 480, seed0=1300)[0][:200, :260].copy()   # the coarse levels are only a few cells wide
    low = (synth.orb_frames(1, 640, 480, seed0=1301)[0] // 8 + 100).astype(np.uint8)   # low contrast: cells fall back to minTh
    orc = O.OrbOracle(1000, 1.2, 8, 25, 7)
    for img in (flat, noise, small, low):
        r, o = R.orb_extract(img, 1000, 1.2, 8, 25, 7), orc.extract(img)
        assert _same(r, o) == 0
    assert len(R.orb_extract(flat, 1000, 1.2, 8, 25, 7)[0]) == 0
    img = synth.orb_frames(1, 640, 480, seed0=1302)[0]
    r, o = R.orb_extract(img, 1000, 1.2, 8, 25, 7, lapping=(200, 400)), orc.extract(img, lapping=(200, 400))
    assert _same(r, o) == 0 and 0 < r[2] < len(r[0])                # keypoints inside the lapping area are packed from the back


def test_pyramid_levels_equal_the_reference():
    img = synth.orb_frames(1, 640, 480, seed0=1400)[0]
    orc = O.OrbOracle(1000, 1.2, 8, 25, 7)
    orc.extract(img)
    for l in range(8):
        assert np.array_equal(R.pyramid_level(img, l), orc.level(l))


def _match_problem(seed, n=1000, w=640, h=480, inlier_frac=0.6):
    rng = np.random.default_rng(seed)
    p1 = rng.uniform([0, 0], [w - 1, h - 1], (n, 2)).astype(np.float32)
    shift = rng.uniform(-12, 12, 2)
    p2 = (p1 + shift + rng.normal(0, 0.7, (n, 2))).astype(np.float32)
    p2 = np.clip(p2, 0, [w - 1, h - 1]).astype(np.float32)
    tr = np.arange(n)
    bad = rng.random(n) > inlier_frac
    tr[bad] = rng.integers(0, n, bad.sum())
    return p1, p2, np.stack([np.arange(n), tr], 1).astype(np.int32)


@pytest.mark.parametrize("seed,n,frac", [(0, 1000, 0.6), (1, 1000, 0.2), (2, 300, 0.9), (3, 40, 0.5), (4, 2000, 0.7)])
def test_gms_equals_the_reference(seed, n, frac):
    """gms_matcher::GetInlierMask(false, false) -- what SearchWithGMS calls (src/ORBmatcher.cc:767-768) -- compiled from the
    reference vs the oracle: the mask and the count."""
    p1, p2, m = _match_problem(seed, n, inlier_frac=frac)
    rm, rn = R.gms(p1, (640, 480), p2, (640, 480), m)
    om, on = O.gms_filter(p1, (640, 480), p2, (640, 480), m)
    assert rn == on and np.array_equal(rm, om)
    if frac >= 0.5 and n >= 300:
        assert rn > 0.3 * n * frac


def test_gms_degenerate_inputs_equal_the_reference():
    p1, p2, m = _match_problem(9, 500)
    for pts1, pts2, mm in ((p1, p2, m[:1]), (p1, p2, m[:9]), (p1[:1].repeat(500, 0), p2, m), (p1, p2[:1].repeat(500, 0), m)):
        rm, rn = R.gms(pts1, (640, 480), pts2, (640, 480), mm)
        om, on = O.gms_filter(pts1, (640, 480), pts2, (640, 480), mm)
        assert rn == on and np.array_equal(rm, om)
    # points on the right / bottom image edge: x == width maps to cell 20 and is rejected by GetGridIndexLeft (:165)
    q1 = p1.copy(); q1[::7, 0] = 639.99
    rm, rn = R.gms(q1, (640, 480), p2, (640, 480), m)
    om, on = O.gms_filter(q1, (640, 480), p2, (640, 480), m)
    assert rn == on and np.array_equal(rm, om)


def test_gms_on_real_orb_matches_equals_the_reference():
    """The configs[1] chain: oracle ORB on two frames of one scene -> BF-Hamming -> GMS, reference-compiled vs oracle."""
    f = synth.orb_frames(2, 640, 480, group=2, seed0=1500)
    orc = O.OrbOracle(1000, 1.2, 8, 25, 7)
    (k1, d1, _), (k2, d2, _) = orc.extract(f[0]), orc.extract(f[1])
    idx, _ = O.bf_match(d1, d2)
    m = np.stack([np.arange(len(idx), dtype=np.int32), idx], 1)
    a, b = np.stack([k1["x"], k1["y"]], 1), np.stack([k2["x"], k2["y"]], 1)
    rm, rn = R.gms(a, (640, 480), b, (640, 480), m)
    om, on = O.gms_filter(a, (640, 480), b, (640, 480), m)
    assert rn == on and np.array_equal(rm, om) and rn > 200
