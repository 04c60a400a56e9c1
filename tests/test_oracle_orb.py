"""Pins the CPU oracle's OpenCV-owned stages bit-exact against the cv2 wheel and
checks the reference-owned logic against an independent cv2/numpy restatement."""
import math

import numpy as np
import pytest

from oracle import oracle as O

cv2 = pytest.importorskip("cv2")

LEVELS = [(640, 480), (533, 400), (444, 333), (370, 278), (309, 231), (257, 193), (214, 161), (179, 134)]


def test_tables_match_survey():
    e = O.OrbOracle(1000, 1.2, 8, 25, 7)
    sf, npl, um = e.tables()
    assert npl.tolist() == [217, 181, 151, 126, 105, 87, 73, 60]
    assert um.tolist() == [15, 15, 15, 15, 14, 14, 14, 13, 13, 12, 11, 10, 9, 8, 6, 3]
    assert [e.level_size(640, 480, l) for l in range(8)] == LEVELS
    assert sf[1] == np.float32(1.2)


def test_resize_area_matches_cv2_on_every_pyramid_step():
    rng = np.random.default_rng(0)
    for (sw, sh), (dw, dh) in zip(LEVELS[:-1], LEVELS[1:]):
        src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
        ref = cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA)
        got = O.resize_area(src, dw, dh)
        assert np.array_equal(ref, got), (sw, sh, dw, dh, int((ref != got).sum()))


@pytest.mark.parametrize("sw,sh,dw,dh", [(752, 480, 627, 400), (100, 77, 83, 64), (1280, 720, 1067, 600)])
def test_resize_area_other_sizes(sw, sh, dw, dh):
    rng = np.random.default_rng(1)
    src = rng.integers(0, 256, (sh, sw), dtype=np.uint8)
    assert np.array_equal(cv2.resize(src, (dw, dh), interpolation=cv2.INTER_AREA), O.resize_area(src, dw, dh))


def test_blur7_matches_cv2():
    rng = np.random.default_rng(2)
    for (w, h) in [(640, 480), (179, 134), (33, 21)]:
        src = rng.integers(0, 256, (h, w), dtype=np.uint8)
        ref = cv2.GaussianBlur(src, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101)
        assert np.array_equal(ref, O.blur7(src))


def test_fast_atan2_matches_cv2():
    rng = np.random.default_rng(3)
    ys = rng.integers(-60000, 60000, 5000)
    xs = rng.integers(-60000, 60000, 5000)
    for y, x in list(zip(ys, xs)) + [(0, 0), (1, 1), (-1, -1), (0, 5), (5, 0), (-5, 0), (0, -5)]:
        assert O.fast_atan2(y, x) == np.float32(cv2.fastAtan2(float(y), float(x))), (y, x)


@pytest.mark.parametrize("thr,shape,seed", [(25, (120, 160), 0), (7, (47, 41), 1), (10, (200, 200), 2), (7, (7, 7), 3),
                                            (20, (6, 40), 4)])
def test_fast_matches_cv2(thr, shape, seed, frames4):
    rng = np.random.default_rng(seed)
    # smooth-ish random image so corners and score plateaus both occur
    img = cv2.GaussianBlur(rng.integers(0, 256, shape, dtype=np.uint8), (3, 3), 0)
    det = cv2.FastFeatureDetector_create(thr, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16)
    ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(img)]
    got = [tuple(r) for r in O.fast(img, thr).tolist()]
    assert ref == got
    crop = np.ascontiguousarray(frames4[0][100:100 + shape[0] * 2, 50:50 + shape[1] * 2])
    ref = [(int(k.pt[0]), int(k.pt[1]), int(k.response)) for k in det.detect(crop)]
    assert ref == [tuple(r) for r in O.fast(crop, thr).tolist()]


def _cv2_candidates(im, ini, mn):
    """ComputeKeyPointsOctTree cell loop (ORBextractor.cc:794-851) with cv2.FAST on each cell."""
    h, w = im.shape
    minB, maxBX, maxBY = 16, w - 16, h - 16
    width, height = float(maxBX - minB), float(maxBY - minB)
    nC, nR = int(width / 35), int(height / 35)
    wC, hC = math.ceil(width / nC), math.ceil(height / nR)
    out = []
    for i in range(nR):
        iniY = minB + i * hC
        maxY = iniY + hC + 6
        if iniY >= maxBY - 3:
            continue
        maxY = min(maxY, maxBY)
        for j in range(nC):
            iniX = minB + j * wC
            maxX = iniX + wC + 6
            if iniX >= maxBX - 6:
                continue
            maxX = min(maxX, maxBX)
            roi = np.ascontiguousarray(im[iniY:maxY, iniX:maxX])
            k = cv2.FastFeatureDetector_create(ini, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(roi)
            if not k:
                k = cv2.FastFeatureDetector_create(mn, True, cv2.FAST_FEATURE_DETECTOR_TYPE_9_16).detect(roi)
            out += [(p.pt[0] + j * wC, p.pt[1] + i * hC, p.response) for p in k]
    return np.array(out, np.float32).reshape(-1, 3)


def test_pyramid_blur_and_cell_fast_against_cv2(frames4):
    e = O.OrbOracle(1000, 1.2, 8, 25, 7)
    img = frames4[1]
    kps, desc, mono = e.extract(img)
    assert mono == len(kps) and 900 <= len(kps) <= 1024
    prev = img
    for l in range(8):
        lw, lh = LEVELS[l]
        cur = prev if l == 0 else cv2.resize(prev, (lw, lh), interpolation=cv2.INTER_AREA)
        assert np.array_equal(cur, e.level(l))
        b = e.level(l, blurred=True)
        assert np.array_equal(cv2.GaussianBlur(cur, (7, 7), 2, sigmaY=2, borderType=cv2.BORDER_REFLECT_101), b)
        assert np.array_equal(_cv2_candidates(cur, 25, 7), e.candidates(l)), l
        prev = cur


def test_orientation_and_descriptor_against_numpy(frames4):
    """IC_Angle (ORBextractor.cc:71-95) and computeOrbDescriptor (:99-160) restated in numpy/python."""
    from geoflowslam_b200.pattern import ORB_PATTERN
    e = O.OrbOracle(1000, 1.2, 8, 25, 7)
    kps, desc, _ = e.extract(frames4[2])
    sf, _, um = e.tables()
    rng = np.random.default_rng(5)
    for idx in rng.choice(len(kps), 60, replace=False):
        k = kps[idx]
        l = int(k["octave"])
        x, y = (k["x"], k["y"]) if l == 0 else (k["x"] / sf[l], k["y"] / sf[l])
        cx, cy = int(round(float(x))), int(round(float(y)))
        im = e.level(l).astype(np.int64)
        m10 = m01 = 0
        for v in range(-15, 16):
            d = int(um[abs(v)])
            for u in range(-d, d + 1):
                m10 += u * im[cy + v, cx + u]
                m01 += v * im[cy + v, cx + u]
        assert np.float32(cv2.fastAtan2(float(m01), float(m10))) == k["angle"]
        bl = e.level(l, blurred=True)
        ang = np.float32(k["angle"]) * np.float32(math.pi / np.float32(180.0))
        a, b = np.float32(math.cos(float(ang))), np.float32(math.sin(float(ang)))
        bits = []
        for t in range(256):
            vals = []
            for p in range(2):
                px, py = np.float32(ORB_PATTERN[4 * t + 2 * p]), np.float32(ORB_PATTERN[4 * t + 2 * p + 1])
                r1 = np.float32(np.float32(px * b) + np.float32(py * a))
                r2 = np.float32(np.float32(px * a) - np.float32(py * b))
                ry = int(math.copysign(math.floor(abs(float(r1)) + 0.5), float(r1)))
                rx = int(math.copysign(math.floor(abs(float(r2)) + 0.5), float(r2)))
                vals.append(int(bl[cy + ry, cx + rx]))
            bits.append(1 if vals[0] < vals[1] else 0)
        d = np.packbits(np.array(bits, np.uint8), bitorder="little")
        assert np.array_equal(d, desc[idx]), idx


def test_keypoint_budget_and_order_properties(frames4):
    e = O.OrbOracle(1000, 1.2, 8, 25, 7)
    kps, desc, _ = e.extract(frames4[0])
    _, npl, _ = e.tables()
    oct_ = kps["octave"]
    assert np.all(np.diff(oct_) >= 0)  # levels are emitted in order
    for l in range(8):
        n = int((oct_ == l).sum())
        assert n <= npl[l] + 3
    assert desc.shape == (len(kps), 32)


def test_empty_image_returns_minus_one():
    e = O.OrbOracle()
    assert e.L.gfo_orb_extract(e.h, None, 0, 0, 0, 0, 0, None, None, 0, None) == -1
