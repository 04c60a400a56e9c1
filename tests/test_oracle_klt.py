"""CPU oracle of the optical-flow front end (oracle/klt_oracle.cpp) pinned against the cv2 4.13 wheel, the OpenCV the
reference's calls resolve to here (cv::buildOpticalFlowPyramid reference src/Frame.cc:373, cv::calcOpticalFlowPyrLK
src/ORBmatcher.cc:2224,2271): pyramid and Scharr derivatives bit-exact; tracked positions within 5e-3 px (OpenCV sums
the integer products in float SIMD lanes, the oracle exactly in int64), status flags equal."""
import cv2
import numpy as np
import pytest

from geoflowslam_b200 import synth
from oracle import oracle as O

CRIT = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
FLAGS = cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS


@pytest.fixture(scope="module")
def pair():
    f = synth.orb_frames(2, 640, 480, group=8, seed0=1000)
    pts = cv2.goodFeaturesToTrack(f[0], 600, 0.01, 7).reshape(-1, 2).astype(np.float32)
    return f[0], f[1], pts


@pytest.mark.parametrize("shape", [(480, 640), (333, 445), (61, 35), (2, 3)])
def test_pyr_down_and_scharr_bit_exact(shape):
    rng = np.random.default_rng(shape[0])
    img = rng.integers(0, 256, shape, dtype=np.uint8)
    assert np.array_equal(O.pyr_down(img), cv2.pyrDown(img))
    if min(shape) >= 3:
        _, pyr = cv2.buildOpticalFlowPyramid(img, (5, 5), 0)
        assert np.array_equal(O.scharr_deriv(img), pyr[1])


def test_pyramid_matches_build_optical_flow_pyramid(pair):
    a, _, _ = pair
    n, cv_pyr = cv2.buildOpticalFlowPyramid(a, (35, 35), 3)
    assert n == 3
    mine = O.klt_unpack(O.klt_build_pyramid(a, 3), 640, 480, 3)
    for (img, der), ci, cd in zip(mine, cv_pyr[0::2], cv_pyr[1::2]):
        assert np.array_equal(img, ci) and np.array_equal(der, cd)


@pytest.mark.parametrize("max_level", [3, 1, 0])
def test_lk_matches_cv2(pair, max_level):
    a, b, pts = pair
    pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
    init = pts + np.float32(0.7)
    cn, cs, ce = cv2.calcOpticalFlowPyrLK(a, b, pts, init.copy(), winSize=(35, 35), maxLevel=max_level, criteria=CRIT, flags=FLAGS)
    on, os_, oe = O.klt_calc(pa, pb, 640, 480, 3, pts, init=init, win=35, max_level=max_level)
    cs = cs.ravel().astype(bool)
    assert (cs == os_.astype(bool)).mean() >= 0.995
    both = cs & os_.astype(bool)
    d = np.linalg.norm(on - cn, axis=1)[both]
    assert np.quantile(d, 0.99) < 5e-3 and np.median(d) < 1e-4, (np.median(d), d.max())
    assert np.allclose(oe[both], ce.ravel()[both], rtol=1e-4, atol=1e-7)


def test_lk_points_at_and_beyond_the_border(pair):
    a, b, _ = pair
    pts = np.array([[0.0, 0.0], [639.0, 479.0], [3.2, 240.7], [636.9, 10.1], [320.5, 1.5], [-40.0, 100.0], [700.0, 500.0], [17.0, 462.0]], np.float32)
    pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
    for lvl in (3, 0):
        cn, cs, ce = cv2.calcOpticalFlowPyrLK(a, b, pts, pts.copy(), winSize=(35, 35), maxLevel=lvl, criteria=CRIT, flags=FLAGS)
        on, os_, oe = O.klt_calc(pa, pb, 640, 480, 3, pts, init=pts, win=35, max_level=lvl)
        assert np.array_equal(cs.ravel().astype(bool), os_.astype(bool))
        ok = os_.astype(bool)
        assert np.abs(on[ok] - cn[ok]).max() < 2e-2


def _fb_klt_cv2(a, b, kps, priors, win=35, nlvl=3, ferr=15.0, maxd=0.5):
    """ORBmatcher::fbKltTracking written with the cv2 calls the reference makes"""
    pr, st, er = cv2.calcOpticalFlowPyrLK(a, b, kps, priors.copy(), winSize=(win, win), maxLevel=nlvl, criteria=CRIT, flags=FLAGS)
    st = st.ravel().astype(bool); er = er.ravel()
    h, w = a.shape
    good = st & ~(er > ferr) & (pr[:, 0] >= 1) & (pr[:, 0] < w - 1) & (pr[:, 1] >= 1) & (pr[:, 1] < h - 1)
    idx = np.nonzero(good)[0]
    status = good.copy()
    if len(idx):
        back, st2, _ = cv2.calcOpticalFlowPyrLK(b, a, pr[idx], kps[idx].copy(), winSize=(win, win), maxLevel=0, criteria=CRIT, flags=FLAGS)
        st2 = st2.ravel().astype(bool)
        dist = np.linalg.norm(kps[idx].astype(np.float64) - back.astype(np.float64), axis=1)
        status[idx] = st2 & ~(dist > maxd)
    return pr, status


def test_fb_klt_tracking_matches_cv2_composition(pair):
    a, b, pts = pair
    pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
    pr_o, st_o = O.fb_klt_tracking(pa, pb, 640, 480, 3, pts, pts)
    pr_c, st_c = _fb_klt_cv2(a, b, pts, pts)
    assert (st_o == st_c).mean() >= 0.99 and st_o.sum() > 0.5 * len(pts)
    both = st_o & st_c
    assert np.quantile(np.linalg.norm(pr_o - pr_c, axis=1)[both], 0.99) < 5e-3
    # empty input: untouched outputs
    pr_e, st_e = O.fb_klt_tracking(pa, pb, 640, 480, 3, np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert len(pr_e) == 0 and len(st_e) == 0


@pytest.mark.parametrize("shape,clip,tiles", [((480, 640), 3.0, (8, 8)), ((333, 445), 3.0, (8, 8)), ((100, 120), 40.0, (4, 6)), ((64, 64), 3.0, (8, 8)),
                                              ((480, 640), 0.0, (8, 8))])
def test_clahe_bit_exact_vs_cv2(shape, clip, tiles):
    rng = np.random.default_rng(shape[1])
    imgs = [rng.integers(0, 256, shape, dtype=np.uint8), (rng.integers(0, 40, shape) + 100).astype(np.uint8), np.full(shape, 7, np.uint8)]
    if shape == (480, 640):
        imgs.append(synth.orb_frames(1, 640, 480, group=8, seed0=1000)[0])
    c = cv2.createCLAHE(clip, tiles)
    for img in imgs:
        assert np.array_equal(O.clahe(img, clip, tiles), c.apply(img))


def test_qvga_pyramid_stops_early_like_opencv():
    """cv::buildOpticalFlowPyramid stops when the next level would be <= the window (QVGA, 35 px: level 2 is the last) and
    calcOpticalFlowPyrLK clamps maxLevel to that; the oracle tracks from the same level."""
    f = synth.orb_frames(2, 640, 480, group=8, seed0=1000)
    a, b = cv2.resize(f[0], (320, 240), interpolation=cv2.INTER_AREA), cv2.resize(f[1], (320, 240), interpolation=cv2.INTER_AREA)
    n, _ = cv2.buildOpticalFlowPyramid(a, (35, 35), 3)
    assert n == 2
    pts = cv2.goodFeaturesToTrack(a, 300, 0.01, 7).reshape(-1, 2).astype(np.float32)
    init = pts + np.float32(0.6)
    pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
    cn, cs, _ = cv2.calcOpticalFlowPyrLK(a, b, pts, init.copy(), winSize=(35, 35), maxLevel=3, criteria=CRIT, flags=FLAGS)
    on, os_, _ = O.klt_calc(pa, pb, 320, 240, 3, pts, init=init, win=35, max_level=3)
    o2, s2, _ = O.klt_calc(pa, pb, 320, 240, 3, pts, init=init, win=35, max_level=2)
    assert np.array_equal(on, o2) and np.array_equal(os_, s2)            # level 3 is not used
    cs = cs.ravel().astype(bool)
    assert (cs == os_.astype(bool)).mean() >= 0.99
    both = cs & os_.astype(bool)
    d = np.linalg.norm(on - cn, axis=1)[both]
    assert np.quantile(d, 0.99) < 5e-3 and np.median(d) < 1e-4, (np.median(d), d.max())
