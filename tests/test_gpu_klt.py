"""GPU parity of the optical-flow front end (through the C ABI) vs the CPU oracle and vs OpenCV itself.
Pyramid and Scharr derivatives: bit-exact vs the oracle AND vs cv2.buildOpticalFlowPyramid.  Tracker: bit-exact vs the
oracle (both sum the integer products exactly); vs cv2.calcOpticalFlowPyrLK positions within 5e-3 px, status equal."""
import numpy as np
import pytest

from geoflowslam_b200 import synth

pytestmark = pytest.mark.gpu


def _frames(n=2, seed0=1000):
    return synth.orb_frames(n, 640, 480, group=8, seed0=seed0)


def _points(img, n=600):
    import cv2
    return cv2.goodFeaturesToTrack(img, n, 0.01, 7).reshape(-1, 2).astype(np.float32)


def _pyramids(trk, frames):
    import torch
    B, h, w = frames.shape
    d = torch.from_numpy(frames).cuda()
    pyr = torch.zeros((B, trk.pyramid_bytes(w, h)), dtype=torch.uint8, device="cuda")
    trk.build_pyramids_device(d, B, w, h, w, w * h, pyr)
    torch.cuda.synchronize()
    return pyr


def test_pyramid_bit_exact_vs_oracle_and_cv2():
    import cv2
    from geoflowslam_b200 import KltTracker
    from oracle import oracle as O
    for shape, seed in (((480, 640), 1), ((333, 445), 2), ((61, 35), 3)):
        rng = np.random.default_rng(seed)
        frames = rng.integers(0, 256, (3,) + shape, dtype=np.uint8)
        h, w = shape
        trk = KltTracker(max_size=(w, h), levels=3, max_batch=3)
        pyr = _pyramids(trk, frames).cpu().numpy()
        for f in range(3):
            got = trk.unpack_pyramid(pyr[f], w, h)
            want = O.klt_unpack(O.klt_build_pyramid(frames[f], 3), w, h, 3)
            _, cvp = cv2.buildOpticalFlowPyramid(frames[f], (35, 35), 3)
            for (gi, gd), (oi, od), ci, cd in zip(got, want, cvp[0::2], cvp[1::2]):
                assert np.array_equal(gi, oi) and np.array_equal(gd, od)
                assert np.array_equal(gi, ci) and np.array_equal(gd, cd)


@pytest.mark.parametrize("max_level", [3, 0])
def test_calc_matches_oracle_bitwise_and_cv2(max_level):
    import cv2
    import torch
    from geoflowslam_b200 import KltTracker
    from oracle import oracle as O
    fr = _frames()
    pts = _points(fr[0])
    border = np.array([[0.0, 0.0], [639.0, 479.0], [3.2, 240.7], [636.9, 10.1], [320.5, 1.5], [-40.0, 100.0], [700.0, 500.0]], np.float32)
    pts = np.concatenate([pts, border])
    init = pts + np.float32(0.7)
    n = len(pts)
    trk = KltTracker(max_points=1024, max_batch=2)
    pyr = _pyramids(trk, fr)
    d_pts = torch.from_numpy(pts).cuda(); d_next = torch.from_numpy(init.copy()).cuda()
    d_n = torch.tensor([n], dtype=torch.int32, device="cuda")
    d_st = torch.zeros(n, dtype=torch.uint8, device="cuda"); d_er = torch.zeros(n, dtype=torch.float32, device="cuda")
    trk.calc_device(pyr[0], pyr[1], 1, 640, 480, d_pts, d_next, d_n, n, d_st, d_er, win=35, max_level=max_level)
    torch.cuda.synchronize()
    g_next, g_st, g_er = d_next.cpu().numpy(), d_st.cpu().numpy(), d_er.cpu().numpy()
    pa, pb = O.klt_build_pyramid(fr[0], 3), O.klt_build_pyramid(fr[1], 3)
    o_next, o_st, o_er = O.klt_calc(pa, pb, 640, 480, 3, pts, init=init, win=35, max_level=max_level)
    assert np.array_equal(g_st, o_st)
    assert np.array_equal(g_next, o_next) and np.array_equal(g_er, o_er)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    cn, cs, _ = cv2.calcOpticalFlowPyrLK(fr[0], fr[1], pts, init.copy(), winSize=(35, 35), maxLevel=max_level, criteria=crit,
                                         flags=cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS)
    cs = cs.ravel().astype(bool)
    assert (cs == g_st.astype(bool)).mean() >= 0.995
    both = cs & g_st.astype(bool)
    assert np.quantile(np.linalg.norm(g_next - cn, axis=1)[both], 0.99) < 5e-3


def test_fb_klt_tracking_matches_oracle_bitwise():
    from geoflowslam_b200 import KltTracker
    from oracle import oracle as O
    fr = _frames(3, seed0=1016)
    trk = KltTracker(max_points=1024, max_batch=1)
    for a, b in ((fr[0], fr[1]), (fr[1], fr[2])):
        pts = _points(a)
        pr_g, st_g = trk.fbKltTracking(a, b, pts, pts)
        pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
        pr_o, st_o = O.fb_klt_tracking(pa, pb, 640, 480, 3, pts, pts)
        assert np.array_equal(st_g, st_o) and np.array_equal(pr_g, pr_o)
        assert st_g.sum() > 0.5 * len(pts)
    assert trk.last_launches() == 13  # two pyramids (copy + 3 pyrDown + 1 Scharr + 1 padded copies each) + ONE tracking launch
    pr_e, st_e = trk.fbKltTracking(fr[0], fr[1], np.zeros((0, 2), np.float32), np.zeros((0, 2), np.float32))
    assert len(pr_e) == 0 and len(st_e) == 0


def test_fb_klt_batch_is_position_independent():
    import torch
    from geoflowslam_b200 import KltTracker
    fr = _frames(8, seed0=1032)
    B = 7
    trk = KltTracker(max_points=512, max_batch=8)
    pyr = _pyramids(trk, fr)
    kps = np.zeros((B, 512, 2), np.float32); n = np.zeros(B, np.int32)
    for i in range(B):
        p = _points(fr[i], 500)
        kps[i, :len(p)] = p; n[i] = len(p)
    d_k = torch.from_numpy(kps).cuda(); d_p = d_k.clone(); d_n = torch.from_numpy(n).cuda()
    d_s = torch.zeros((B, 512), dtype=torch.uint8, device="cuda")
    trk.fb_track_device(pyr[:B].contiguous(), pyr[1:B + 1].contiguous(), B, 640, 480, d_k, d_p, d_n, 512, d_s)
    torch.cuda.synchronize()
    one = KltTracker(max_points=512, max_batch=1)
    for i in range(B):
        pr, st = one.fbKltTracking(fr[i], fr[i + 1], kps[i, :n[i]], kps[i, :n[i]])
        assert np.array_equal(d_p[i, :n[i]].cpu().numpy(), pr) and np.array_equal(d_s[i, :n[i]].cpu().numpy().astype(bool), st)


def test_capacity_errors_are_loud():
    from geoflowslam_b200 import GfsError, KltTracker
    trk = KltTracker(max_size=(320, 240), max_points=10, max_batch=1)
    img = np.zeros((480, 640), np.uint8)
    with pytest.raises(GfsError):
        trk.fbKltTracking(img, img, np.zeros((4, 2), np.float32), np.zeros((4, 2), np.float32))
    small = np.zeros((240, 320), np.uint8)
    with pytest.raises(GfsError):
        trk.fbKltTracking(small, small, np.zeros((11, 2), np.float32), np.zeros((11, 2), np.float32))


def test_clahe_bit_exact_vs_oracle_and_cv2():
    import cv2
    import torch
    from geoflowslam_b200.klt import clahe_apply, clahe_apply_device
    from oracle import oracle as O
    for shape, clip, tiles in (((480, 640), 3.0, (8, 8)), ((333, 445), 3.0, (8, 8)), ((100, 120), 40.0, (4, 6)), ((480, 640), 0.0, (8, 8))):
        rng = np.random.default_rng(shape[1])
        for img in (rng.integers(0, 256, shape, dtype=np.uint8), (rng.integers(0, 40, shape) + 100).astype(np.uint8), np.full(shape, 7, np.uint8)):
            g = clahe_apply(img, clip, tiles)
            assert np.array_equal(g, O.clahe(img, clip, tiles))
            assert np.array_equal(g, cv2.createCLAHE(clip, tiles).apply(img))
    fr = _frames(4, seed0=1040)
    d = torch.from_numpy(fr).cuda()
    clahe_apply_device(d, 4, 640, 480, 640, 640 * 480, d)  # in place, as Frame::Frame does
    torch.cuda.synchronize()
    for i in range(4):
        assert np.array_equal(d[i].cpu().numpy(), cv2.createCLAHE(3.0, (8, 8)).apply(fr[i]))


@pytest.mark.parametrize("win,levels", [(21, 2), (15, 3), (41, 1), (7, 0)])
def test_other_window_sizes_match_oracle_bitwise(win, levels):
    """windows narrower and wider than a warp, fewer / more levels than the reference's 35 / 3"""
    import torch
    from geoflowslam_b200 import KltTracker
    from oracle import oracle as O
    fr = _frames(2, seed0=1048)
    pts = _points(fr[0], 300)
    pts = np.concatenate([pts, np.array([[0.0, 0.0], [639.0, 479.0], [2.5, 470.2], [630.0, 3.0]], np.float32)])
    n = len(pts)
    trk = KltTracker(levels=3, max_points=512, max_batch=2)
    pyr = _pyramids(trk, fr)
    d_pts = torch.from_numpy(pts).cuda(); d_next = d_pts.clone()
    d_n = torch.tensor([n], dtype=torch.int32, device="cuda")
    d_st = torch.zeros(n, dtype=torch.uint8, device="cuda"); d_er = torch.zeros(n, dtype=torch.float32, device="cuda")
    trk.calc_device(pyr[0], pyr[1], 1, 640, 480, d_pts, d_next, d_n, n, d_st, d_er, win=win, max_level=levels)
    torch.cuda.synchronize()
    pa, pb = O.klt_build_pyramid(fr[0], 3), O.klt_build_pyramid(fr[1], 3)
    o_next, o_st, o_er = O.klt_calc(pa, pb, 640, 480, 3, pts, init=pts, win=win, max_level=levels)
    assert np.array_equal(d_st.cpu().numpy(), o_st)
    assert np.array_equal(d_next.cpu().numpy(), o_next) and np.array_equal(d_er.cpu().numpy(), o_er)


def test_qvga_tracks_from_the_level_opencv_stops_at():
    """QVGA with the 35-px window: cv::buildOpticalFlowPyramid keeps levels 0-2 only; the CUDA tracker starts there too."""
    import cv2
    from geoflowslam_b200 import KltTracker
    from oracle import oracle as O
    fr = _frames()
    a = cv2.resize(fr[0], (320, 240), interpolation=cv2.INTER_AREA); b = cv2.resize(fr[1], (320, 240), interpolation=cv2.INTER_AREA)
    pts = cv2.goodFeaturesToTrack(a, 300, 0.01, 7).reshape(-1, 2).astype(np.float32)
    trk = KltTracker(max_size=(320, 240), max_points=512, max_batch=1)
    pr_g, st_g = trk.fbKltTracking(a, b, pts, pts)
    pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
    pr_o, st_o = O.fb_klt_tracking(pa, pb, 320, 240, 3, pts, pts)
    assert np.array_equal(st_g, st_o) and np.array_equal(pr_g, pr_o)
    crit = (cv2.TERM_CRITERIA_COUNT + cv2.TERM_CRITERIA_EPS, 30, 0.01)
    cn, cs, ce = cv2.calcOpticalFlowPyrLK(a, b, pts, pts.copy(), winSize=(35, 35), maxLevel=3, criteria=crit,
                                          flags=cv2.OPTFLOW_USE_INITIAL_FLOW + cv2.OPTFLOW_LK_GET_MIN_EIGENVALS)
    ok = st_g.astype(bool) & cs.ravel().astype(bool)
    assert ok.sum() > 0.5 * len(pts) and np.quantile(np.linalg.norm(pr_g - cn, axis=1)[ok], 0.99) < 5e-3
