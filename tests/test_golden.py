"""Golden vectors (tests/golden/*.npz, written by scripts/make_golden.py from the cv2 wheel: the OpenCV calls the
reference makes on this path).  CPU: the oracle reproduces them; GPU (marked): the CUDA path reproduces them through the
C ABI.  Integer / byte stages bit-exact; optical-flow tracks within 5e-3 px with equal status."""
import os

import numpy as np
import pytest

from oracle import oracle as O

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _load(name):
    return np.load(os.path.join(G, name))


def test_oracle_orb_primitives():
    g = _load("orb_primitives.npz")
    img = g["img"]
    assert np.array_equal(O.resize_area(img, 133, 100), g["resize_area"])
    assert np.array_equal(O.blur7(img), g["blur7"])
    for thr in (25, 7):
        want = g["fast%d" % thr]
        got = O.fast(img, thr)
        assert len(got) == len(want) > 0
        assert np.array_equal(np.asarray(got, np.float32).reshape(-1, 3), want)


def test_oracle_bf_hamming():
    g = _load("bf_hamming.npz")
    idx, dist = O.bf_match(g["dq"], g["dt"])
    assert np.array_equal(idx, g["train_idx"]) and np.array_equal(dist, g["dist"])


def test_oracle_optical_flow():
    g = _load("optical_flow.npz")
    a, b = g["prev"], g["cur"]
    h, w = a.shape
    assert np.array_equal(O.clahe(a), g["clahe"])
    lv = O.klt_unpack(O.klt_build_pyramid(a, 2), w, h, 2)
    assert np.array_equal(lv[1][0], g["pyr_img1"]) and np.array_equal(lv[2][0], g["pyr_img2"])
    for l in range(3):
        assert np.array_equal(lv[l][1], g["der%d" % l])
    pa, pb = O.klt_build_pyramid(a, 2), O.klt_build_pyramid(b, 2)
    nxt, st, er = O.klt_calc(pa, pb, w, h, 2, g["pts"], init=g["pts"], win=21, max_level=2)
    assert np.array_equal(st, g["status"])
    ok = st.astype(bool)
    assert ok.sum() > 60 and np.abs(nxt[ok] - g["next"][ok]).max() < 5e-3
    assert np.allclose(er[ok], g["min_eig"][ok], rtol=1e-4, atol=1e-7)
    # the warp between the two frames is known: the flow must point the right way
    assert np.median(nxt[ok, 0] - g["pts"][ok, 0]) > 1.0 and np.median(nxt[ok, 1] - g["pts"][ok, 1]) < -0.5


@pytest.mark.gpu
def test_gpu_bf_hamming_golden():
    from geoflowslam_b200 import ORBmatcher
    g = _load("bf_hamming.npz")
    idx, dist = ORBmatcher.bf_match(g["dq"], g["dt"])
    assert np.array_equal(idx, g["train_idx"]) and np.array_equal(dist, g["dist"])


@pytest.mark.gpu
def test_gpu_optical_flow_golden():
    import torch
    from geoflowslam_b200 import KltTracker
    from geoflowslam_b200.klt import clahe_apply
    g = _load("optical_flow.npz")
    a, b = g["prev"], g["cur"]
    h, w = a.shape
    assert np.array_equal(clahe_apply(a), g["clahe"])
    trk = KltTracker(max_size=(w, h), levels=2, max_points=256, max_batch=2)
    d = torch.from_numpy(np.stack([a, b])).cuda()
    pyr = torch.zeros((2, trk.pyramid_bytes(w, h)), dtype=torch.uint8, device="cuda")
    trk.build_pyramids_device(d, 2, w, h, w, w * h, pyr)
    lv = trk.unpack_pyramid(pyr[0].cpu().numpy(), w, h)
    assert np.array_equal(lv[1][0], g["pyr_img1"]) and np.array_equal(lv[2][0], g["pyr_img2"])
    for l in range(3):
        assert np.array_equal(lv[l][1], g["der%d" % l])
    pts = g["pts"]
    n = len(pts)
    d_p = torch.from_numpy(pts).cuda(); d_n = d_p.clone()
    d_cnt = torch.tensor([n], dtype=torch.int32, device="cuda")
    d_st = torch.zeros(n, dtype=torch.uint8, device="cuda"); d_er = torch.zeros(n, dtype=torch.float32, device="cuda")
    trk.calc_device(pyr[0], pyr[1], 1, w, h, d_p, d_n, d_cnt, n, d_st, d_er, win=21, max_level=2)
    torch.cuda.synchronize()
    st = d_st.cpu().numpy()
    assert np.array_equal(st, g["status"])
    ok = st.astype(bool)
    assert np.abs(d_n.cpu().numpy()[ok] - g["next"][ok]).max() < 5e-3
    assert np.allclose(d_er.cpu().numpy()[ok], g["min_eig"][ok], rtol=1e-4, atol=1e-7)


@pytest.mark.gpu
def test_gpu_orb_level_golden():
    """level 1 of the extractor's pyramid for a frame whose 1.2x level has the golden size is cv2's INTER_AREA resize"""
    from geoflowslam_b200 import ORBextractor
    g = _load("orb_primitives.npz")
    img = g["img"]
    ex = ORBextractor(200, 1.2, 2, 25, 7, max_size=(160, 120), max_batch=1)
    ex.extract_batch(img[None])
    assert ex.level_size(160, 120, 1) == (133, 100)
    assert np.array_equal(ex.image_pyramid_level(1), g["resize_area"])
