"""GPU parity: CUDA ORB extractor (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ORB_CFG = (1000, 1.2, 8, 25, 7)  # g1_op_icp_lidar_indoor1.yaml:118-131


@pytest.fixture(scope="module")
def ex():
    from geoflowslam_b200 import ORBextractor
    return ORBextractor(*ORB_CFG, max_size=(752, 480), max_batch=8)


def _oracle():
    from oracle import oracle as O
    return O.OrbOracle(*ORB_CFG)


def _assert_same(kg, dg, ko, do):
    assert len(kg) == len(ko)
    for f in ("x", "y", "size", "angle", "response", "octave"):
        assert np.array_equal(kg[f], ko[f]), f
    assert np.array_equal(dg, do)


def test_single_frame_stages_and_output(ex, frames4):
    o = _oracle()
    ko, do, mo = o.extract(frames4[0])
    mono, kg, dg = ex(frames4[0])
    for l in range(8):
        assert np.array_equal(ex.image_pyramid_level(l), o.level(l)), ("pyramid", l)
        assert np.array_equal(ex.fast_candidates(l), o.candidates(l)), ("fast", l)
        assert np.array_equal(ex.image_pyramid_level(l, blurred=True), o.level(l, blurred=True)), ("blur", l)
    assert mono == mo
    _assert_same(kg, dg, ko, do)


def test_batch_matches_oracle(ex, frames4):
    kps, desc, n, mono = ex.extract_batch(frames4)
    o = _oracle()
    for i in range(len(frames4)):
        ko, do, mo = o.extract(frames4[i])
        assert n[i] == len(ko) and mono[i] == mo
        _assert_same(kps[i, :n[i]], desc[i, :n[i]], ko, do)


@pytest.mark.parametrize("kind", ["noise", "flat", "gradient", "checker", "wide", "pitch4", "pitch2"])
def test_edge_images(kind):
    from geoflowslam_b200 import ORBextractor
    rng = np.random.default_rng(7)
    if kind == "noise":
        img = rng.integers(0, 256, (480, 640), dtype=np.uint8)       # maximal candidate density
    elif kind == "flat":
        img = np.full((480, 640), 128, np.uint8)                      # no corners at all
    elif kind == "gradient":
        img = np.tile(np.arange(640, dtype=np.uint8), (480, 1))
    elif kind == "checker":
        yy, xx = np.mgrid[0:480, 0:640]
        img = (((yy // 8 + xx // 8) % 2) * 200 + 20).astype(np.uint8)  # score plateaus / ties
    elif kind == "pitch4":
        img = rng.integers(0, 256, (301, 644), dtype=np.uint8)        # level-0 rows 4- but not 16-byte aligned (word staging)
    elif kind == "pitch2":
        img = rng.integers(0, 256, (300, 642), dtype=np.uint8)        # level-0 rows unaligned (byte staging)
    else:
        img = rng.integers(0, 256, (280, 752), dtype=np.uint8)        # several initial quadtree nodes
    h, w = img.shape
    e = ORBextractor(*ORB_CFG, max_size=(w, h), max_batch=1)
    mono, kg, dg = e(img)
    ko, do, mo = _oracle().extract(img)
    assert mono == mo
    _assert_same(kg, dg, ko, do)


def test_lapping_area_packing(frames4):
    from geoflowslam_b200 import ORBextractor
    e = ORBextractor(*ORB_CFG, max_size=(640, 480), max_batch=1)
    mono, kg, dg = e(frames4[3], vLappingArea=(200, 400))
    ko, do, mo = _oracle().extract(frames4[3], lapping=(200, 400))
    assert mono == mo and 0 < mono < len(kg)
    _assert_same(kg, dg, ko, do)


def test_other_configs(frames4):
    from geoflowslam_b200 import ORBextractor
    from oracle import oracle as O
    for cfg in [(500, 1.2, 8, 20, 7), (2000, 1.2, 8, 20, 7), (1200, 1.5, 4, 15, 5), (300, 1.1, 3, 30, 10)]:
        e = ORBextractor(*cfg, max_size=(640, 480), max_batch=1)
        mono, kg, dg = e(frames4[1])
        ko, do, mo = O.OrbOracle(*cfg).extract(frames4[1])
        assert mono == mo, cfg
        _assert_same(kg, dg, ko, do)


def test_empty_image_and_errors(ex):
    from geoflowslam_b200 import GfsError
    mono, k, d = ex(np.zeros((0, 0), np.uint8))
    assert mono == -1 and len(k) == 0                                  # ORBextractor.cc:1150
    with pytest.raises(GfsError):
        ex(np.zeros((40, 40), np.uint8))                               # too small for a FAST cell
    with pytest.raises(GfsError):
        ex(np.zeros((2000, 2000), np.uint8))                           # beyond max_size


def test_idempotent_and_batch_order_independent(ex, frames4):
    a = ex.extract_batch(frames4)
    b = ex.extract_batch(frames4[::-1].copy())
    for i in range(4):
        j = 3 - i
        assert a[2][i] == b[2][j]
        assert np.array_equal(a[1][i, :a[2][i]], b[1][j, :b[2][j]])
