"""GPU parity: BF Hamming + GMS kernels (through the C ABI) vs the CPU oracle, bit-exact."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("nq,nt,seed", [(1000, 1000, 0), (37, 1024, 1), (1, 1, 2), (513, 77, 3), (1024, 5, 4),
                                        (3000, 2500, 5)])
def test_bf_match_equals_oracle(nq, nt, seed):
    from geoflowslam_b200 import ORBmatcher
    from oracle import oracle as O
    rng = np.random.default_rng(seed)
    dq = rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    dt = rng.integers(0, 256, (nt, 32), dtype=np.uint8)
    dt[rng.integers(0, nt, nt // 4)] = dt[rng.integers(0, nt, nt // 4)]  # duplicate rows -> ties
    dq[: min(nq, nt) // 2] = dt[: min(nq, nt) // 2]                       # exact matches
    gi, gd = ORBmatcher.bf_match(dq, dt)
    oi, od = O.bf_match(dq, dt)
    assert np.array_equal(gi, oi) and np.array_equal(gd, od)


def test_bf_empty_sets():
    from geoflowslam_b200 import ORBmatcher
    d = np.zeros((5, 32), np.uint8)
    gi, gd = ORBmatcher.bf_match(d, np.zeros((0, 32), np.uint8))
    assert (gi == -1).all() and (gd == -1).all()
    gi, gd = ORBmatcher.bf_match(np.zeros((0, 32), np.uint8), d)
    assert len(gi) == 0


def test_descriptor_distance():
    from geoflowslam_b200 import ORBmatcher
    rng = np.random.default_rng(0)
    for _ in range(5):
        a = rng.integers(0, 256, 32, dtype=np.uint8); b = rng.integers(0, 256, 32, dtype=np.uint8)
        assert ORBmatcher.DescriptorDistance(a, b) == int(np.unpackbits(a ^ b).sum())


def _match_set(seed, n, frac):
    rng = np.random.default_rng(seed)
    p1 = np.stack([rng.uniform(19, 620, n), rng.uniform(19, 460, n)], 1).astype(np.float32)
    p2 = np.clip(p1 + rng.normal(0, 1.0, p1.shape) + [6, -4], 0, [639, 479]).astype(np.float32)
    tr = np.arange(n)
    bad = rng.random(n) > frac
    tr[bad] = rng.integers(0, n, bad.sum())
    return p1, p2, np.stack([np.arange(n), tr], 1).astype(np.int32)


@pytest.mark.parametrize("seed,n,frac", [(0, 1000, 0.6), (1, 1000, 0.1), (2, 300, 0.9), (3, 3000, 0.5), (4, 1, 1.0),
                                         (5, 33, 1.0)])
def test_gms_equals_oracle(seed, n, frac):
    from geoflowslam_b200 import ORBmatcher
    from oracle import oracle as O
    p1, p2, m = _match_set(seed, n, frac)
    gm, gc = ORBmatcher.gms_filter(p1, (640, 480), p2, (640, 480), m)
    om, oc = O.gms_filter(p1, (640, 480), p2, (640, 480), m)
    assert gc == oc and np.array_equal(gm, om)


def test_gms_degenerate():
    from geoflowslam_b200 import ORBmatcher
    from oracle import oracle as O
    p = np.zeros((0, 2), np.float32)
    gm, gc = ORBmatcher.gms_filter(p, (640, 480), p, (640, 480), np.zeros((0, 2), np.int32))
    assert gc == 0 and len(gm) == 0
    p1 = np.full((500, 2), 100.0, np.float32)  # every match in one cell pair (count > 255)
    m = np.stack([np.arange(500), np.arange(500)], 1).astype(np.int32)
    gm, gc = ORBmatcher.gms_filter(p1, (640, 480), p1, (640, 480), m)
    om, oc = O.gms_filter(p1, (640, 480), p1, (640, 480), m)
    assert gc == oc == 500 and np.array_equal(gm, om)
    # points on the image border (x == w -> right index wraps, left index out of range)
    p2 = np.array([[640, 10], [639.9, 479.9], [0, 0], [630, 470]], np.float32)
    m = np.array([[0, 1], [1, 0], [2, 2], [3, 3]], np.int32)
    gm, gc = ORBmatcher.gms_filter(p2, (640, 480), p2, (640, 480), m)
    om, oc = O.gms_filter(p2, (640, 480), p2, (640, 480), m)
    assert gc == oc and np.array_equal(gm, om)


def test_extract_match_gms_pipeline(frames4):
    """configs[1] slice: extract two frames of one scene, BF + GMS, everything vs the oracle."""
    from geoflowslam_b200 import ORBextractor, ORBmatcher
    from oracle import oracle as O
    e = ORBextractor(1000, 1.2, 8, 25, 7, max_size=(640, 480), max_batch=2)
    kps, desc, n, _ = e.extract_batch(frames4[:2])
    k1, d1, k2, d2 = kps[0, :n[0]], desc[0, :n[0]], kps[1, :n[1]], desc[1, :n[1]]
    m, mask, cnt = ORBmatcher().SearchWithGMS(k1, d1, k2, d2, (640, 480))
    oi, _ = O.bf_match(d1, d2)
    assert np.array_equal(m[:, 1], oi)
    p1 = np.stack([k1["x"], k1["y"]], 1); p2 = np.stack([k2["x"], k2["y"]], 1)
    om, oc = O.gms_filter(p1, (640, 480), p2, (640, 480), m)
    assert cnt == oc and np.array_equal(mask, om)
    assert cnt > 100  # same scene under a small homography: most matches are coherent


def test_frontend_batch_equals_oracle():
    """gfs_frontend_run (host buffers) on 6 frames: every output vs the oracle, bit-exact."""
    from geoflowslam_b200 import TrackingFrontend, synth
    from oracle import oracle as O
    frames = synth.orb_frames(6, group=3)
    fe = TrackingFrontend(1000, 1.2, 8, 25, 7, max_size=(640, 480), max_batch=6)
    out = fe.run(frames)
    orc = O.OrbOracle(1000, 1.2, 8, 25, 7)
    ref = [orc.extract(f) for f in frames]
    for i, (ko, do, mo) in enumerate(ref):
        n = out["n"][i]
        assert n == len(ko) and out["mono"][i] == mo
        for f in ("x", "y", "size", "angle", "response", "octave"):
            assert np.array_equal(out["kp"][i, :n][f], ko[f])
        assert np.array_equal(out["desc"][i, :n], do)
    for p in range(5):
        (k1, d1, _), (k2, d2, _) = ref[p], ref[p + 1]
        oi, od = O.bf_match(d1, d2)
        n1 = len(k1)
        assert np.array_equal(out["train_idx"][p, :n1], oi) and np.array_equal(out["dist"][p, :n1], od)
        m = np.stack([np.arange(n1), oi], 1)
        om, oc = O.gms_filter(np.stack([k1["x"], k1["y"]], 1), (640, 480), np.stack([k2["x"], k2["y"]], 1),
                              (640, 480), m)
        assert out["inlier_count"][p] == oc
        assert np.array_equal(out["inlier"][p, :n1].astype(bool), om)
    # frames 2 -> 3 straddle two different scenes: GMS must reject (almost) everything
    assert out["inlier_count"][2] < out["inlier_count"][0] // 4


def test_frontend_pipelined_host_path_equals_device_path():
    """batch > chunk with pinned buffers takes the chunked, stream-overlapped path: same bits."""
    import torch
    from geoflowslam_b200 import TrackingFrontend, synth
    frames = synth.orb_frames(144, group=8)                    # 144 > 64: three chunks
    fe = TrackingFrontend(1000, 1.2, 8, 25, 7, max_size=(640, 480), max_batch=144)
    ref = fe.run(frames)                                       # pageable buffers: single-shot path
    keep = []

    def pinned(shape, dt):
        nb = int(np.prod(shape)) * np.dtype(dt).itemsize
        t = torch.empty(max(nb, 1), dtype=torch.uint8).pin_memory(); keep.append(t)
        return t.numpy()[:nb].view(dt).reshape(shape)

    h_imgs = pinned(frames.shape, np.uint8); h_imgs[...] = frames
    out = fe.alloc_host_outputs(144, pinned)
    fe.run(h_imgs, out=out)
    for k in ref:
        if k in ("kp", "desc"):
            for i in range(144):
                n = ref["n"][i]
                assert np.array_equal(ref[k][i, :n], out[k][i, :n]), (k, i)
        elif k in ("train_idx", "dist", "inlier"):
            for i in range(143):
                n = ref["n"][i]
                assert np.array_equal(ref[k][i, :n], out[k][i, :n]), (k, i)
        else:
            assert np.array_equal(ref[k], out[k]), k
