"""Projection-window search and depth back-projection: oracle self-checks (CPU) and GPU parity."""
import numpy as np
import pytest

from geoflowslam_b200._lib import KP_DTYPE
from oracle import oracle as O

GRID = (0.0, 0.0, 64.0 / 640.0, 48.0 / 480.0)  # mnMinX, mnMinY, mfGridElementWidthInv, mfGridElementHeightInv


def make_case(seed, n=1000, nq=900, mode=0, contention=0.2, occupied_frac=0.1):
    rng = np.random.default_rng(seed)
    k = np.zeros(n, KP_DTYPE)
    k["x"] = rng.uniform(19, 620, n).astype(np.float32); k["y"] = rng.uniform(19, 460, n).astype(np.float32)
    k["octave"] = rng.integers(0, 8, n); k["angle"] = rng.uniform(0, 360, n).astype(np.float32)
    k[: n // 20]["x"] = rng.uniform(-30, 700, n // 20)  # some undistorted points outside the image / grid
    desc = rng.integers(0, 256, (n, 32), dtype=np.uint8)
    ur = np.where(rng.random(n) < 0.7, k["x"] - rng.uniform(5, 40, n), -1).astype(np.float32)
    occ = (rng.random(n) < occupied_frac).astype(np.uint8)
    q = np.zeros(nq, O.PROJ_QUERY_DTYPE)
    tgt = rng.integers(0, n, nq)
    dup = rng.random(nq) < contention                      # several map points competing for one keypoint
    tgt[dup] = tgt[rng.integers(0, nq, dup.sum())]
    q["u"] = k["x"][tgt] + rng.normal(0, 2.0, nq).astype(np.float32)
    q["v"] = k["y"][tgt] + rng.normal(0, 2.0, nq).astype(np.float32)
    lvl = k["octave"][tgt]
    sf = (1.2 ** lvl).astype(np.float32)
    if mode == 0:
        q["radius"] = np.float32(15.0) * sf
        q["min_level"] = lvl - 1; q["max_level"] = lvl + 1
        fwd = rng.random(nq) < 0.2
        q["min_level"][fwd] = lvl[fwd]; q["max_level"][fwd] = -1        # bForward: GetFeaturesInArea(u, v, r, octave)
    else:
        q["radius"] = np.where(rng.random(nq) < 0.5, 2.5, 4.0).astype(np.float32) * np.float32(3.0) * sf
        q["min_level"] = lvl - 1; q["max_level"] = lvl
    q["ur"] = np.where(ur[tgt] > 0, ur[tgt] + rng.normal(0, 1.5, nq), -1).astype(np.float32)
    q["angle"] = (k["angle"][tgt] + rng.normal(0, 4, nq) + np.where(rng.random(nq) < 0.1, 150, 0)).astype(np.float32) % 360
    q["blocks"] = (rng.random(nq) < 0.8).astype(np.int32)
    d = desc[tgt].copy()
    flips = rng.integers(0, 256, (nq, 32), dtype=np.uint8) & rng.integers(0, 256, (nq, 32), dtype=np.uint8) & \
        rng.integers(0, 256, (nq, 32), dtype=np.uint8)
    far = rng.random(nq) < 0.1
    d ^= np.where(far[:, None], rng.integers(0, 256, (nq, 32), dtype=np.uint8), flips)
    q["desc"] = d
    q["radius"][rng.random(nq) < 0.05] = -1.0              # map points the caller filtered out
    return q, k, ur, desc, occ


@pytest.mark.parametrize("mode", [0, 1])
def test_oracle_projection_search_properties(mode):
    q, k, ur, desc, occ = make_case(10 + mode, mode=mode)
    assign, nm = O.search_by_projection(mode, q, k, ur, desc, occ, GRID, nnratio=0.9)
    matched = assign >= 0
    assert 0 < nm and matched.sum() > 100
    assert not (matched & (occ > 0)).any()                 # occupied keypoints are never reassigned
    for idx in np.nonzero(matched)[0][:200]:
        qi = q[assign[idx]]
        assert abs(k["x"][idx] - qi["u"]) < qi["radius"] and abs(k["y"][idx] - qi["v"]) < qi["radius"]
        assert int(np.unpackbits(desc[idx] ^ qi["desc"]).sum()) <= 100
    # without contention and blocking the result must not depend on the order of the map points
    q2, k2, ur2, d2, _ = make_case(20 + mode, mode=mode, contention=0.0, occupied_frac=0.0)
    q2["blocks"] = 0
    a1, n1 = O.search_by_projection(mode, q2, k2, ur2, d2, None, GRID, check_orientation=False)
    perm = np.random.default_rng(0).permutation(len(q2))
    a2, n2 = O.search_by_projection(mode, q2[perm], k2, ur2, d2, None, GRID, check_orientation=False)
    assert n1 == n2
    same = (a1 >= 0) == (a2 >= 0)
    assert same.mean() > 0.99                              # (two map points may still pick one keypoint: last one wins)


def test_oracle_depth_to_cloud():
    rng = np.random.default_rng(1)
    depth = rng.uniform(-1, 12, (480, 640)).astype(np.float32)
    pts = O.depth_to_cloud(depth, 3, 606.986, 607.011, 311.519, 247.26)
    v, u = np.mgrid[0:480:3, 0:640:3]
    d = depth[0:480:3, 0:640:3]
    ok = (d > 0) & (d < 10)
    assert len(pts) == ok.sum() and np.array_equal(pts[:, 2], d[ok]) and (pts[:, 3] == 1).all()
    x = ((u[ok].astype(np.float32) - np.float32(311.519)) * d[ok] / np.float32(606.986)).astype(np.float32)
    assert np.array_equal(pts[:, 0], x)


@pytest.mark.gpu
@pytest.mark.parametrize("mode,seed,n,nq", [(0, 1, 1000, 900), (1, 2, 1000, 3000), (0, 3, 50, 400), (1, 4, 2000, 10),
                                            (0, 5, 1008, 1008)])
def test_gpu_projection_search_equals_oracle(mode, seed, n, nq):
    from geoflowslam_b200.matcher import search_by_projection
    q, k, ur, desc, occ = make_case(seed, n=n, nq=nq, mode=mode, contention=0.4)
    for check_ori in (True, False):
        ga, gn = search_by_projection(mode, q, k, ur, desc, occ, GRID, nnratio=0.8, check_orientation=check_ori)
        oa, on = O.search_by_projection(mode, q, k, ur, desc, occ, GRID, nnratio=0.8, check_orientation=check_ori)
        assert gn == on and np.array_equal(ga, oa)


@pytest.mark.gpu
def test_gpu_projection_search_edge_cases():
    from geoflowslam_b200.matcher import search_by_projection
    q, k, ur, desc, occ = make_case(7, n=300, nq=200)
    a, n = search_by_projection(0, q[:0], k, ur, desc, occ, GRID)          # no map points
    assert n == 0 and (a == -1).all()
    q["blocks"] = 1; q["u"] = k["x"][0]; q["v"] = k["y"][0]; q["desc"] = desc[0]; q["radius"] = 20; q["min_level"] = -1
    q["max_level"] = -1; q["ur"] = -1                                      # every map point wants keypoint 0's area
    ga, gn = search_by_projection(0, q, k, ur, desc, None, GRID, check_orientation=False)
    oa, on = O.search_by_projection(0, q, k, ur, desc, None, GRID, check_orientation=False)
    assert gn == on and np.array_equal(ga, oa)


@pytest.mark.gpu
def test_gpu_depth_to_cloud_equals_oracle():
    from geoflowslam_b200.matcher import depth_to_cloud
    rng = np.random.default_rng(2)
    for shape, stride in [((480, 640), 3), ((480, 640), 1), ((37, 53), 4)]:
        depth = rng.uniform(-1, 12, shape).astype(np.float32)
        g = depth_to_cloud(depth, stride, 606.986, 607.011, 311.519, 247.26)
        o = O.depth_to_cloud(depth, stride, 606.986, 607.011, 311.519, 247.26)
        assert g.shape == o.shape and np.array_equal(g, o)
    assert len(depth_to_cloud(np.zeros((48, 64), np.float32), 3, 600, 600, 32, 24)) == 0
