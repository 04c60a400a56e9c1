"""ORBmatcher::SearchByProjectionWithOF (reference src/ORBmatcher.cc:2303-2497): the product's array version
(geoflowslam_b200.matcher.search_by_projection_with_of) against the statement-by-statement restatement in the oracle,
both driven by the same fbKltTracking (the CPU oracle here; the CUDA tracker is bit-exact with it, tests/test_gpu_klt.py)
and by OpenCV's own findFundamentalMat, as in the reference."""
import numpy as np
import pytest

K = (606.986, 607.011, 311.519, 247.260)     # g1_op_icp_lidar_indoor1.yaml:25-28
BOUNDS = (0.0, 640.0, 0.0, 480.0)


def _scenario(seed, n=260):
    import cv2
    from geoflowslam_b200 import synth
    rng = np.random.default_rng(seed)
    fr = synth.orb_frames(2, 640, 480, group=8, seed0=4000 + seed)
    last = cv2.goodFeaturesToTrack(fr[0], n, 0.01, 7).reshape(-1, 2).astype(np.float32)
    cur = cv2.goodFeaturesToTrack(fr[1], 40, 0.01, 25).reshape(-1, 2).astype(np.float32)
    state = rng.choice([0, 1, 1, 1, 2], size=len(last)).astype(np.int32)
    # map points: the last frame's keypoints back-projected at random depths in the last camera frame (= world)
    z = rng.uniform(0.8, 5.0, len(last)).astype(np.float32)
    X = np.stack([(last[:, 0] - K[2]) / K[0] * z, (last[:, 1] - K[3]) / K[1] * z, z], 1).astype(np.float32)
    X[rng.random(len(last)) < 0.05] *= -1.0                      # a few behind the camera (invzc < 0)
    bad = rng.random(len(last)) < 0.15                           # wrong map points: the projection lands 20-60 px away, the
    X[bad, :2] += (rng.uniform(0.03, 0.1, (int(bad.sum()), 2)) * rng.choice([-1, 1], (int(bad.sum()), 2)) * z[bad, None]).astype(np.float32)  # first pass loses them or the F check does
    a = 0.01
    R = np.array([[np.cos(a), -np.sin(a), 0], [np.sin(a), np.cos(a), 0], [0, 0, 1]], np.float32)
    t = np.array([0.02, -0.01, 0.03], np.float32)
    return fr, cur, last, state, X, R, t


@pytest.mark.parametrize("seed", [0, 1])
def test_product_equals_oracle_restatement(seed):
    from geoflowslam_b200.matcher import search_by_projection_with_of
    from oracle import oracle as O
    fr, cur, last, state, X, R, t = _scenario(seed)
    tracker = O.fb_klt_tracking_images
    m0 = np.zeros((480, 640), np.uint8)
    n_o, tracked_o, mask_o = O.search_by_projection_with_of(cur, last, state, X, R, t, K, BOUNDS, fr[0], fr[1], m0.copy(),
                                                            tracker=tracker)
    mask_p = m0.copy()
    n_p, ids_p, pts_p = search_by_projection_with_of(tracker, cur, last, state, X, R, t, K, BOUNDS, fr[0], fr[1], mask_p)
    assert n_p == n_o and n_p > 30
    assert [int(i) for i in ids_p] == [i for i, _ in tracked_o]
    assert np.array_equal(pts_p, np.array([p for _, p in tracked_o], np.float32).reshape(-1, 2))
    assert np.array_equal(mask_p, mask_o)
    # no two accepted points share a mask disc centre, none sits on a pixel occupied by a current keypoint
    assert len(set(map(tuple, np.round(pts_p).astype(int)))) == len(pts_p)
    assert not set(ids_p.tolist()) & set(np.flatnonzero(state == 2).tolist())


def test_everything_rejected_and_empty_inputs():
    from geoflowslam_b200.matcher import search_by_projection_with_of
    from oracle import oracle as O
    fr, cur, last, state, X, R, t = _scenario(2, n=60)
    mask = np.zeros((480, 640), np.uint8)
    # no trackable point at all: every map point bad
    n, ids, pts = search_by_projection_with_of(O.fb_klt_tracking_images, cur, last, np.full(len(last), 2), X, R, t, K, BOUNDS,
                                               fr[0], fr[1], mask)
    assert n == 0 and len(ids) == 0 and pts.shape == (0, 2)
    # a mask that is already full rejects every tracked point (isPointNearby)
    full = np.full((480, 640), 255, np.uint8)
    n, ids, pts = search_by_projection_with_of(O.fb_klt_tracking_images, cur, last, np.zeros(len(last), np.int32), X, R, t, K,
                                               BOUNDS, fr[0], fr[1], full)
    assert n == 0


def test_filter_outliers_equals_oracle_restatement():
    from geoflowslam_b200.matcher import filter_outliers
    from oracle import oracle as O
    fr, cur, last, state, X, R, t = _scenario(3)
    has = state == 1
    n_o, out_o = O.filter_outliers(last, has, X, R, t, K, 1.0)
    n_p, out_p = filter_outliers(last, has, X, R, t, K, 1.0)
    assert n_p == n_o and np.array_equal(out_p, out_o)
    assert 0 < n_p < int(has.sum()) and out_p.any()          # the displaced map points are rejected
    # at most 8 pairs: no F check, nothing flagged (:233)
    few = np.zeros(len(last), bool); few[:8] = True
    n8, out8 = filter_outliers(last, few, X, R, t, K, 1.0)
    assert n8 == 0 and not out8.any()
