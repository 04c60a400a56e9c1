"""Pins the GICP oracle (no reference fixtures are vendored): brute-force kNN (the pattern of
small_gicp's kdtree_synthetic_test.cpp:24-130), numpy eigen-decomposition, scipy expm, finite
differences, and pose recovery on synthetic RGB-D clouds."""
import numpy as np
import pytest

from geoflowslam_b200 import synth
from oracle import oracle as O


def test_voxelgrid_matches_numpy_groupby():
    rng = np.random.default_rng(0)
    pts = np.concatenate([rng.uniform(-1, 1, (5000, 3)), np.ones((5000, 1))], 1).astype(np.float32)
    pts[:50, :3] = 0.011  # many points in one voxel
    got = O.voxelgrid(pts, 0.05)
    key = np.floor(pts[:, :3].astype(np.float64) / 0.05).astype(np.int64)
    order = np.lexsort((np.arange(len(pts)), key[:, 0], key[:, 1], key[:, 2]))  # z-major key order, then index
    ks = key[order]
    starts = np.r_[0, np.nonzero((np.diff(ks, axis=0) != 0).any(1))[0] + 1, len(pts)]
    ref = []
    for a, b in zip(starts[:-1], starts[1:]):
        s = np.zeros(3)
        for i in order[a:b]:
            s = s + pts[i, :3].astype(np.float64)
        ref.append(s / float(b - a))
    assert np.array_equal(np.array(ref), got)
    assert len(O.voxelgrid(np.zeros((0, 4), np.float32), 0.05)) == 0


@pytest.mark.parametrize("dist", ["uniform", "normal", "planes"])
def test_kdtree_knn_against_brute_force(dist):
    rng = np.random.default_rng(1)
    if dist == "uniform":
        pts = rng.uniform(-1, 1, (3000, 3))
    elif dist == "normal":
        pts = rng.normal(0, 1, (3000, 3))
    else:
        pts = rng.uniform(-1, 1, (3000, 3)); pts[:, 2] = np.round(pts[:, 2] * 2) / 2 + rng.normal(0, 1e-3, 3000)
    q = rng.uniform(-1.2, 1.2, (200, 3))
    idx, d2, f = O.kdtree_knn(pts, q, 20)
    D = ((pts[None] - q[:, None]) ** 2).sum(2)
    ref = np.argsort(D, 1, kind="stable")[:, :20]
    assert (f == 20).all()
    assert np.array_equal(idx, ref)
    assert np.allclose(d2, np.take_along_axis(D, ref, 1), rtol=0, atol=1e-12)
    idx, d2, f = O.kdtree_knn(pts[:7], q, 10)  # fewer points than k
    assert (f == 7).all() and (idx[:, 7:] == -1).all()


def test_eig3_direct_against_numpy():
    rng = np.random.default_rng(2)
    for _ in range(200):
        A = rng.normal(0, 1, (3, 3)); A = A @ A.T * rng.uniform(1e-6, 1e2)
        ev, V = O.eig3(A)
        w, U = np.linalg.eigh(A)
        assert np.allclose(ev, w, rtol=1e-9, atol=1e-12 * abs(w).max())
        assert np.allclose(V.T @ V, np.eye(3), atol=1e-7)
        assert np.allclose(A @ V, V * ev, atol=1e-7 * abs(w).max())


def test_covariance_is_regularised_plane():
    rng = np.random.default_rng(3)
    n = np.array([0.3, -0.5, 0.8]); n /= np.linalg.norm(n)
    basis = np.linalg.svd(n[None])[2][1:]
    pts = (rng.uniform(-1, 1, (400, 2)) @ basis) + rng.normal(0, 1e-4, (400, 1)) * n
    cov = O.covariances(pts, 10)
    C = np.zeros((400, 3, 3))
    C[:, 0, 0], C[:, 0, 1], C[:, 0, 2], C[:, 1, 1], C[:, 1, 2], C[:, 2, 2] = cov.T
    C[:, 1, 0], C[:, 2, 0], C[:, 2, 1] = C[:, 0, 1], C[:, 0, 2], C[:, 1, 2]
    ref = np.eye(3) - 0.999 * np.outer(n, n)
    assert np.abs(C - ref).max() < 5e-3
    w = np.linalg.eigvalsh(C)
    assert np.allclose(w[:, 0], 1e-3, atol=1e-9) and np.allclose(w[:, 1:], 1.0, atol=1e-9)
    assert np.array_equal(O.covariances(pts[:4], 10), np.tile([1, 0, 0, 1, 0, 1.0], (4, 1)))  # < 5 neighbours


def test_se3_exp_against_scipy():
    from scipy.linalg import expm
    rng = np.random.default_rng(4)
    for a in list(rng.normal(0, 0.5, (20, 6))) + [np.zeros(6), np.r_[1e-7, 0, 0, 1, 2, 3]]:
        w, t = a[:3], a[3:]
        X = np.zeros((4, 4)); X[:3, :3] = [[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]]; X[:3, 3] = t
        assert np.allclose(O.se3_exp(a), expm(X), atol=1e-12)


def test_align_recovers_pose_and_reports_reference_fields():
    tgt, src, T = synth.gicp_pair(2000, n_target=20000)
    r = O.gicp_align(tgt, src)
    assert r["converged"] and r["num_inliers"] > 0.8 * r["n_source"]
    assert np.abs(r["T"] - T).max() < 1e-3
    assert np.allclose(r["H"], r["H"].T) and np.all(np.linalg.eigvalsh(r["H"]) > 0)
    assert 0 <= r["iterations"] < 20 and r["inner_evals"] >= r["iterations"] + 1
    # thread count must not change the result (fixed summation order)
    r1 = O.gicp_align(tgt, src, threads=1)
    assert np.array_equal(r["T"], r1["T"]) and r["error"] == r1["error"]


def test_align_gradient_by_finite_differences():
    """b = sum J^T M r is the gradient of e(T exp(d)) at d = 0: check it through max_iter = 1 runs."""
    tgt, src, T = synth.gicp_pair(2001, n_target=6000)
    r0 = O.gicp_align(tgt, src, T0=T, max_iter=1, threads=1)
    g = np.zeros(6)
    h = 1e-6
    for k in range(6):
        d = np.zeros(6); d[k] = h
        ep = O.gicp_align(tgt, src, T0=T @ O.se3_exp(d), max_iter=1, threads=1)["error"]
        em = O.gicp_align(tgt, src, T0=T @ O.se3_exp(-d), max_iter=1, threads=1)["error"]
        g[k] = (ep - em) / (2 * h)
    # correspondences / Mahalanobis matrices move with T, so agreement is loose but sign+scale must hold
    assert np.allclose(g, r0["b"], rtol=0.2, atol=0.05 * np.abs(r0["b"]).max())


def test_align_degenerate_inputs():
    tgt, src, _ = synth.gicp_pair(2002, n_target=3000)
    far = src.copy(); far[:, :3] += 50.0           # no correspondence within 0.1 m
    r = O.gicp_align(tgt, far)
    assert r["num_inliers"] == 0 and not np.isnan(r["T"]).any()
    r = O.gicp_align(tgt, tgt)                      # identical clouds: identity, converged at once
    assert r["converged"] and np.allclose(r["T"], np.eye(4), atol=1e-9)


def test_predict_state_icp_gating_and_pose_update():
    """Tracking::PredictStateICP (Tracking.cc:3364-3413) around RegisterPointClouds: target = last cloud, source = current
    cloud, init = Tcw_last * Tcw_cur^-1, accepted iff converged && num_inliers > 200, new pose = T^-1 * Tcw_last."""
    from geoflowslam_b200 import synth
    from geoflowslam_b200.gicp import predict_state_icp
    from oracle import oracle as O
    tgt, src, T_true = synth.gicp_pair(2100, n_target=6000)
    calls = []

    def register(t, s, T0):
        calls.append((len(t), len(s), np.array(T0)))
        return O.gicp_align(t, s, T0, threads=2)

    Tl = np.eye(4, dtype=np.float32); Tl[:3, 3] = [0.1, -0.2, 0.05]
    Tc_guess = Tl.copy()                                             # constant-pose prior: init = identity
    r = predict_state_icp(register, Tl, Tc_guess, tgt, src)
    assert calls[0][0] == len(tgt) and calls[0][1] == len(src) and np.allclose(calls[0][2], np.eye(4), atol=1e-6)
    assert r["ok"] and r["result"]["converged"] and r["result"]["num_inliers"] > 200
    # T_target_source maps current-camera points into the last camera: Tcw_cur = T^-1 * Tcw_last
    assert np.allclose(r["Tcw"], np.linalg.inv(r["delta"]) @ Tl.astype(np.float64), atol=1e-5)
    assert np.allclose(r["delta"], T_true, atol=5e-3) and r["Tcw"].dtype == np.float32
    assert abs(r["pos_error"] - np.linalg.norm(np.linalg.inv(r["delta"])[:3, 3])) < 1e-6
    # gates: too few points -> no registration at all; too few inliers -> pose untouched
    n0 = len(calls)
    assert not predict_state_icp(register, Tl, Tc_guess, tgt[:9], src)["ok"] and len(calls) == n0
    few = predict_state_icp(register, Tl, Tc_guess, tgt[:150], src[:150])
    assert not few["ok"] and np.array_equal(few["Tcw"], Tc_guess)


def test_local_ba_icp_edges_gating():
    """The EdgeICP block of LocalInertialBA (Optimizer.cc:3262-3318): which keyframe pairs are registered and which
    registrations become edges."""
    from geoflowslam_b200 import synth
    from geoflowslam_b200.gicp import local_ba_icp_edges
    from oracle import oracle as O
    tgt, src, T_true = synth.gicp_pair(2101, n_target=8000)
    T_true = np.asarray(T_true, np.float64)
    # three keyframes: 0 has no predecessor, 1 follows 0 (clouds tgt -> src), 2 follows 1 but tracks well (> 75 inliers)
    T0 = np.eye(4); T1 = np.linalg.inv(T_true) @ T0; T2 = T1.copy()
    calls = []

    def register(t, s, init):
        calls.append(np.array(init))
        return O.gicp_align(t, s, init, threads=2)

    e = local_ba_icp_edges(register, [T0, T1, T2], [-1, 0, 1], [20, 30, 120], [tgt, src, src])
    assert len(calls) == 1 and np.allclose(calls[0], T0 @ np.linalg.inv(T1), atol=1e-12)   # only keyframe 1 is registered
    assert e["n_icp"] == 1 and e["icp_kf1"].tolist() == [0] and e["icp_kf2"].tolist() == [1]
    rel = np.eye(4); rel[:3, :3] = e["icp_Rt"][0, :9].reshape(3, 3); rel[:3, 3] = e["icp_Rt"][0, 9:]
    assert np.allclose(rel, T_true, atol=5e-3) and np.allclose(rel, e["results"][0][1]["T"])
    # a bad initial guess that the registration corrects by more than 0.1 m in x-y is not trusted
    Tbad = T1.copy(); Tbad[0, 3] += 0.09; Tbad[1, 3] -= 0.09
    e2 = local_ba_icp_edges(register, [T0, Tbad], [-1, 0], [20, 30], [tgt, src])
    r2 = e2["results"][0][1]
    d = np.asarray(r2["T"]) @ np.linalg.inv(T0 @ np.linalg.inv(Tbad))
    expect = bool(r2["converged"] and r2["num_inliers"] > 400 and r2["error"] / r2["num_inliers"] < 0.01
                  and np.hypot(d[0, 3], d[1, 3]) < 0.1)
    assert e2["n_icp"] == int(expect)
    # too few inliers: tiny clouds never make an edge
    assert local_ba_icp_edges(register, [T0, T1], [-1, 0], [20, 30], [tgt[:300], src[:300]])["n_icp"] == 0
