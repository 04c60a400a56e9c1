"""Parity at BASELINE.json's full sizes, through properties that do not need the oracle on every unit:
position independence inside a batch (the same frame / pair / problem gives the same bits wherever it
sits), a brute-force check of the matcher's argmin on sampled queries, sortedness of the GMS inputs,
and the oracle itself on a sample of the batch."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_configs1_batch_1024_frames():
    """1024 VGA frames (32 distinct, tiled) through gfs_frontend_run: every copy of a frame yields the
    same keypoints / descriptors, every copy of a pair the same matches; a sample is checked against
    the oracle; BF distances are re-derived with numpy popcounts and are the row minima."""
    from geoflowslam_b200 import TrackingFrontend, synth
    from oracle import oracle as O
    distinct = synth.orb_frames(32, group=8)
    frames = np.concatenate([distinct] * 32)                     # frame i = distinct[i % 32]
    fe = TrackingFrontend(1000, 1.2, 8, 25, 7, max_size=(640, 480), max_batch=1024)
    out = fe.run(frames)
    n = out["n"]
    assert n.shape == (1024,) and n.min() > 200 and n.max() <= fe.stride
    for i in range(32, 1024):                                    # position independence
        j = i % 32
        assert n[i] == n[j]
        assert np.array_equal(out["kp"][i, :n[i]], out["kp"][j, :n[j]])
        assert np.array_equal(out["desc"][i, :n[i]], out["desc"][j, :n[j]])
    for p in range(32, 1023):
        q = p % 32
        if q == 31:
            continue                                             # pair (31 -> 0 of the next tile) has no twin in tile 0
        assert np.array_equal(out["train_idx"][p, :n[p]], out["train_idx"][q, :n[q]])
        assert out["inlier_count"][p] == out["inlier_count"][q]
        assert np.array_equal(out["inlier"][p, :n[p]], out["inlier"][q, :n[q]])
    orc = O.OrbOracle(1000, 1.2, 8, 25, 7)
    for i in (0, 7, 500, 1023):                                  # oracle on a sample
        ko, do, mo = orc.extract(frames[i])
        assert n[i] == len(ko)
        for f in ("x", "y", "size", "angle", "response", "octave"):
            assert np.array_equal(out["kp"][i, :n[i]][f], ko[f])
        assert np.array_equal(out["desc"][i, :n[i]], do)
    pop = np.array([bin(v).count("1") for v in range(256)], np.int32)
    rng = np.random.default_rng(0)
    for p in rng.integers(0, 1023, 6):                           # matcher: argmin property on sampled pairs
        d1, d2 = out["desc"][p, :n[p]], out["desc"][p + 1, :n[p + 1]]
        for qi in rng.integers(0, n[p], 20):
            dist = pop[np.bitwise_xor(d2, d1[qi])].sum(1)
            assert out["dist"][p, qi] == dist.min() and out["train_idx"][p, qi] == int(dist.argmin())
    assert out["inlier_count"].shape[0] >= 1023 and (out["inlier_count"][:1023] >= 0).all()


def test_configs2_full_size_pairs():
    """50 000-point pairs: oracle parity on one pair, position independence across a batch of 16."""
    from geoflowslam_b200 import RegistrationGICP, synth
    from oracle import oracle as O
    pairs = [synth.gicp_pair(2000 + i, n_target=50000) for i in range(2)]
    stride = max(max(len(t), len(s)) for t, s, _ in pairs)
    reg = RegistrationGICP(max_points=stride, max_pairs=16)
    tg = np.zeros((16, stride, 4), np.float32); sr = np.zeros((16, stride, 4), np.float32)
    nt = np.zeros(16, np.int32); ns = np.zeros(16, np.int32)
    for i in range(16):
        t, s, _ = pairs[i % 2]
        tg[i, :len(t)] = t; sr[i, :len(s)] = s; nt[i] = len(t); ns[i] = len(s)
    res = reg.align_batch(tg, nt, sr, ns, np.tile(np.eye(4), (16, 1, 1)))
    for i in range(2, 16):
        for k in ("T", "H", "b"):
            assert np.array_equal(res[i][k], res[i % 2][k]), k
        assert res[i]["iterations"] == res[i % 2]["iterations"] and res[i]["num_inliers"] == res[i % 2]["num_inliers"]
    o = O.gicp_align(pairs[0][0], pairs[0][1])
    g = res[0]
    assert g["iterations"] == o["iterations"] and g["num_inliers"] == o["num_inliers"] and bool(g["converged"]) == o["converged"]
    assert np.allclose(np.array(g["T"]).reshape(4, 4), o["T"], rtol=1e-7, atol=1e-7)
    assert np.allclose(np.array(g["H"]).reshape(6, 6), o["H"], rtol=1e-6, atol=1e-6 * np.abs(o["H"]).max())
    # the recovered transform maps source onto target: far below the initial misalignment
    Tt = pairs[0][2]
    assert np.abs(np.array(g["T"]).reshape(4, 4)[:3, 3] - Tt[:3, 3]).max() < 0.01


def test_configs3_batch_of_64_problems():
    """64 copies of two configs[3] problems: identical bits per copy, oracle parity on one."""
    from geoflowslam_b200 import Optimizer, synth
    from oracle import oracle as O
    probs = [synth.ba_problem(seed=3000 + i) for i in range(2)]
    opt = Optimizer(max_kf=21, max_points=3000, max_obs=16384, max_inertial=20, max_batch=64)
    res = opt.LocalInertialBA_batch([probs[i % 2] for i in range(64)])
    for i in range(2, 64):
        for k in ("kf_twb", "kf_Rwb", "pt_xyz", "obs_chi2", "obs_outlier"):
            assert np.array_equal(res[i][k], res[i % 2][k]), k
        assert res[i]["lm_trials"] == res[i % 2]["lm_trials"]
    o = O.ba_solve(probs[1])
    g = res[1]
    assert g["lm_trials"] == o["lm_trials"] and g["iterations_done"] == o["iterations_done"]
    assert np.allclose(g["kf_twb"], o["kf_twb"], atol=1e-6) and np.allclose(g["pt_xyz"], o["pt_xyz"], atol=1e-6)


def test_pose_inertial_batch_of_512_frames():
    """Size-independent properties at a tracking-rate batch: every frame of the batch is solved exactly like the same
    frame alone (the batch shares no state), results are permutation-equivariant, and the prior that comes out is
    symmetric positive semi-definite."""
    from geoflowslam_b200 import PoseInertialOptimizer, synth
    uniq = [synth.pose_inertial_problem(6300 + i, mode=i % 2, n_obs=200 + 37 * (i % 7)) for i in range(16)]
    order = np.random.default_rng(0).permutation(512)
    batch = [uniq[j % 16] for j in order]
    opt = PoseInertialOptimizer(max_obs=512, max_batch=512)
    res = opt.optimize_batch(batch)
    one = PoseInertialOptimizer(max_obs=512, max_batch=1)
    ref = [one.optimize_batch([p])[0] for p in uniq]
    for j, r in zip(order, res):
        o = ref[j % 16]
        assert np.array_equal(r["twb"], o["twb"]) and np.array_equal(r["Rwb"], o["Rwb"]) and np.array_equal(r["outlier"], o["outlier"])
        assert np.array_equal(r["H"], o["H"]) and r["n_inliers"] == o["n_inliers"]
    for o in ref:
        H = o["H"]
        assert np.allclose(H, H.T, rtol=0, atol=1e-9 * np.abs(H).max())
        assert np.linalg.eigvalsh((H + H.T) / 2).min() > -1e-9 * np.abs(H).max()
