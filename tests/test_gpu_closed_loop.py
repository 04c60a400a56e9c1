"""BASELINE.json configs[4] / "ATE vs ref": the closed-loop RGB-D-inertial tracker (geoflowslam_b200/tracker.py) over the WHOLE
path -- ORB extraction, SearchByProjection (+ retry) / SearchWithGMS, depth -> cloud, PredictStateICP's GICP gate,
PoseOptimization, SearchLocalPoints, PoseInertialOptimizationLast{KeyFrame,Frame}, keyframes, LocalInertialBA -- once on the
CUDA library and once on the CPU oracle, same host logic.  Every integer decision (match counts, outlier counts, ICP
accept / iterations / inliers, keyframe contents, BA window / LM trials / erased observations) must be IDENTICAL and the
trajectories must agree to 1e-4 (north_star: "ATE within 1e-4 of the reference").  scripts/config4_closed_loop.py runs the
8-sequence, one-per-GPU version under torchrun."""
import numpy as np
import pytest

from geoflowslam_b200 import imu, synth, tracker

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("seed,n_frames,kf_every", [(4000, 16, 3), (4001, 12, 4)])
def test_closed_loop_equals_oracle(seed, n_frames, kf_every, tmp_path):
    from oracle.tracker_backend import OracleBackend
    seq = synth.room_sequence(seed, n_frames=n_frames)
    g = tracker.run_tracker(seq, tracker.CudaBackend(), kf_every=kf_every)
    o = tracker.run_tracker(seq, OracleBackend(), kf_every=kf_every)
    assert len(g["decisions"]) == len(o["decisions"])
    for a, b in zip(g["decisions"], o["decisions"]):
        assert a == b, "decision differs: cuda %s vs oracle %s" % (a, b)
    assert any(d[0] == "ba" for d in g["decisions"]) and any(d[0] == "icp" and d[2] == 1 for d in g["decisions"])
    assert np.abs(g["twb"] - o["twb"]).max() < 1e-4 and np.abs(g["Rwb"] - o["Rwb"]).max() < 1e-4
    gt = seq["twb"][:n_frames]
    ate_g, ate_o = imu.ate_rmse(g["twb"], gt), imu.ate_rmse(o["twb"], gt)
    assert abs(ate_g - ate_o) < 1e-4 and ate_g < 1e-2
    # SaveTrajectoryTUM of both runs: the files agree to the printed precision
    paths = []
    for name, r in (("cuda", g), ("oracle", o)):
        Twc = []
        for R, p in zip(r["Rwb"], r["twb"]):
            T = np.eye(4); T[:3, :3] = R @ seq["Rbc"]; T[:3, 3] = R @ seq["tbc"] + p
            Twc.append(T)
        path = str(tmp_path / (name + ".txt"))
        imu.save_trajectory_tum(path, seq["stamps"][:n_frames], Twc)
        paths.append(path)
    a, b = (np.loadtxt(p) for p in paths)
    assert a.shape == (n_frames, 8) and np.abs(a - b).max() < 1e-4
