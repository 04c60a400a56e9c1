"""The closed-loop tracker's host logic (geoflowslam_b200/tracker.py) on the CPU oracle: every stage of the path is driven in
the reference's order with the reference's gates, and the trajectory stays on the ground truth."""
import numpy as np

from geoflowslam_b200 import imu, synth, tracker
from oracle.tracker_backend import OracleBackend


def test_tracker_on_the_oracle_follows_the_ground_truth():
    seq = synth.room_sequence(8200, n_frames=10)
    r = tracker.run_tracker(seq, OracleBackend(), kf_every=3)
    gt = seq["twb"][:10]
    d = r["decisions"]
    assert d[0][0] == "init" and d[0][1] > 900
    icp = [x for x in d if x[0] == "icp"]; tr = [x for x in d if x[0] == "track"]; ba = [x for x in d if x[0] == "ba"]; kf = [x for x in d if x[0] == "kf"]
    assert len(icp) == 9 and all(x[2] == 1 and x[4] > 200 for x in icp)       # PredictStateICP accepted: converged && inliers > 200
    assert len(tr) == 9 and all(x[2] == "proj" and x[3] >= 25 for x in tr)    # SearchByProjection found enough without the retry
    assert all(x[8] > 300 for x in tr)                                        # PoseInertialOptimization inliers
    assert [x[7] for x in tr][:2] == [0, 1]                                   # LastKeyFrame on the first frame, LastFrame afterwards
    assert len(kf) == 3 and len(ba) == 2 and all(x[-1] == 0 for x in ba)      # LocalInertialBA from the third keyframe on, none failed
    assert tr[6][7] == 0                                                      # the frame after a local BA optimises against the keyframe
    assert imu.ate_rmse(r["twb"], gt) < 5e-3 and np.linalg.norm(r["twb"][-1] - gt[-1]) < 1e-2
    # without the visual correction the inertial dead reckoning alone drifts further than that within these frames
    assert r["n_keyframes"] == 4 and r["n_map_points"] > 1200


def test_gms_fallback_and_retry_gates():
    """A frame whose motion prediction is useless: SearchByProjection finds too few matches twice and the tracker falls back
    to SearchWithGMS (the gates of TrackWithMotionModelICP, src/Tracking.cc:3640-3665)."""
    seq = synth.room_sequence(8201, n_frames=3)
    t = tracker.Tracker(seq, OracleBackend(), kf_every=10, use_icp=False)
    real_predict = t.predict_state_imu

    def bad_predict(F, last):
        real_predict(F, last)
        F.Rwb = F.Rwb @ synth._rot(np.array([0.0, 0.0, 1.2]))      # 70 degrees of yaw: every projection leaves the image

    t.predict_state_imu = bad_predict
    r = t.run(2)
    tr = [x for x in r["decisions"] if x[0] == "track"]
    assert tr[0][2] == "gms" and tr[0][3] > 100
