"""The closed-loop chain (geoflowslam_b200/chain.py) on the CPU oracle alone: fbKltTracking -> IMU preintegration ->
PoseInertialOptimizationLast{KeyFrame,Frame} on an 8-frame synthetic RGB-D-inertial sequence stays on ground truth."""
import numpy as np

from geoflowslam_b200 import chain, imu, synth
from oracle import oracle as O


class OracleBackend:
    def fb_klt(self, a, b, kps, priors):
        pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
        return O.fb_klt_tracking(pa, pb, a.shape[1], a.shape[0], 3, kps, priors)

    def preintegrate(self, rows, bias6):
        return O.imu_preintegrate(rows, bias6, *synth.imu_calib_noise())

    def pose_inertial(self, prob):
        return O.pose_inertial_optimize(prob)


def test_oracle_chain_tracks_ground_truth():
    seq = synth.vio_sequence(8001, n_frames=8)
    o = chain.run_chain(seq, OracleBackend())
    gt = seq["twb"][:8]
    assert o["n_tracked"][0] > 150 and o["n_tracked"][-1] > 0.9 * o["n_tracked"][0]
    assert imu.ate_rmse(o["twb"], gt) < 2e-3 and np.linalg.norm(o["twb"][-1] - gt[-1]) < 5e-3
    # dead reckoning alone (no visual correction) would have drifted further: biases start at zero
    dr = seq["twb"][0] + seq["vel"][0] * seq["stamps"][7]
    assert np.linalg.norm(o["twb"][-1] - gt[-1]) < np.linalg.norm(dr - gt[-1])
    for k in range(8):
        assert abs(np.linalg.det(o["Rwb"][k]) - 1) < 1e-5
