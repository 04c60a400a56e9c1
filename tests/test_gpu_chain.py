"""Closed loop: a 12-frame synthetic RGB-D-inertial sequence through fbKltTracking -> IMU preintegration ->
PoseInertialOptimizationLast{KeyFrame,Frame} (geoflowslam_b200/chain.py), once on the CUDA library and once on the CPU
oracle.  BASELINE.json asks for the ATE within 1e-4 of the reference's: the two trajectories agree to 1e-5 m here, and
both stay within a centimetre of ground truth."""
import numpy as np
import pytest

from geoflowslam_b200 import chain, imu, synth

pytestmark = pytest.mark.gpu


class OracleBackend:
    def fb_klt(self, a, b, kps, priors):
        from oracle import oracle as O
        pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
        return O.fb_klt_tracking(pa, pb, a.shape[1], a.shape[0], 3, kps, priors)

    def preintegrate(self, rows, bias6):
        from oracle import oracle as O
        return O.imu_preintegrate(rows, bias6, *synth.imu_calib_noise())

    def pose_inertial(self, prob):
        from oracle import oracle as O
        return O.pose_inertial_optimize(prob)


def test_closed_loop_trajectory_matches_oracle_and_ground_truth():
    seq = synth.vio_sequence(8000, n_frames=12)
    g = chain.run_chain(seq, chain.CudaBackend())
    o = chain.run_chain(seq, OracleBackend())
    gt = seq["twb"][:12]
    ate_g, ate_o = imu.ate_rmse(g["twb"], gt), imu.ate_rmse(o["twb"], gt)
    assert g["n_tracked"] == o["n_tracked"] and g["n_inliers"] == o["n_inliers"]
    assert np.abs(g["twb"] - o["twb"]).max() < 1e-5 and np.abs(g["Rwb"] - o["Rwb"]).max() < 1e-5
    assert abs(ate_g - ate_o) < 1e-5
    assert ate_g < 0.01 and min(g["n_inliers"][1:]) > 100, (ate_g, g["n_inliers"])
    # without the correction the IMU-only dead reckoning drifts: the chain must beat it
    assert np.linalg.norm(g["twb"][-1] - gt[-1]) < 0.02
