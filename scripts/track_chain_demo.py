#!/usr/bin/env python
"""Closed-loop check of BASELINE.json's "ATE vs ref": the synthetic RGB-D-inertial sequence of synth.vio_sequence through
geoflowslam_b200/chain.py on the CUDA library and on the CPU oracle; prints both ATEs and the largest pose difference.
  python scripts/track_chain_demo.py [n_frames]"""
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoflowslam_b200 import chain, imu, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


class OracleBackend:
    def fb_klt(self, a, b, kps, priors):
        pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
        return O.fb_klt_tracking(pa, pb, a.shape[1], a.shape[0], 3, kps, priors)

    def preintegrate(self, rows, bias6):
        return O.imu_preintegrate(rows, bias6, *synth.imu_calib_noise())

    def pose_inertial(self, prob):
        return O.pose_inertial_optimize(prob)


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
    seq = synth.vio_sequence(8000, n_frames=n)
    gt = seq["twb"][:n]
    t0 = time.perf_counter(); g = chain.run_chain(seq, chain.CudaBackend()); tg = time.perf_counter() - t0
    t0 = time.perf_counter(); o = chain.run_chain(seq, OracleBackend()); to = time.perf_counter() - t0
    print("frames %d, path length %.3f m, landmarks %d -> %d tracked at the end" % (n, np.linalg.norm(np.diff(gt, axis=0), axis=1).sum(), g["n_tracked"][0], g["n_tracked"][-1]))
    print("ATE (RMSE after rigid alignment): cuda %.6f m, oracle %.6f m, |difference| %.2e m" % (imu.ate_rmse(g["twb"], gt), imu.ate_rmse(o["twb"], gt), abs(imu.ate_rmse(g["twb"], gt) - imu.ate_rmse(o["twb"], gt))))
    print("largest difference between the two trajectories: position %.2e m, rotation matrix entry %.2e" % (np.abs(g["twb"] - o["twb"]).max(), np.abs(g["Rwb"] - o["Rwb"]).max()))
    print("final position error vs ground truth: cuda %.4f m, oracle %.4f m" % (np.linalg.norm(g["twb"][-1] - gt[-1]), np.linalg.norm(o["twb"][-1] - gt[-1])))
    print("inlier counts equal: %s; wall time (host-pointer calls, one frame at a time): cuda %.2f s, oracle %.2f s" % (g["n_inliers"] == o["n_inliers"], tg, to))


if __name__ == "__main__":
    main()
