#!/usr/bin/env python
"""BASELINE configs[0] / [4] plumbing on the CPU (no GPU): a synthetic RGB-D-inertial sequence is written to disk in the
layout the reference's example drivers read (geoflowslam_b200/dataset.py), read back through the restated loaders and
frame loop, pushed through the closed-loop chain (geoflowslam_b200/chain.py) on the CPU oracle, and the estimated
trajectory is saved in the SaveTrajectoryTUM format together with its ATE against the generator's ground truth.
  python scripts/config0_plumbing.py [n_frames] [out_dir]"""
import os
import sys
import tempfile

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from geoflowslam_b200 import chain, dataset, imu, synth  # noqa: E402
from oracle import oracle as O  # noqa: E402


class OracleBackend:
    def fb_klt(self, a, b, kps, priors):
        pa, pb = O.klt_build_pyramid(a, 3), O.klt_build_pyramid(b, 3)
        return O.fb_klt_tracking(pa, pb, a.shape[1], a.shape[0], 3, kps, priors)

    def preintegrate(self, rows, bias6):
        return O.imu_preintegrate(rows, bias6, *synth.imu_calib_noise())

    def pose_inertial(self, prob):
        return O.pose_inertial_optimize(prob)


def write_to_disk(seq, root):
    rows = []
    for k, block in enumerate(seq["imu"]):
        t = float(seq["stamps"][k])
        for r in block:                     # a sample is stamped at the end of the interval it covers
            t += float(r[6])
            rows.append([t, *r[0:3], *r[3:6]])
    return dataset.write_sequence(root, seq["stamps"], seq["frames"], seq["depth"], imu=np.array(rows), inertial=True)


def read_from_disk(root, assoc, truth):
    """-> the dict chain.run_chain takes, with images / depth / inertial samples from disk and only the initial state and
    the calibration from `truth`"""
    rgb, dep, ts = dataset.load_images(assoc, inertial=True)
    t_imu, acc, gyr = dataset.load_imu(os.path.join(root, "imu", "imu.txt"))
    frames, depth = zip(*(dataset.read_frame(root, a, b) for a, b in zip(rgb, dep)))
    groups = {ni: ab for ni, ab, _ in dataset.frame_measurements(ts, t_imu)}
    blocks = []
    for k in range(1, len(ts)):
        a, b = groups[k]
        edges = np.concatenate([[ts[k - 1]], t_imu[a:b]])
        blocks.append(np.concatenate([acc[a:b], gyr[a:b], np.diff(edges)[:, None].astype(np.float32)], 1).astype(np.float32))
    seq = dict(truth)
    seq.update(frames=np.stack(frames), depth=np.stack(depth), stamps=ts, imu=blocks)
    return seq


def run(n_frames=12, out_dir=None, seed=8001):
    truth = synth.vio_sequence(seed, n_frames=n_frames)
    root = out_dir or tempfile.mkdtemp(prefix="gfs_seq_")
    assoc = write_to_disk(truth, root)
    seq = read_from_disk(root, assoc, truth)
    est = chain.run_chain(seq, OracleBackend())
    traj = os.path.join(root, "CameraTrajectory.txt")
    Twb = [np.block([[est["Rwb"][k], est["twb"][k][:, None]], [np.zeros((1, 3)), np.ones((1, 1))]]) for k in range(len(est["twb"]))]
    imu.save_trajectory_tum(traj, seq["stamps"][:len(Twb)], Twb)
    ate = imu.ate_rmse(est["twb"], truth["twb"][:len(est["twb"])])
    return dict(root=root, trajectory=traj, ate=ate, est=est, seq=seq, truth=truth)


if __name__ == "__main__":
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 12
    r = run(n, sys.argv[2] if len(sys.argv) > 2 else None)
    print("sequence in %s (%d frames, %d inertial samples)" % (r["root"], len(r["seq"]["frames"]), sum(len(b) for b in r["seq"]["imu"])))
    print("trajectory -> %s, ATE (RMSE after rigid alignment) %.6f m, landmarks %d -> %d" %
          (r["trajectory"], r["ate"], r["est"]["n_tracked"][0], r["est"]["n_tracked"][-1]))
