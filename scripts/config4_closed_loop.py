#!/usr/bin/env python
"""BASELINE.json configs[4]: independent synthetic RGB-D-inertial sequences (seeds 4000 + rank), one per GPU, through the
closed-loop tracker (geoflowslam_b200/tracker.py) on the CUDA library and on the CPU oracle; per sequence the two
trajectories, their ATE against the ground truth and the number of differing decisions.

  python scripts/config4_closed_loop.py [--frames 30] [--out DIR]                      (one sequence on cuda:0)
  torchrun --nproc-per-node 8 scripts/config4_closed_loop.py --frames 30 --out DIR     (eight sequences, one per GPU)"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np  # noqa: E402


def run(seed, n_frames, kf_every, out_dir=None, oracle=True):
    from geoflowslam_b200 import imu, synth, tracker
    seq = synth.room_sequence(seed, n_frames=n_frames)
    gt = seq["twb"][:n_frames]
    t0 = time.perf_counter()
    g = tracker.run_tracker(seq, tracker.CudaBackend(), kf_every=kf_every)
    t_gpu = time.perf_counter() - t0
    res = dict(seed=seed, frames=n_frames, keyframes=g["n_keyframes"], map_points=g["n_map_points"], ate_cuda=imu.ate_rmse(g["twb"], gt),
               seconds_cuda=t_gpu, ba_calls=sum(1 for d in g["decisions"] if d[0] == "ba"),
               icp_accepted=sum(1 for d in g["decisions"] if d[0] == "icp" and d[2] == 1))
    runs = [("cuda", g)]
    if oracle:
        from oracle.tracker_backend import OracleBackend
        t0 = time.perf_counter()
        o = tracker.run_tracker(seq, OracleBackend(), kf_every=kf_every)
        res.update(seconds_oracle=time.perf_counter() - t0, ate_oracle=imu.ate_rmse(o["twb"], gt),
                   ate_difference=abs(imu.ate_rmse(o["twb"], gt) - res["ate_cuda"]),
                   max_position_difference=float(np.abs(g["twb"] - o["twb"]).max()), max_rotation_difference=float(np.abs(g["Rwb"] - o["Rwb"]).max()),
                   differing_decisions=sum(1 for a, b in zip(g["decisions"], o["decisions"]) if a != b) + abs(len(g["decisions"]) - len(o["decisions"])))
        runs.append(("oracle", o))
    if out_dir:
        os.makedirs(out_dir, exist_ok=True)
        for name, r in runs:
            Twc = []
            for R, p in zip(r["Rwb"], r["twb"]):
                T = np.eye(4); T[:3, :3] = R @ seq["Rbc"]; T[:3, 3] = R @ seq["tbc"] + p
                Twc.append(T)
            imu.save_trajectory_tum(os.path.join(out_dir, "seq%d_%s.txt" % (seed, name)), seq["stamps"][:n_frames], Twc)
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--frames", type=int, default=30)
    ap.add_argument("--kf-every", type=int, default=5)
    ap.add_argument("--out", default=None)
    ap.add_argument("--no-oracle", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1")); local = int(os.environ.get("LOCAL_RANK", "0"))
    import torch
    torch.cuda.set_device(local)
    res = run(4000 + rank, args.frames, args.kf_every, args.out, oracle=not args.no_oracle)
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("gloo")
        allres = [None] * world
        dist.all_gather_object(allres, res)
        dist.destroy_process_group()
    else:
        allres = [res]
    if rank == 0:
        out = dict(config="configs[4]: %d independent sequences, one per GPU, %d frames each" % (world, args.frames), sequences=allres,
                   worst_ate_difference=max((r.get("ate_difference", 0.0) for r in allres)),
                   differing_decisions=sum(r.get("differing_decisions", 0) for r in allres))
        print(json.dumps(out))


if __name__ == "__main__":
    main()
