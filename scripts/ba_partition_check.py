"""torchrun --nproc-per-node N scripts/ba_partition_check.py [--mode nccl|callback] [--json]

Landmark-partitioned LocalInertialBA (BASELINE configs[3], "1->8 GPU edge-partitioned with NCCL J^T J all-reduce") against the
single-GPU solve of the same problem: every rank linearises the edges of the landmarks it owns, the Schur-reduced pose system
and the chi2 / gain-ratio partials are summed over the ranks, every rank solves the reduced system and back-substitutes its own
landmarks.  mode nccl: ncclAllReduce called directly on the solve stream (gfs_ba_set_partition_nccl); mode callback: the
torch.distributed callback of round 1 (a host sync + a Python call per buffer)."""
import argparse
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from geoflowslam_b200 import Optimizer, synth

ap = argparse.ArgumentParser()
ap.add_argument("--mode", default="nccl", choices=["nccl", "callback"])
ap.add_argument("--json", action="store_true")
ap.add_argument("--reps", type=int, default=5)
args = ap.parse_args()
rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = synth.ba_problem(seed=3000)
single = Optimizer(max_batch=1).LocalInertialBA(p)
opt = Optimizer(max_batch=1)
if args.mode == "nccl":
    opt.set_partition_nccl(rank, world)
else:
    opt.set_partition(rank, world)
opt.upload([p])
opt.solve_uploaded()   # warm-up (NCCL channel set-up)
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(args.reps):
    opt.solve_uploaded()
e1.record(); torch.cuda.synchronize()
ms_part = e0.elapsed_time(e1) / args.reps
part = opt.download()[0]
own = np.arange(p["n_points"]) % world == rank
ok = part["iterations_done"] == single["iterations_done"] and part["lm_trials"] == single["lm_trials"]
errs = {k: float(np.abs(part[k] - single[k]).max()) for k in ("kf_Rwb", "kf_twb", "kf_vel", "kf_bg", "kf_ba")}
errs["pt_owned"] = float(np.abs(part["pt_xyz"][own] - single["pt_xyz"][own]).max())
ok = bool(ok and max(errs.values()) < 1e-6)
s0 = Optimizer(max_batch=1); s0.upload([p]); s0.solve_uploaded(); torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record()
for _ in range(args.reps):
    s0.solve_uploaded()
f1.record(); torch.cuda.synchronize()
ms_single = f0.elapsed_time(f1) / args.reps
t = torch.tensor([ms_part, ms_single, float(ok), max(errs.values())], dtype=torch.float64, device="cuda")
mx = t.clone(); dist.all_reduce(mx, op=dist.ReduceOp.MAX)
mn = t.clone(); dist.all_reduce(mn, op=dist.ReduceOp.MIN)
if args.json:
    if rank == 0:
        print(json.dumps(dict(workload="configs[3]: one LocalInertialBA problem (20 KF x 3000 MP x ~15k edges), landmarks partitioned p % world", mode=args.mode,
                              n_gpus=world, ok=bool(mn[2].item() > 0.5), max_state_difference_vs_single_gpu=float(mx[3].item()),
                              ms_partitioned=float(mx[0].item()), ms_single_gpu=float(mx[1].item()), lm_trials=int(part["lm_trials"]),
                              nccl_launches_per_solve=int(opt.last_nccl_calls()) if args.mode == "nccl" else None,
                              note="a latency-bound solve: partitioning one problem costs more in collectives than it saves in edges (SURVEY.md 8e); throughput scales by batching independent problems per GPU")), flush=True)
else:
    print("rank %d/%d %s partitioned BA %s  max errs %s  partitioned %.2f ms vs single-GPU %.2f ms" %
          (rank, world, args.mode, "OK" if ok else "MISMATCH", errs, ms_part, ms_single), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
