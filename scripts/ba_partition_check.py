"""torchrun --nproc-per-node N scripts/ba_partition_check.py : landmark-partitioned LocalInertialBA over
NCCL vs the single-GPU solve of the same problem (run on the GPU box)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from geoflowslam_b200 import Optimizer, synth

rank, local, world = int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"]), int(os.environ["WORLD_SIZE"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
p = synth.ba_problem(seed=3000)
single = Optimizer(max_batch=1).LocalInertialBA(p)
opt = Optimizer(max_batch=1)
opt.set_partition(rank, world)
opt.upload([p]); 
torch.cuda.synchronize(); dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); opt.solve_uploaded(); e1.record(); torch.cuda.synchronize()
part = opt.download()[0]
own = np.arange(p["n_points"]) % world == rank
ok = part["iterations_done"] == single["iterations_done"] and part["lm_trials"] == single["lm_trials"]
errs = {k: float(np.abs(part[k] - single[k]).max()) for k in ("kf_Rwb", "kf_twb", "kf_vel", "kf_bg", "kf_ba")}
errs["pt_owned"] = float(np.abs(part["pt_xyz"][own] - single["pt_xyz"][own]).max())
ok = ok and max(errs.values()) < 1e-6
s0 = Optimizer(max_batch=1); s0.upload([p]); s0.solve_uploaded(); torch.cuda.synchronize()
f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
f0.record(); s0.solve_uploaded(); f1.record(); torch.cuda.synchronize()
print("rank %d/%d partitioned BA %s  max errs %s  partitioned %.2f ms vs single-GPU %.2f ms" %
      (rank, world, "OK" if ok else "MISMATCH", errs, e0.elapsed_time(e1), f0.elapsed_time(f1)), flush=True)
dist.barrier()
dist.destroy_process_group()
sys.exit(0 if ok else 1)
