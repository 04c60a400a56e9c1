#!/bin/bash
# Round 2, GPU session 51 (8 GPUs): the headline at N = 8 under torchrun, as the driver launches it.
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus 8 --steps 20 --warmup 3 > $O/r2s51_bench_track_n8.json 2> $O/r2s51_bench_track_n8.err
python - <<PY
import json
for l in open("$O/r2s51_bench_track_n8.json"):
    if l.startswith("{"):
        d = json.loads(l); print("N=8", d["value"], d["e2e"]["value"], d["n_gpus"], d["ms_per_step"], d["clocks"])
PY
tail -3 $O/r2s51_bench_track_n8.err
