#!/bin/bash
# Round 2, GPU session 46 (2 GPUs): the tests that need two GPUs + the headline at N = 2 under torchrun.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_ba_partition.py tests/test_gpu_closed_loop.py -x -q > $O/r2s46_tests_2gpu.log 2>&1; tail -4 $O/r2s46_tests_2gpu.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus 2 --steps 20 --warmup 3 > $O/r2s46_bench_track_n2.json 2> $O/r2s46_bench_track_n2.err
python - <<PY
import json
for l in open("$O/r2s46_bench_track_n2.json"):
    if l.startswith("{"):
        d = json.loads(l); print("N=2", d["value"], d["e2e"]["value"], d["n_gpus"], d["ms_per_step"])
PY
