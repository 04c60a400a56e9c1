#!/bin/bash
# Round 2, GPU session 38: k_nn_corr3 with a CTA-level queue over several tiles, fallback seeded with the best of the ten neighbours.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py -x -q > $O/r2s38_tests.log 2>&1; tail -5 $O/r2s38_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s38_bench_gicp_track_$name.json 2> $O/r2s38_bench_gicp_track_$name.err
  python - <<PY
import json
for l in open("$O/r2s38_bench_gicp_track_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1))
PY
}
run t1 GFS_GICP_NN_TILES=1
run t2 GFS_GICP_NN_TILES=2
run t4 GFS_GICP_NN_TILES=4
run t8 GFS_GICP_NN_TILES=8
run t16 GFS_GICP_NN_TILES=16
run t8nn7 GFS_GICP_NN_TILES=8 GFS_GICP_NN=7
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/r2s38_bench_track.json 2> $O/r2s38_bench_track.err
python - <<PY
import json
for l in open("$O/r2s38_bench_track.json"):
    if l.startswith("{"):
        d = json.loads(l); print("track", d["value"], d["e2e"]["value"], d["roofline"].get("gicp_stage_ms_per_step"))
PY
