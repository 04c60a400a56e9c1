#!/bin/bash
# Round 2, GPU session 37: optimiser kernels launched over the unfinished pairs only (compacted list), check interval sweep.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py -x -q > $O/r2s37_tests.log 2>&1; tail -5 $O/r2s37_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s37_bench_gicp_track_$name.json 2> $O/r2s37_bench_gicp_track_$name.err
  python - <<PY
import json
for l in open("$O/r2s37_bench_gicp_track_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1))
PY
}
run c2 GFS_GICP_CHECK_EVERY=2
run c1 GFS_GICP_CHECK_EVERY=1
run c3 GFS_GICP_CHECK_EVERY=3
run c2e16 GFS_GICP_ERR_PPT=16
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/r2s37_bench_track.json 2> $O/r2s37_bench_track.err
python - <<PY
import json
for l in open("$O/r2s37_bench_track.json"):
    if l.startswith("{"):
        d = json.loads(l); print("track", d["value"], d["e2e"]["value"], d["roofline"].get("gicp_stage_ms_per_step"))
PY
