#!/bin/bash
# Round 2, GPU session 36: k_error with several points per thread (batched gathers), k_linearize register caps.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py -x -q > $O/r2s36_tests.log 2>&1; tail -5 $O/r2s36_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s36_bench_gicp_track_$name.json 2> $O/r2s36_bench_gicp_track_$name.err
  python - <<PY
import json
for l in open("$O/r2s36_bench_gicp_track_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1))
PY
}
run e2 GFS_GICP_ERR_PPT=2
run e4 GFS_GICP_ERR_PPT=4
run e8 GFS_GICP_ERR_PPT=8
run e16 GFS_GICP_ERR_PPT=16
run l4 GFS_GICP_LIN_MINB=4
run l5 GFS_GICP_LIN_MINB=5
run l4p4 GFS_GICP_LIN_MINB=4 GFS_GICP_LIN_PPT=4
