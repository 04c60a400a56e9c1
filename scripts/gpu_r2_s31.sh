#!/bin/bash
# Round 2, GPU session 31: 10-NN search split from the covariance (GFS_GICP_KNN=2..5 = 6 / 7 / 5 / 4 CTAs per SM), k_linearize points per thread.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py -x -q > $O/r2s31_tests.log 2>&1; tail -5 $O/r2s31_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s31_bench_gicp_track_$name.json 2> $O/r2s31_bench_gicp_track_$name.err
  python - <<PY
import json
for l in open("$O/r2s31_bench_gicp_track_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1), d["roofline"].get("gicp_stage_ms_per_step"))
PY
}
run knn0 GFS_GICP_KNN=0
run knn2 GFS_GICP_KNN=2
run knn3 GFS_GICP_KNN=3
run knn4 GFS_GICP_KNN=4
run knn5 GFS_GICP_KNN=5
run ppt8 GFS_GICP_LIN_PPT=8
run ppt16 GFS_GICP_LIN_PPT=16
