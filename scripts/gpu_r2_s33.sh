#!/bin/bash
# Round 2, GPU session 33: dense sorted grid (GFS_GICP_GRID=1) against the hash grid, CTAs per SM of the two search kernels.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py -x -q > $O/r2s33_tests.log 2>&1; tail -15 $O/r2s33_tests.log
run() {  # name, env...
  local name=$1; shift
  env "$@" timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s33_bench_gicp_track_$name.json 2> $O/r2s33_bench_gicp_track_$name.err
  python - <<PY
import json
for l in open("$O/r2s33_bench_gicp_track_$name.json"):
    if l.startswith("{"):
        d = json.loads(l); print("$name", round(d["value"], 1))
PY
}
run hash GFS_GICP_GRID=0
run hash_nn8 GFS_GICP_GRID=0 GFS_GICP_NN=8
run hash_nn9 GFS_GICP_GRID=0 GFS_GICP_NN=9
run dense GFS_GICP_GRID=1
run dense_nn8 GFS_GICP_NN=8
run dense_nn9 GFS_GICP_NN=9
run dense_knn3 GFS_GICP_KNN=3
run dense_knn4 GFS_GICP_KNN=4
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/r2s33_bench_track.json 2> $O/r2s33_bench_track.err
python - <<PY
import json
for l in open("$O/r2s33_bench_track.json"):
    if l.startswith("{"):
        d = json.loads(l); print("track", d["value"], d["e2e"]["value"], d["roofline"].get("gicp_stage_ms_per_step"))
PY
