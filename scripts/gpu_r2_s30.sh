#!/bin/bash
# Round 2, GPU session 30: k_linearize with the recursive-halving warp reduction and 1 / 2 / 4 points per thread.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py -x -q > $O/r2s30_tests.log 2>&1; tail -5 $O/r2s30_tests.log
for ppt in 1 2 4; do
  GFS_GICP_LIN_PPT=$ppt timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s30_bench_gicp_track_ppt$ppt.json 2> $O/r2s30_bench_gicp_track_ppt$ppt.err
  python - <<PY
import json
for l in open("$O/r2s30_bench_gicp_track_ppt$ppt.json"):
    if l.startswith("{"):
        d = json.loads(l); print("ppt$ppt", d["value"], d["roofline"].get("gicp_stage_ms_per_step"))
PY
done
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/r2s30_bench_track.json 2> $O/r2s30_bench_track.err
python - <<PY
import json
for l in open("$O/r2s30_bench_track.json"):
    if l.startswith("{"):
        d = json.loads(l); print("track", d["value"], d["e2e"]["value"], d["roofline"].get("gicp_stage_ms_per_step"))
PY
