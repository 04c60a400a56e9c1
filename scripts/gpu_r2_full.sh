#!/bin/bash
# Whole GPU suite + headline bench + GICP benches after a change of defaults (outputs gpurun_out/r2s34_*).
set -x
O=gpurun_out
mkdir -p $O
timeout 1500 python -m pytest tests -m gpu -x -q > $O/r2s34_tests.log 2>&1; tail -3 $O/r2s34_tests.log
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2s34_bench_track.json 2> $O/r2s34_bench_track.err; tail -c 600 $O/r2s34_bench_track.json
timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s34_bench_gicp_track.json 2> $O/r2s34_bench_gicp_track.err
timeout 600 python bench.py --workload gicp --batch 512 --steps 3 --warmup 3 --no-cpu > $O/r2s34_bench_gicp.json 2> $O/r2s34_bench_gicp.err
python - <<PY
import json
for f in ("track", "gicp_track", "gicp"):
    for l in open("$O/r2s34_bench_%s.json" % f):
        if l.startswith("{"):
            d = json.loads(l); print(f, round(d["value"], 1), d.get("e2e", {}).get("value"), d["roofline"].get("gicp_stage_ms_per_step"))
PY
