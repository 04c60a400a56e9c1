#!/bin/bash
# Round 2, GPU session 40: k_klt_track re-stages the J window only when its integer origin moved; GICP state after the clean-up.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_klt.py tests/test_gpu_gicp.py tests/test_gpu_closed_loop.py tests/test_gpu_flow_search.py -x -q > $O/r2s40_tests.log 2>&1; tail -5 $O/r2s40_tests.log
timeout 600 python bench.py --workload klt --steps 10 --warmup 3 --no-cpu > $O/r2s40_bench_klt.json 2> $O/r2s40_bench_klt.err
timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s40_bench_gicp_track.json 2> $O/r2s40_bench_gicp_track.err
timeout 900 python bench.py --steps 20 --warmup 3 --no-cpu > $O/r2s40_bench_track.json 2> $O/r2s40_bench_track.err
python - <<PY
import json
for f in ("klt", "gicp_track", "track"):
    for l in open("$O/r2s40_bench_%s.json" % f):
        if l.startswith("{"):
            d = json.loads(l); print(f, round(d["value"], 1), d.get("e2e", {}).get("value"), d["config"].get("stage_ms_one_stream"))
PY
