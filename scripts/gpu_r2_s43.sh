#!/bin/bash
# Round 2, GPU session 43: k_klt_track windows through tensor-map TMA loads -- parity + A/B.
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_klt.py tests/test_gpu_flow_search.py tests/test_gpu_closed_loop.py tests/test_gpu_chain.py -x -q > $O/r2s43_tests.log 2>&1; echo "rc=$?" >> $O/r2s43_tests.log; tail -8 $O/r2s43_tests.log
for v in 1 0; do
  GFS_KLT_TMAP=$v timeout 300 python bench.py --workload klt --steps 10 --warmup 3 --no-cpu > $O/r2s43_bench_klt_tmap$v.json 2> $O/r2s43_bench_klt_tmap$v.err
  python - <<PY
import json
for l in open("$O/r2s43_bench_klt_tmap$v.json"):
    if l.startswith("{"):
        d = json.loads(l); print("tmap$v", round(d["value"], 1))
PY
done
