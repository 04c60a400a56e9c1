#!/bin/bash
# Round 2, GPU session 42: FAST cell ROIs staged with one tensor-map TMA load per CTA (UTMALDG) -- parity + A/B.
set -x
O=gpurun_out
mkdir -p $O
timeout 300 python -m pytest tests/test_gpu_orb.py -x -q > $O/r2s42_tests.log 2>&1; echo "rc=$?" >> $O/r2s42_tests.log; tail -8 $O/r2s42_tests.log
for v in 1 0; do
  GFS_ORB_TMAP=$v timeout 300 python bench.py --workload orb --steps 10 --warmup 3 --no-cpu > $O/r2s42_bench_orb_tmap$v.json 2> $O/r2s42_bench_orb_tmap$v.err
  python - <<PY
import json
for l in open("$O/r2s42_bench_orb_tmap$v.json"):
    if l.startswith("{"):
        d = json.loads(l); print("tmap$v", round(d["value"], 1), d["config"].get("stage_ms"), d["roofline"].get("kernel"))
PY
done
