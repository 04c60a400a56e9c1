#!/bin/bash
# Round 2, GPU session 44: grid cell size sweep on the dense grid (the 0.11 m optimum was measured on the hash grid).
O=gpurun_out
mkdir -p $O
for c in 0.06 0.07 0.08 0.09 0.10 0.11 0.125; do
  GFS_GICP_CELL=$c GFS_GICP_DENSE_CAP=8388608 timeout 600 python bench.py --workload gicp --gicp-track --batch 128 --steps 5 --warmup 3 --no-cpu > $O/r2s44_cell$c.json 2> $O/r2s44_cell$c.err
  python - <<PY
import json
for l in open("$O/r2s44_cell$c.json"):
    if l.startswith("{"):
        d = json.loads(l); print("cell $c", round(d["value"], 1))
PY
done
