#!/bin/bash
# Round 2, GPU session 48: ncu capture of k_klt_track with the tensor-map window loads.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_klt_track -s 2 -c 1 -o $O/r2s48_klt_track -f python bench.py --workload klt --batch 256 --steps 1 --warmup 1 --no-cpu > $O/r2s48_ncu_a.log 2>&1
ls -la $O/r2s48*
