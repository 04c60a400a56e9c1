#!/bin/bash
# Round 2, GPU session 54: launch list of the default command (steady-state steps: the first 1400 launches are skipped).
set -x
O=gpurun_out
mkdir -p $O
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --launch-skip 1400 --launch-count 420 --csv --log-file $O/r2s54_track_launches_b512.csv python bench.py --steps 2 --warmup 3 --no-cpu > $O/r2s54_ncu_list.log 2>&1
tail -2 $O/r2s54_ncu_list.log | cut -c1-300
wc -l $O/r2s54_track_launches_b512.csv
