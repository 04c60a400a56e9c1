#!/bin/bash
# Round 2, GPU session 1: parity of the new GICP search variants + tracking mode, A/B timings, ncu before/after pages.
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > $O/r2s1_gpu.txt
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s1_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s1_tests.log
for v in "0 0" "1 0" "1 1" "1 2" "0 2"; do
  set -- $v
  GFS_GICP_ORDER=$1 GFS_GICP_NN=$2 timeout 600 python bench.py --workload gicp --batch 128 --steps 3 --warmup 1 --no-cpu > $O/r2s1_gicp_o$1_n$2.json 2> $O/r2s1_gicp_o$1_n$2.err
done
timeout 600 python bench.py --workload gicp --batch 128 --steps 4 --warmup 2 --no-cpu --gicp-track > $O/r2s1_gicp_track.json 2> $O/r2s1_gicp_track.err
# launch lists (serialised, cold): before = first-generation search, after = defaults
GFS_GICP_ORDER=0 GFS_GICP_NN=0 timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2s1_gicp_launches_before.csv python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu > $O/r2s1_ncu_list_before.log 2>&1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file $O/r2s1_gicp_launches_after.csv python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu > $O/r2s1_ncu_list_after.log 2>&1
# full pages: the two hot kernels before and after (one launch each; -s skips the warm-up step's launches of that kernel)
GFS_GICP_ORDER=0 GFS_GICP_NN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nn_corr -s 14 -c 1 -o $O/r2s1_nn_corr_before -f python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu > $O/r2s1_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nn_corr2 -s 14 -c 1 -o $O/r2s1_nn_corr_after -f python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu > $O/r2s1_ncu_b.log 2>&1
GFS_GICP_ORDER=0 GFS_GICP_NN=0 timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_cov -s 1 -c 1 -o $O/r2s1_knn_cov_before -f python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu > $O/r2s1_ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_cov -s 1 -c 1 -o $O/r2s1_knn_cov_after -f python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu > $O/r2s1_ncu_d.log 2>&1
ls -la $O | tail -30
tail -3 $O/r2s1_tests.log
cat $O/r2s1_gicp_*.json | python -c "
import sys, json
for l in sys.stdin:
    try:
        d = json.loads(l); print(d['config'].get('variant'), d['config']['workload'][-30:], round(d['value'],1), d['unit'], round(d['ms_per_step'],2))
    except Exception as e: print('bad line', e)
"
