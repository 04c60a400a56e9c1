#!/bin/bash
# Round 2, GPU session 35: ncu captures of the dense-grid search kernels (k_nn_corr3, k_knn_search), k_linearize and k_cov_nbr; launch list.
set -x
O=gpurun_out
mkdir -p $O
B="python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu --gicp-track"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file $O/r2s35_gicp_launches.csv $B > $O/r2s35_ncu_list.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nn_corr3 -s 14 -c 1 -o $O/r2s35_nn_corr3 -f $B > $O/r2s35_ncu_a.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_knn_search -s 2 -c 1 -o $O/r2s35_knn_search -f $B > $O/r2s35_ncu_b.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_linearize -s 14 -c 1 -o $O/r2s35_linearize -f $B > $O/r2s35_ncu_c.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_cov_nbr -s 2 -c 1 -o $O/r2s35_cov_nbr -f $B > $O/r2s35_ncu_d.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_error -s 14 -c 1 -o $O/r2s35_error -f $B > $O/r2s35_ncu_e.log 2>&1
ls -la $O/*.ncu-rep
