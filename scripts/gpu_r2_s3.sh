#!/bin/bash
# Round 2, GPU session 3: closed-loop tracker parity (configs[4]) on one GPU.
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_closed_loop.py tests/test_gpu_chain.py -x -q > $O/r2s3_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s3_tests.log
tail -30 $O/r2s3_tests.log
timeout 900 python scripts/config4_closed_loop.py --frames 30 --kf-every 5 --out $O/r2s3_traj > $O/r2s3_config4.json 2> $O/r2s3_config4.err; echo "config4 rc=$?"
tail -3 $O/r2s3_config4.err; cat $O/r2s3_config4.json
