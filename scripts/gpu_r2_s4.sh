#!/bin/bash
# Round 2, GPU session 4 (2 GPUs): partitioned BA over NCCL, sequence-sharded track bench, configs[4] on two GPUs.
set -x
O=gpurun_out
mkdir -p $O
nvidia-smi -L
timeout 900 python -m pytest tests/test_gpu_ba_partition.py -x -q > $O/r2s4_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s4_tests.log
tail -15 $O/r2s4_tests.log
for m in nccl callback; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29741 scripts/ba_partition_check.py --mode $m --json > $O/r2s4_ba_partition_$m.json 2> $O/r2s4_ba_partition_$m.err; echo "$m rc=$?"; cat $O/r2s4_ba_partition_$m.json; tail -3 $O/r2s4_ba_partition_$m.err
done
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29742 bench.py --gpus 2 --steps 10 --warmup 3 > $O/r2s4_track_2gpu.json 2> $O/r2s4_track_2gpu.err; echo "track2 rc=$?"; tail -3 $O/r2s4_track_2gpu.err; cut -c1-600 $O/r2s4_track_2gpu.json
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29743 scripts/config4_closed_loop.py --frames 20 > $O/r2s4_config4_2gpu.json 2> $O/r2s4_config4_2gpu.err; echo "config4 rc=$?"; cat $O/r2s4_config4_2gpu.json
