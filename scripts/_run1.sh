set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s3_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s3_tests.log
tail -5 gpurun_out/s3_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s3_bench_v2.json 2> gpurun_out/s3_bench_v2.err
GFS_FAST_V1=1 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s3_bench_v1.json 2> gpurun_out/s3_bench_v1.err
timeout 300 ncu --set full --clock-control none --import-source on -k k_fast_cells2 -c 1 -f -o gpurun_out/s3_fast2 python bench.py --steps 1 --warmup 1 --no-cpu --batch 256 > gpurun_out/s3_ncu.log 2>&1
cat gpurun_out/s3_bench_v2.json | cut -c1-1800
