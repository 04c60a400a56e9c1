set -x
python -m pytest tests/test_gpu_orb.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q > gpurun_out/s4_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s4_tests.log
tail -5 gpurun_out/s4_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s4_bench.json 2> gpurun_out/s4_bench.err
timeout 300 ncu --set full --clock-control none --import-source on -k k_fast_cells2 -c 1 -f -o gpurun_out/s4_fast2 python bench.py --steps 1 --warmup 1 --no-cpu --batch 256 > gpurun_out/s4_ncu.log 2>&1
cat gpurun_out/s4_bench.json | cut -c1-1500
