set -x
python -m pytest tests/test_gpu_orb.py tests/test_gpu_match.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q > gpurun_out/s13_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s13_tests.log
tail -3 gpurun_out/s13_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s13_bench.json 2> gpurun_out/s13_bench.err
cp geoflowslam_b200/libgfs_b200.so /tmp/lib_keep.so
for v in th40 th56; do cp geoflowslam_b200/_variant_$v.so geoflowslam_b200/libgfs_b200.so; python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s13_bench_$v.json 2>/dev/null; done
cp /tmp/lib_keep.so geoflowslam_b200/libgfs_b200.so
