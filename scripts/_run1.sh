set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s14_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s14_tests.log
tail -3 gpurun_out/s14_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s14_bench.json 2> gpurun_out/s14_bench.err
for cfg in "4 6" "4 10" "6 8" "6 10" "8 12"; do set -- $cfg; GFS_FRONTEND_STREAMS=$1 GFS_FRONTEND_CHUNKS=$2 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s14_bench_s$1_c$2.json 2>/dev/null; done
ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/s14_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu > gpurun_out/s14_ncu_list.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k k_orient_desc -c 1 -f -o gpurun_out/s14_od python bench.py --steps 1 --warmup 1 --no-cpu --batch 256 > gpurun_out/s14_ncu.log 2>&1
timeout 200 python bench.py --workload track --steps 3 --warmup 1 > gpurun_out/s14_track.json 2> gpurun_out/s14_track.err
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/s14_smoke.log 2>&1
