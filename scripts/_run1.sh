set -x
python -m pytest tests -m gpu -x -q > gpurun_out/s8_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s8_tests.log
tail -3 gpurun_out/s8_tests.log
python bench.py --steps 10 --warmup 3 > gpurun_out/s8_bench.json 2> gpurun_out/s8_bench.err
for s in 6 8; do GFS_FRONTEND_STREAMS=$s python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s8_bench_s${s}.json 2>/dev/null; done
GFS_FRONTEND_STREAMS=8 GFS_FRONTEND_CHUNKS=12 python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s8_bench_s8_c12.json 2>/dev/null
timeout 200 compute-sanitizer --tool memcheck python -m pytest tests/test_gpu_orb.py -m gpu -x -q -k "single_frame or pitch or wide" > gpurun_out/s8_memcheck.log 2>&1
tail -5 gpurun_out/s8_memcheck.log
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s8_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["roofline"]["stages"].items()})
    except Exception as e: print(f, e)
PY
