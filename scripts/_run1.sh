set -x
python -m pytest tests/test_gpu_orb.py tests/test_gpu_fullsize.py tests/test_golden.py -m gpu -x -q > gpurun_out/s11_tests.log 2>&1; echo "tests rc=$?" >> gpurun_out/s11_tests.log
tail -3 gpurun_out/s11_tests.log
python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s11_bench.json 2> gpurun_out/s11_bench.err
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s11_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3), {k:round(v["ms"],3) for k,v in d["roofline"]["stages"].items()})
    except Exception as e: print(f, e)
PY
