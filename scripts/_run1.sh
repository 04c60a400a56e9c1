set -x
for s in 2 3 4; do for c in 6 8 10; do GFS_FRONTEND_STREAMS=$s GFS_FRONTEND_CHUNKS=$c python bench.py --steps 10 --warmup 3 --no-cpu > gpurun_out/s7_bench_s${s}_c$c.json 2>/dev/null; done; done
timeout 300 ncu --set full --clock-control none --import-source on -k k_orient_desc -c 1 -f -o gpurun_out/s7_od python bench.py --steps 1 --warmup 1 --no-cpu --batch 256 > gpurun_out/s7_ncu.log 2>&1
python - <<'PY'
import json,glob
for f in sorted(glob.glob("gpurun_out/s7_bench*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1]); print(f, round(d["value"]), round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"],3))
    except Exception as e: print(f, e)
PY
