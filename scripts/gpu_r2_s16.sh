#!/bin/bash
# Round 2, GPU session 16: full validation -- every GPU test, smoke, the default bench and the reference arm, sub-workload lines.
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2s16_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s16_tests.log; tail -4 $O/r2s16_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2s16_smoke.log 2>&1; echo "smoke rc=$?"; tail -6 $O/r2s16_smoke.log
timeout 900 python bench.py --impl reference --steps 5 --warmup 1 > $O/r2s16_track_ref.json 2> $O/r2s16_track_ref.err; echo "ref rc=$?"
timeout 900 python bench.py --steps 20 --warmup 3 > $O/r2s16_track.json 2> $O/r2s16_track.err; echo "track rc=$?"; tail -3 $O/r2s16_track.err
timeout 600 python bench.py --workload orb --steps 10 --warmup 3 > $O/r2s16_orb.json 2> $O/r2s16_orb.err; echo "orb rc=$?"
timeout 600 python bench.py --workload gicp --steps 3 --warmup 1 > $O/r2s16_gicp.json 2> $O/r2s16_gicp.err; echo "gicp rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file $O/r2s16_track_launches.csv python bench.py --steps 1 --warmup 3 --batch 32 --no-cpu > $O/r2s16_ncu_list.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s16_*.json')):
    try:
        d = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', d.get('e2e', {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
        if f.endswith('track.json'):
            print(json.dumps(d['config']['stage_ms_one_stream']), json.dumps(d['roofline']['gicp_stage_ms_per_step']), d['roofline']['kernel'], d['roofline']['frac'])
            print(json.dumps(d['cpu_baseline'].get('cv2_cross_check')), json.dumps(d['config']['sub_lines']['ate_vs_oracle']))
    except Exception as e:
        print(f, 'bad', e)
PY
