#!/bin/bash
# Round 2, final validation: what the driver runs at round end (GPU tests, smoke, both bench arms) + the sub-workload lines.
set -x
O=gpurun_out
mkdir -p $O
timeout 1200 python -m pytest tests -m gpu -x -q > $O/r2final_tests.log 2>&1; echo "tests rc=$?" >> $O/r2final_tests.log; tail -4 $O/r2final_tests.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > $O/r2final_smoke.log 2>&1; echo "smoke rc=$?"; tail -7 $O/r2final_smoke.log
timeout 900 python bench.py --impl reference --gpus 1 --steps 5 --warmup 1 > $O/r2final_track_ref.json 2> $O/r2final_track_ref.err; echo "ref rc=$?"
( time timeout 900 python bench.py --gpus 1 --steps 20 --warmup 3 > $O/r2final_track.json 2> $O/r2final_track.err ) 2> $O/r2final_track.time; echo "track rc=$?"; tail -3 $O/r2final_track.err; cat $O/r2final_track.time
for w in gicp ba lba pose pose_inertial klt orb; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $O/r2final_$w.json 2> $O/r2final_$w.err; echo "$w rc=$?"
done
timeout 600 python bench.py --workload gicp --steps 5 --warmup 2 --batch 128 --gicp-track > $O/r2final_gicp_track.json 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2final_*.json')):
    try:
        d = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', d.get('e2e', {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
        if f.endswith('final_track.json'):
            print(json.dumps(d['config']['stage_ms_one_stream']), json.dumps(d['roofline']['gicp_stage_ms_per_step']), d['roofline']['kernel'], d['roofline']['frac'], d['gpu_launches'])
            print(json.dumps(d['config']['sub_lines']))
    except Exception as e:
        print(f, 'bad', e)
PY
