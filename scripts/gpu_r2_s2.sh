#!/bin/bash
# Round 2, GPU session 2: the BASELINE-metric bench (track + LocalBA composite) + one line per sub-workload.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests -m gpu -x -q > $O/r2s2_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s2_tests.log
tail -3 $O/r2s2_tests.log
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2s2_track.json 2> $O/r2s2_track.err; echo "track rc=$?"; tail -5 $O/r2s2_track.err
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > $O/r2s2_track_ref.json 2> $O/r2s2_track_ref.err; echo "ref rc=$?"
for w in gicp ba lba pose pose_inertial klt orb; do
  timeout 600 python bench.py --workload $w --steps 5 --warmup 3 > $O/r2s2_$w.json 2> $O/r2s2_$w.err; echo "$w rc=$?"; tail -2 $O/r2s2_$w.err
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s2_*.json')):
    try:
        d = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', d.get('e2e', {}).get('value'), 'cpu', (d.get('cpu_baseline') or {}).get('value'))
        if 'track.json' in f:
            print(json.dumps(d['config']['stage_ms_one_stream']), json.dumps(d['roofline']['gicp_stage_ms_per_step']), d['roofline']['kernel'], d['roofline']['frac'])
    except Exception as e:
        print(f, 'bad', e)
PY
