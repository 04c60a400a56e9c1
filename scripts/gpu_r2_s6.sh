#!/bin/bash
# Round 2, GPU session 6: seven-cell fast path of the correspondence search; default bench with sub-lines.
set -x
O=gpurun_out
mkdir -p $O
timeout 900 python -m pytest tests/test_gpu_gicp.py -x -q > $O/r2s6_tests.log 2>&1; echo "tests rc=$?" >> $O/r2s6_tests.log
tail -5 $O/r2s6_tests.log
for v in "0 1" "0 5" "0 6"; do
  set -- $v
  GFS_GICP_KNN=$1 GFS_GICP_NN=$2 timeout 600 python bench.py --workload gicp --batch 128 --steps 3 --warmup 1 --no-cpu --gicp-track > $O/r2s6_gicp_k$1_n$2.json 2> $O/r2s6_gicp_k$1_n$2.err
  tail -2 $O/r2s6_gicp_k$1_n$2.err
done
timeout 900 python bench.py --steps 10 --warmup 3 > $O/r2s6_track.json 2> $O/r2s6_track.err; echo "track rc=$?"; tail -5 $O/r2s6_track.err
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_nn_corr2 -s 14 -c 1 -o $O/r2s6_nn_corr_fast -f python bench.py --workload gicp --batch 64 --steps 1 --warmup 1 --no-cpu --gicp-track > $O/r2s6_ncu_b.log 2>&1
python - <<'PY'
import json, glob
for f in sorted(glob.glob('gpurun_out/r2s6_*.json')):
    try:
        d = json.loads(open(f).read().strip().split('\n')[-1])
        print(f, round(d['value'], 1), d['unit'], 'ms/step', round(d['ms_per_step'], 2), 'e2e', d.get('e2e', {}).get('value'))
        if 'track' in f:
            print(json.dumps(d['config']['stage_ms_one_stream']), json.dumps(d['roofline']['gicp_stage_ms_per_step']), d['roofline']['kernel'], d['roofline']['frac'])
            print(json.dumps(d['config']['sub_lines'], indent=1))
    except Exception as e:
        print(f, 'bad', e)
PY
