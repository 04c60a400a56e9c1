#!/usr/bin/env python
"""Sum an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel.  usage: ncu_launch_shares.py file.csv"""
import csv, collections, sys
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
h = rows[0]; ix = {n: i for i, n in enumerate(h)}
t = collections.defaultdict(float); c = collections.Counter()
scale = {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 'nsecond': 1e-6, 'usecond': 1e-3, 'msecond': 1.0}
for r in rows[1:]:
    if len(r) != len(h): continue
    k = r[ix['Kernel Name']].split('(')[0]
    t[k] += float(r[ix['Metric Value']].replace(',', '')) * scale.get(r[ix['Metric Unit']], 1e-6); c[k] += 1
tot = sum(t.values())
print(f"{'kernel':44s} {'n':>5s} {'total ms':>10s} {'avg us':>9s} {'share':>6s}")
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print(f"{k[:44]:44s} {c[k]:5d} {v:10.3f} {1e3 * v / c[k]:9.1f} {100 * v / tot:5.1f}%")
