#!/usr/bin/env python
"""Aggregate an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel."""
import csv, sys
from collections import defaultdict
rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ik, iv, iu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= iv:
        continue
    v = float(r[iv].replace(',', ''))
    u = r[iu]
    v = v / 1e3 if u == 'us' else v / 1e6 if u == 'ns' else v * 1e3 if u == 's' else v
    name = r[ik].split('(')[0][:60]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print("total ms %.3f launches %d" % (tot, sum(v[0] for v in agg.values())))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print("%-62s n=%5d  ms=%9.3f  share=%5.1f%%  avg_us=%8.1f" % (k, v[0], v[1], 100 * v[1] / tot, 1e3 * v[1] / v[0]))
