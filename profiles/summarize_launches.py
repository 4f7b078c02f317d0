#!/usr/bin/env python3
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel name.
usage: python profiles/summarize_launches.py gpurun_out/launches.csv > profiles/<name>.txt"""
import collections
import csv
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[hi]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi:
        continue
    try:
        v = float(r[vi].replace(',', ''))
    except ValueError:
        continue
    v = v / 1e3 if r[ui] == 'ns' else (v * 1e3 if r[ui] == 'ms' else v)
    name = r[ki].split('(')[0][:64]
    agg[name][0] += 1
    agg[name][1] += v
tot = sum(v[1] for v in agg.values())
print('# per-launch device time (ncu, cold-cache, serialised): compare SHARES, not absolutes')
print('%-66s %6s %12s %10s %7s' % ('kernel', 'n', 'total_us', 'avg_us', 'share'))
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print('%-66s %6d %12.1f %10.1f %6.1f%%' % (k, v[0], v[1], v[1] / v[0], 100 * v[1] / tot))
