#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, average, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[1:]:
    v = float(r[vi].replace(',', '')); v = v / 1e3 if r[ui] in ('ns', 'nsecond') else v
    n = r[ki].split('(')[0]; agg[n][0] += 1; agg[n][1] += v
tot = sum(v[1] for v in agg.values())
print(f"{'kernel':64s} {'launches':>8s} {'total us':>12s} {'avg us':>10s} {'share':>7s}")
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:64s} {v[0]:8d} {v[1]:12.1f} {v[1] / v[0]:10.1f} {v[1] / tot:7.3f}")
