#!/usr/bin/env python
"""Aggregates an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel: count, total, mean, share."""
import collections, csv, sys
rows = [r for r in csv.reader(open(sys.argv[1])) if r and not r[0].startswith('==')]
hdr = rows[0]
ki, vi, ui = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    k = r[ki].split('(')[0].split('::')[-1]
    v = float(r[vi].replace(',', '')) * {'ns': 1e-3, 'us': 1.0, 'ms': 1e3}.get(r[ui], 1.0)
    a = agg.setdefault(k, [0, 0.0]); a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print(f'{len(rows) - 1} launches, {tot:.1f} us in total (cold-cache, serialised: compare shares, not absolutes)')
for k, (c, t) in sorted(agg.items(), key=lambda x: -x[1][1]):
    print(f'{k:42s} n={c:4d} total={t:10.1f} us  mean={t / c:8.1f} us  share={t / tot * 100:5.1f} %')
