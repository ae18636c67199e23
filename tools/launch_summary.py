#!/usr/bin/env python3
"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list by kernel."""
import collections, csv, io, sys
rows = [l for l in open(sys.argv[1]) if l.startswith('"')]
tot = collections.defaultdict(lambda: [0, 0.0])
for x in csv.DictReader(io.StringIO(''.join(rows))):
    if x['Metric Name'] != 'gpu__time_duration.sum':
        continue
    k = x['Kernel Name'].split('(')[0]
    v = float(x['Metric Value'].replace(',', ''))
    u = x['Metric Unit']
    v = v / 1e3 if u == 'ns' else v * 1e3 if u == 'ms' else v
    tot[k][0] += 1; tot[k][1] += v
s = sum(v[1] for v in tot.values())
print(f"{'kernel':42s} {'n':>4s} {'total us':>12s} {'share':>7s}")
for k, v in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print(f"{k:42s} {v[0]:4d} {v[1]:12.1f} {100 * v[1] / s:6.1f}%")
print(f"{'total':42s} {'':4s} {s:12.1f}")
