#!/usr/bin/env python
"""Aggregate an `ncu --metrics gpu__time_duration.sum --csv` launch list per kernel. Usage: launch_summary.py launches.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
i = [k for k, r in enumerate(rows) if r and r[0] == 'ID'][0]
hdr = rows[i]
kn, mv, mu = hdr.index('Kernel Name'), hdr.index('Metric Value'), hdr.index('Metric Unit')
agg, order = {}, []
for r in rows[i + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0][-50:]
    v = float(r[mv].replace(',', ''))
    v = v * 1000 if r[mu] == 'ms' else v / 1000 if r[mu] == 'ns' else v
    if name not in agg:
        agg[name] = []
        order.append(name)
    agg[name].append(v)
tot = sum(sum(v) / len(v) for v in agg.values())
for n in order:
    a = sum(agg[n]) / len(agg[n])
    print(f"{n:52s} launches={len(agg[n]):3d} avg={a:9.1f} us  share={a / tot * 100:5.1f}%")
print(f"{'sum of per-kernel averages':52s} {tot:24.1f} us")
