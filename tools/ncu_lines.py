#!/usr/bin/env python
"""Summarise an `ncu --page source --csv --print-source sass,cuda` export per CUDA source line:
instructions executed and stall samples. Usage: ncu_lines.py export.csv [top_n]"""
import csv
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
rows = list(csv.reader(open(path)))
hdr = None
acc = {}
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[2] != "-":
        continue
    try:
        line = int(r[0])
        ins = int(r[hdr.index("Instructions Executed")])
        smp = int(r[hdr.index("# Samples")])
    except ValueError:
        continue
    stalls = {h: int(v) for h, v in zip(hdr, r) if h.startswith("stall_") and "(" not in h and v.isdigit() and int(v)}
    a = acc.setdefault(line, [r[1][:110], 0, 0, {}])
    a[1] += ins
    a[2] += smp
    for k, v in stalls.items():
        a[3][k] = a[3].get(k, 0) + v
tot_i = sum(a[1] for a in acc.values()) or 1
tot_s = sum(a[2] for a in acc.values()) or 1
print(f"total instructions {tot_i}, samples {tot_s}")
for line, a in sorted(acc.items(), key=lambda kv: -kv[1][2])[:top]:
    st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{line:4d} inst {a[1] / tot_i * 100:5.1f}% smp {a[2] / tot_s * 100:5.1f}% [{st}] {a[0].strip()}")
