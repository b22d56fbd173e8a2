#!/usr/bin/env python
"""Summarise `ncu -i rep --page source --csv --print-source sass,cuda [--kernel-name regex:..]` per CUDA source line
(file:line): share of executed warp instructions and of stall samples, top stall reasons.
Usage: ncu_lines.py export.csv [top_n] [inst|smp]"""
import csv
import os
import sys

path, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 30
key = sys.argv[3] if len(sys.argv) > 3 else "inst"
rows = list(csv.reader(open(path, errors="replace")))
hdr, cur_file, acc = None, "?", {}
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        cur_file = os.path.basename(r[1])
        continue
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
    a = acc.setdefault((cur_file, line), [r[1][:100], 0, 0, {}])
    a[1] += ins
    a[2] += smp
    for k, v in stalls.items():
        a[3][k] = a[3].get(k, 0) + v
tot_i = sum(a[1] for a in acc.values()) or 1
tot_s = sum(a[2] for a in acc.values()) or 1
print(f"total instructions {tot_i}, samples {tot_s}")
by_file = {}
for (f, _), a in acc.items():
    b = by_file.setdefault(f, [0, 0]); b[0] += a[1]; b[1] += a[2]
for f, (i, s) in sorted(by_file.items(), key=lambda kv: -kv[1][0]):
    print(f"  {f:18s} inst {i / tot_i * 100:5.1f}% smp {s / tot_s * 100:5.1f}%")
idx = 1 if key == "inst" else 2
for (f, line), a in sorted(acc.items(), key=lambda kv: -kv[1][idx])[:top]:
    st = ",".join(f"{k[6:]}:{v}" for k, v in sorted(a[3].items(), key=lambda kv: -kv[1])[:3])
    print(f"{f}:{line:<4d} inst {a[1] / tot_i * 100:5.1f}% smp {a[2] / tot_s * 100:5.1f}% [{st}] {a[0].strip()}")
