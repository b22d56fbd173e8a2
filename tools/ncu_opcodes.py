#!/usr/bin/env python
"""Executed-instruction histogram by SASS opcode from an `ncu --page source --csv --print-source sass,cuda` export.
Usage: ncu_opcodes.py export.csv"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
hdr, op, tot = None, {}, 0
seen = set()
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr) or r[0].strip() != "":
        continue
    addr, sass = r[2].strip(), r[3].strip()
    if not sass or addr in seen:
        continue
    seen.add(addr)
    try:
        ins = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    t = sass.split()
    o = t[1] if t[0].startswith("@") and len(t) > 1 else t[0]
    o = ".".join(o.split(".")[:2]) if o.startswith(("IMAD", "SHF", "LDS", "STS", "LDG", "STG", "ATOMS")) else o.split(".")[0]
    op[o] = op.get(o, 0) + ins
    tot += ins
for k, v in sorted(op.items(), key=lambda kv: -kv[1])[:40]:
    print(f"{k:16s} {v / 1e6:9.1f}M {v / tot * 100:5.1f}%")
print(f"total {tot / 1e6:.1f}M")
