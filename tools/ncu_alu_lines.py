#!/usr/bin/env python
"""Per CUDA source line: executed ALU-pipe instructions (LOP3/SHF/IADD3/VIADD/ISETP/SEL/LEA/VIMNMX/PRMT/...) from an
`ncu --page source --csv --print-source sass,cuda` export. Usage: ncu_alu_lines.py export.csv [top_n]"""
import csv, sys
ALU = ("LOP3", "SHF", "IADD3", "VIADD", "ISETP", "SEL", "LEA", "VIMNMX", "PRMT", "PLOP3", "IABS", "FLO", "BREV", "MOV", "IADD", "LOP", "BMSK", "SGXT", "VABSDIFF", "IMNMX")
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 30
hdr, cur, acc, seen, tot, alltot = None, None, {}, set(), 0, 0
for r in rows:
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr is None or len(r) != len(hdr):
        continue
    if r[0].strip():
        cur = (r[0].strip(), r[1].strip()[:100])
        continue
    addr, sass = r[2].strip(), r[3].strip()
    if not sass or addr in seen or cur is None:
        continue
    seen.add(addr)
    try:
        ins = int(r[hdr.index("Instructions Executed")])
    except ValueError:
        continue
    t = sass.split()
    o = (t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]
    alltot += ins
    if o in ALU:
        acc[cur] = acc.get(cur, 0) + ins
        tot += ins
print(f"ALU-pipe instructions {tot / 1e6:.1f}M of {alltot / 1e6:.1f}M")
for (ln, src), v in sorted(acc.items(), key=lambda kv: -kv[1])[:top]:
    print(f"{ln:>5s} {v / 1e6:7.1f}M {v / tot * 100:5.1f}%  {src}")
