#!/usr/bin/env python
"""Static view of the candidate-round loop of k_pileup<16576,false> in a built library: SASS between the first
instruction of `const int j = r + lane;` and the last of `if (m_after) break;` (count_chunk), with source lines.
Usage: sass_loop.py lib.so path/to/kernels.cuh [out.txt]. Prints instruction and opcode-class counts."""
import collections, glob, os, re, subprocess, sys, tempfile

lib, kfile = sys.argv[1], sys.argv[2]
out = sys.argv[3] if len(sys.argv) > 3 else None
tmp = tempfile.mkdtemp()
subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, stdout=subprocess.DEVNULL)
cubin = glob.glob(os.path.join(tmp, "*.cubin"))[0]
L = subprocess.run(["nvdisasm", "-g", "-c", cubin], capture_output=True, text=True).stdout.split("\n")
st = [i for i, l in enumerate(L) if l.startswith(".text._ZN5mgatk8k_pileupILi16576ELb0")][0]
en = [i for i in range(st + 1, len(L)) if L[i].startswith("//-----")][0]
src = open(kfile).read().split("\n")
find = lambda s: [i + 1 for i, l in enumerate(src) if s in l][0]
l0, l1 = find("const int j = r + lane;"), find("if (m_after) break;")
cf = cur = None
rows = []
for l in L[st:en]:
    m = re.match(r'\s*//## File "(.*)", line (\d+)', l)
    if m:
        cf, cur = m.group(1).split("/")[-1], int(m.group(2))
        continue
    m = re.match(r"\s*/\*([0-9a-f]+)\*/\s+(.*?);", l)
    if m:
        rows.append((int(m.group(1), 16), cf, cur, m.group(2).strip()))
    elif l.startswith(".L_"):
        rows.append((None, None, None, l.strip()))
i0 = min(i for i, r in enumerate(rows) if r[1] == "kernels.cuh" and r[2] == l0)
i1 = max(i for i, r in enumerate(rows) if r[1] == "kernels.cuh" and r[2] == l1)
rows = rows[i0:i1 + 1]
ops = collections.Counter()
for a, f, ln, s in rows:
    if a is not None:
        t = s.split()
        ops[(t[1] if t[0].startswith("@") and len(t) > 1 else t[0]).split(".")[0]] += 1
print("instructions", sum(ops.values()), dict(ops.most_common(14)))
if out:
    with open(out, "w") as fh:
        for a, f, ln, s in rows:
            fh.write(s + "\n" if a is None else f"{a:05x} {f}:{ln:<5d} {s}\n")
