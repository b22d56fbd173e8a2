#!/usr/bin/env python
"""Host-side throughput of the native dense-plane text writer (csrc/textio.cpp) on planes of the bench shape:
rows/s and MB/s of text for the five output.*.txt.gz files, next to the per-row Python formatting + gzip.open it
replaced (timed on a slice of the cells). CPU only. Usage: bench_textio.py [cells] [threads] [level]"""
import gzip, os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mgatk2_b200.engine import pos_pad
from mgatk2_b200.textio import write_plane_file

n_cells = int(sys.argv[1]) if len(sys.argv) > 1 else 500
threads = int(sys.argv[2]) if len(sys.argv) > 2 else 0
level = int(sys.argv[3]) if len(sys.argv) > 3 else 9
P = 16569
rng = np.random.default_rng(1)
planes = np.zeros((n_cells, 11, pos_pad(P)), np.uint16)
depth = rng.poisson(3.0, size=(n_cells, 8, P)).astype(np.uint16)          # ~12x coverage over 8 base x strand planes
planes[:, :8, :P] = depth
planes[:, 10, :P] = depth.sum(axis=1)
names = [f"ACGTACGTACGT{c:04d}-1" for c in range(n_cells)]
cells = list(range(n_cells))
files = (("A", 0, 1), ("C", 2, 3), ("G", 4, 5), ("T", 6, 7), ("coverage", 10, -1))
with tempfile.TemporaryDirectory() as d:
    t0 = time.perf_counter()
    rows = sum(write_plane_file(os.path.join(d, f"output.{n}.txt.gz"), planes, P, None, a, b, cells, names, level, threads) for n, a, b in files)
    t_native = time.perf_counter() - t0
    size = sum(os.path.getsize(os.path.join(d, f"output.{n}.txt.gz")) for n, _, _ in files)
    text = sum(len(gzip.open(os.path.join(d, f"output.{n}.txt.gz")).read()) for n, _, _ in files)
    k = max(1, n_cells // 25)                                               # the Python way on a slice
    t0 = time.perf_counter()
    for n, a, b in files:
        parts = []
        for c in range(k):
            va, vb = planes[c, a, :P], planes[c, b, :P] if b >= 0 else None
            if vb is None:
                pos = np.nonzero(va)[0]
                parts.append("".join(f"{p + 1},{names[c]},{v}\n" for p, v in zip(pos.tolist(), va[pos].tolist())))
            else:
                pos = np.nonzero((va > 0) | (vb > 0))[0]
                parts.append("".join(f"{p + 1},{names[c]},{x},{y}\n" for p, x, y in zip(pos.tolist(), va[pos].tolist(), vb[pos].tolist())))
        with gzip.open(os.path.join(d, f"py.{n}.txt.gz"), "wb", compresslevel=level) as f:
            f.write("".join(parts).encode())
    t_py = (time.perf_counter() - t0) * n_cells / k
print(f"{n_cells} cells, {rows / 1e6:.1f} M rows, {text / 1e6:.0f} MB of text -> {size / 1e6:.0f} MB gzip level {level}: native {t_native:.2f} s "
      f"({rows / t_native / 1e6:.1f} M rows/s, {text / t_native / 1e6:.0f} MB/s, {threads or os.cpu_count()} threads); "
      f"python per-row formatting + gzip.open {t_py:.1f} s (extrapolated from {k} cells) = {t_py / t_native:.0f}x")
