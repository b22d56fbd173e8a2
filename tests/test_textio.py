"""Native dense-plane text writer (csrc/textio.cpp) against the row rule of the reference's
IncrementalTextWriter (src/file_io/writers.py:440-486) spelled out in Python: rows "pos,barcode,fwd,rev" where either
count is non-zero, "pos,barcode,depth" for the coverage file, 1-based positions, cells in list order, exact values
for saturated entries, any number of threads, multi-member gzip read back as one stream."""
import gzip
import os
import re

import numpy as np
import pytest

from mgatk2_b200.engine import OVERFLOW_DTYPE, pos_pad
from mgatk2_b200.textio import EXPORTS, load, write_plane_file
from tests.helpers import ROOT


def expected_text(planes, P, overflow, a, b, cells, names):
    exact = {(int(o["cell"]), int(o["plane_pos"]) >> 24, int(o["plane_pos"]) & 0xFFFFFF): int(o["value"]) for o in overflow}
    out = []
    for c, bc in zip(cells, names):
        va = planes[c, a, :P].astype(np.int64)
        vb = planes[c, b, :P].astype(np.int64) if b >= 0 else np.zeros(P, np.int64)
        for p in np.nonzero((va > 0) | (vb > 0))[0].tolist():
            x = exact.get((c, a, p), int(va[p]))
            if b >= 0:
                out.append(f"{p + 1},{bc},{x},{exact.get((c, b, p), int(vb[p]))}\n")
            else:
                out.append(f"{p + 1},{bc},{x}\n")
    return "".join(out).encode()


def test_exports_match_header():
    header = open(os.path.join(ROOT, "include", "mgatk2_textio.h")).read()
    declared = set(re.findall(r"\b(mgatk_text_[a-z_]+)\s*\(", header))
    assert declared == set(EXPORTS)
    lib = load()
    for sym in declared:
        assert hasattr(lib, sym), sym


@pytest.mark.parametrize("threads,level", [(1, 9), (3, 1), (0, 6)])
def test_rows_match_reference_rule(tmp_path, threads, level):
    rng = np.random.default_rng(threads * 10 + level)
    n_cells, P = 41, 16569
    planes = np.zeros((n_cells, 11, pos_pad(P)), np.uint16)
    dense = rng.random((n_cells, 11, P)) < 0.3
    planes[:, :, :P] = np.where(dense, rng.integers(1, 400, size=(n_cells, 11, P)), 0).astype(np.uint16)
    planes[5, :, :] = 0                                        # a listed cell without rows
    planes[:, :, P:] = 777                                     # padding must never be written
    ovf = []
    for c, pl, p, v in ((3, 0, 0, 70000), (3, 1, 0, 65535), (7, 10, P - 1, 1 << 20), (9, 6, 100, 123456), (9, 7, 100, 65536)):
        planes[c, pl, p] = 65535
        ovf.append((c, (pl << 24) | p, v))
    planes[11, 2, 55] = 65535                                   # saturated without a list entry: written as is
    overflow = np.array(ovf, dtype=OVERFLOW_DTYPE)
    cells = [int(c) for c in rng.permutation(n_cells)[:33]]
    for must in (3, 5, 7, 9, 11):
        if must not in cells:
            cells.append(must)
    names = [("BC%05d-1" % c) if c % 3 else ("X" * (1 + c % 40) + "-%d" % c) for c in cells]
    for fname, a, b in (("A", 0, 1), ("T", 6, 7), ("coverage", 10, -1)):
        path = tmp_path / f"output.{fname}.txt.gz"
        rows = write_plane_file(path, planes, P, overflow, a, b, cells, names, level=level, threads=threads)
        got = gzip.open(path).read()
        want = expected_text(planes, P, overflow, a, b, cells, names)
        assert got == want, fname
        assert rows == want.count(b"\n")


def test_empty_and_errors(tmp_path):
    planes = np.zeros((2, 11, pos_pad(100)), np.uint16)
    path = tmp_path / "e.txt.gz"
    assert write_plane_file(path, planes, 100, None, 0, 1, [], []) == 0
    assert gzip.open(path).read() == b""
    assert write_plane_file(path, planes, 100, None, 10, -1, [1, 0], ["b", "a"]) == 0 and gzip.open(path).read() == b""
    with pytest.raises(OSError):
        write_plane_file(path, planes, 100, None, 0, 1, [2], ["x"])                    # cell out of range
    with pytest.raises(OSError):
        write_plane_file(tmp_path / "no_such_dir" / "f.gz", planes, 100, None, 0, 1, [0], ["x"])
    with pytest.raises(ValueError):
        write_plane_file(path, planes, 100, None, 0, 1, [0, 1], ["x"])
