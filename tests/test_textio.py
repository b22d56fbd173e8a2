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


@pytest.mark.parametrize("name", ["synth_run_default", "synth_tenx"])
def test_dense_writer_reproduces_reference_files(tmp_path, name):
    """DenseTextWriter (native rows + gzip members, Python for the small tables) on the result the oracle gives for a
    golden input: every file the reference's own IncrementalTextWriter wrote for those records (tests/golden, generated
    by running the reference) comes out byte for byte after gunzip. CPU only: the oracle stands in for the GPU here;
    tests/test_gpu_dropin.py does the same through the CUDA path."""
    from mgatk2_b200 import PipelineConfig
    from mgatk2_b200.engine import N_PLANES, PileupResult
    from mgatk2_b200.readers import ReadsByBarcode
    from mgatk2_b200.writers import DenseTextWriter
    from oracle.oracle import make_params, run_oracle
    from tests.helpers import load_golden
    d, batch, barcodes, params = load_golden(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    if "txt_A" not in d:
        pytest.skip("golden without text outputs")
    n_cells, P = len(barcodes), 16569
    ora = run_oracle(batch, make_params(n_cells, params["min_baseq"], params["min_mapq"], params["min_distance_from_end"],
                                        params["dedup_mode"], params["max_strand_bias"], params["min_reads_per_cell"],
                                        max_read_extent=batch.max_read_extent()), n_threads=4)
    planes = np.zeros((n_cells, N_PLANES, pos_pad(P)), np.uint16)
    for b in range(4):
        for s in range(2):
            planes[:, 2 * b + s, :P] = np.minimum(ora.counts[:, :, b, s], 65535)
    planes[:, 8, :P] = np.minimum(ora.tn5[:, :, 0], 65535)
    planes[:, 9, :P] = np.minimum(ora.tn5[:, :, 1], 65535)
    planes[:, 10, :P] = np.minimum(ora.coverage, 65535)
    assert ora.coverage.max() < 65535                                  # (no overflow list needed for these inputs)
    res = PileupResult(planes, ora.cell_qc, dict(ora.stats), ora.base_totals, np.zeros(0, OVERFLOW_DTYPE), P,
                       params["min_reads_per_cell"])
    ok = ((batch.flag & 0x904) == 0) & (batch.bc_idx >= 0)
    cells, first = np.unique(batch.bc_idx[ok], return_index=True)      # first-seen order, as BAMReader hands it on
    order = cells[np.argsort(first, kind="stable")]
    order = order[ora.cell_qc["n_reads"][order] > 0]
    cfg = PipelineConfig(min_baseq=params["min_baseq"], min_mapq=params["min_mapq"], max_strand_bias=params["max_strand_bias"],
                         skip_deduplication=params["dedup_mode"] == 2, use_fragment_length_dedup=params["dedup_mode"] == 0,
                         min_reads_per_cell=params["min_reads_per_cell"], sequential=True)
    writer = DenseTextWriter(tmp_path, cfg, barcodes, compresslevel=6)
    results = writer.write_result(ReadsByBarcode(barcodes, order, res), cfg)
    writer.finalize(tmp_path / "qc")
    assert len(results) == int(d["exp_alive"].sum())
    for base in ("A", "C", "G", "T", "coverage"):
        assert gzip.open(tmp_path / "output" / f"output.{base}.txt.gz").read() == d[f"txt_{base}"].tobytes(), base
    assert (tmp_path / "output" / "output.depthTable.txt").read_bytes() == d["txt_depthTable"].tobytes()
    assert (tmp_path / "output" / "chrM_refAllele.txt").read_bytes() == d["txt_refAllele"].tobytes()
    assert (tmp_path / "qc" / "cell_stats.csv").read_bytes() == d["txt_cell_stats"].tobytes()
