"""Data layer of the reference's QC report (src/analysis/report.py:84-130,164-262) from the reference's own committed
HDF5 files: the vectorised reductions against the report's per-position Python loops restated here."""
import os

import numpy as np

from mgatk2_b200.h5lite import H5Reader
from mgatk2_b200.report import report_inputs

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "tests", "golden", "ref_hdf5")


def test_report_inputs_from_the_reference_files():
    cpath, mpath = os.path.join(REF, "counts.h5"), os.path.join(REF, "metadata.h5")
    got = report_inputs((cpath, mpath))
    c, m = H5Reader(cpath), H5Reader(mpath)
    fwd = c.objects["tn5_cuts_fwd"].read().sum(axis=1).astype(np.int64)          # report.py:88-90
    rev = c.objects["tn5_cuts_rev"].read().sum(axis=1).astype(np.int64)
    np.testing.assert_array_equal(got["tn5_cuts_fwd"], fwd)
    np.testing.assert_array_equal(got["tn5_cuts_rev"], rev)
    np.testing.assert_array_equal(got["read_start_sites"], m.objects["coverage"].read().sum(axis=1).astype(np.int64))
    ref = [x.decode() for x in m.objects["reference"].read()]
    total = fwd + rev
    want = {a + b: 0 for a in "ACGT" for b in "ACGT"}                             # report.py:182-195, the loop as written there
    for pos in range(len(total) - 1):
        if total[pos] > 0:
            d = ref[pos] + ref[pos + 1]
            if d in want:
                want[d] += int(total[pos])
    assert got["dinucleotide_counts"] == want and sum(want.values()) > 0
    n = sum(want.values())
    assert all(abs(got["dinucleotide_percent"][k] - v / n * 100) < 1e-12 for k, v in want.items())
    md, gc = m.objects["mean_depth"].read(), m.objects["genome_coverage"].read()
    mask = (md > 0) & (gc > 0)
    np.testing.assert_array_equal(got["mean_depth"], md[mask])
    np.testing.assert_array_equal(got["genome_coverage"], gc[mask])


def test_report_inputs_from_a_device_result():
    from mgatk2_b200.engine import CELL_QC_DTYPE, OVERFLOW_DTYPE, PileupResult
    rng = np.random.default_rng(2)
    P, ppad = 500, 512
    planes = np.zeros((4, 11, ppad), np.uint16)
    planes[:, :10, :P] = rng.integers(0, 4, (4, 10, P))
    planes[:, 10, :P] = planes[:, :8, :P].sum(1)
    planes[3] = 0
    qc = np.zeros(4, CELL_QC_DTYPE)
    qc["n_reads"] = [5, 5, 5, 0]
    qc["sum_depth"] = planes[:, 10, :P].sum(1)
    qc["covered"] = (planes[:, 10, :P] > 0).sum(1)
    totals = np.stack([planes[:, 2 * b, :P].astype(np.int64).sum(0) + planes[:, 2 * b + 1, :P].sum(0) for b in range(4)], 1)
    res = PileupResult(planes, qc, {}, totals, np.zeros(0, OVERFLOW_DTYPE), P, 1)
    got = report_inputs(res)
    np.testing.assert_array_equal(got["tn5_cuts_fwd"], planes[:3, 8, :P].sum(0))
    ref = res.reference_alleles().astype(str)
    tot = planes[:3, 8, :P].sum(0).astype(np.int64) + planes[:3, 9, :P].sum(0)
    want = {}
    for pos in range(P - 1):
        if tot[pos] > 0 and ref[pos] in "ACGT" and ref[pos + 1] in "ACGT":
            want[ref[pos] + ref[pos + 1]] = want.get(ref[pos] + ref[pos + 1], 0) + int(tot[pos])
    assert {k: v for k, v in got["dinucleotide_counts"].items() if v} == want
    assert len(got["mean_depth"]) == 3
