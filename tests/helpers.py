"""Shared comparison helpers: golden fixture loading and dense-result equality."""
from __future__ import annotations

import glob
import os

import numpy as np

from mgatk2_b200.batch import ReadBatch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = sorted(glob.glob(os.path.join(ROOT, "tests", "golden", "*.npz")))
PARAM_KEYS = ("min_baseq", "min_mapq", "min_distance_from_end", "dedup_mode", "max_strand_bias", "min_reads_per_cell")


def golden_ids():
    return [os.path.basename(p)[:-4] for p in GOLDEN]


def load_golden(path):
    d = np.load(path)
    batch = ReadBatch.from_npz_dict(d)
    params = {k: d["param_" + k].item() for k in PARAM_KEYS}
    return d, batch, [str(x) for x in d["barcodes"]], params


def derived_qc(cell_qc, min_reads_per_cell, mito_length):
    """Host-side float64 QC exactly as processors.py:33-39 computes it."""
    n = cell_qc["n_reads"].astype(np.int64)
    alive = (n >= max(1, min_reads_per_cell)) & (cell_qc["sum_depth"] > 0)
    frag = np.where(cell_qc["n_paired"] > 0, n // 2, n)
    with np.errstate(divide="ignore", invalid="ignore"):
        mean = cell_qc["sum_depth"].astype(np.float64) / cell_qc["covered"].astype(np.float64)
    breadth = cell_qc["covered"].astype(np.float64) / mito_length
    return alive, frag, mean, breadth


def assert_matches_golden(d, params, counts, tn5, coverage, cell_qc, stats, mito_length=16569):
    """counts [C,P,4,2], tn5 [C,P,2], coverage [C,P] unsaturated; cell_qc structured; stats dict."""
    np.testing.assert_array_equal(counts, d["exp_counts"])
    np.testing.assert_array_equal(tn5, d["exp_tn5"])
    np.testing.assert_array_equal(coverage, d["exp_coverage"])
    total, filtered, n_barcodes, dup_len, dup_pos = (int(x) for x in d["exp_stats"])
    assert stats["total_reads"] == total
    assert stats["filtered_reads"] == filtered
    assert stats["dup_with_length"] == dup_len
    assert stats["dup_position_only"] == dup_pos
    assert int((cell_qc["n_reads"] > 0).sum()) == n_barcodes
    np.testing.assert_array_equal(cell_qc["n_reads"].astype(np.int64), d["exp_n_reads_in"])
    alive, frag, mean, breadth = derived_qc(cell_qc, params["min_reads_per_cell"], mito_length)
    np.testing.assert_array_equal(alive.astype(np.uint8), d["exp_alive"])
    a = alive
    np.testing.assert_array_equal(cell_qc["n_reads"][a].astype(np.float64), d["exp_qc"][a, 0])
    np.testing.assert_array_equal(frag[a].astype(np.float64), d["exp_qc"][a, 1])
    np.testing.assert_array_equal(mean[a], d["exp_qc"][a, 2])          # float64, bit-exact
    np.testing.assert_array_equal(breadth[a], d["exp_breadth"][a])
    # writers.py:187-193 depth statistics follow from the coverage plane
    cov = d["exp_coverage"]
    for c in np.nonzero(a)[0]:
        dep = cov[c][cov[c] > 0]
        assert cell_qc["max_depth"][c] == dep.max()
        assert (cell_qc["median_lo"][c] + cell_qc["median_hi"][c]) / 2 == np.median(dep)
        assert cell_qc["covered"][c] == len(dep) and cell_qc["sum_depth"][c] == dep.sum()


def assert_result_equals_oracle(res, ora, dense=True):
    """GPU PileupResult vs OracleResult: bit-exact on every integer the path produces."""
    for k in ("total_reads", "stage1_reads", "filtered_reads", "dup_with_length", "dup_position_only", "n_empty_seq"):
        assert res.stats[k] == ora.stats[k], (k, res.stats[k], ora.stats[k])
    for f in ("n_reads", "n_paired", "sum_depth", "covered", "max_depth", "median_lo", "median_hi"):   # medians exact above 65535 too
        np.testing.assert_array_equal(res.cell_qc[f], ora.cell_qc[f], err_msg=f)
    np.testing.assert_array_equal(res.base_totals, ora.base_totals)
    if dense:
        np.testing.assert_array_equal(res.coverage(), ora.coverage)
        np.testing.assert_array_equal(res.tn5(), ora.tn5)
        np.testing.assert_array_equal(res.counts(), ora.counts)
        # padding beyond mito_length stays zero
        assert not res.planes[:, :, res.mito_length:].any()
