"""TEST INFRASTRUCTURE — ctypes wrapper around oracle/mgatk2_oracle.c (not product code).

`run_oracle(batch, params)` executes the plain-C restatement of the reference's
readers.py / processors.py / pileup.py path and returns dense unsaturated arrays in the
reference's natural layout.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libmgatk2_oracle.so")


class ParamsC(ctypes.Structure):
    _fields_ = [
        ("min_baseq", ctypes.c_int32), ("min_mapq", ctypes.c_int32),
        ("min_distance_from_end", ctypes.c_int32), ("dedup_mode", ctypes.c_int32),
        ("max_strand_bias", ctypes.c_double), ("min_reads_per_cell", ctypes.c_int32),
        ("mito_length", ctypes.c_int32), ("n_cells", ctypes.c_int32), ("max_read_extent", ctypes.c_int32),
        ("flags", ctypes.c_int32),
    ]


class OracleOutputsC(ctypes.Structure):
    _fields_ = [("counts", ctypes.c_void_p), ("tn5", ctypes.c_void_p), ("coverage", ctypes.c_void_p),
                ("cell_qc", ctypes.c_void_p), ("stats", ctypes.c_void_p), ("base_totals", ctypes.c_void_p),
                ("keep", ctypes.c_void_p)]


CELL_QC_DTYPE = np.dtype([("n_reads", "<u4"), ("n_paired", "<u4"), ("sum_depth", "<u8"), ("covered", "<u4"),
                          ("max_depth", "<u4"), ("median_lo", "<u4"), ("median_hi", "<u4")])
STATS_FIELDS = ("total_reads", "stage1_reads", "filtered_reads", "dup_with_length", "dup_position_only",
                "n_empty_seq", "n_overflow", "error_bits")


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "mgatk2_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.mgatk_oracle_run.restype = ctypes.c_int
        _lib.mgatk_oracle_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
    return _lib


@dataclass
class OracleResult:
    counts: np.ndarray | None      # uint32 [n_cells, P, 4, 2]
    tn5: np.ndarray | None         # uint32 [n_cells, P, 2]
    coverage: np.ndarray | None    # uint32 [n_cells, P]
    cell_qc: np.ndarray            # CELL_QC_DTYPE [n_cells]
    stats: dict
    base_totals: np.ndarray | None # int64 [P, 4]
    keep: np.ndarray               # uint8 [n_records]


def make_params(n_cells, min_baseq=20, min_mapq=30, min_distance_from_end=5, dedup_mode=0,
                max_strand_bias=1.0, min_reads_per_cell=1, mito_length=16569, max_read_extent=0, flags=0) -> ParamsC:
    return ParamsC(int(min_baseq), int(min_mapq), int(min_distance_from_end), int(dedup_mode),
                   float(max_strand_bias), int(min_reads_per_cell), int(mito_length), int(n_cells),
                   int(max_read_extent), int(flags))


def run_oracle(batch, params: ParamsC, n_threads: int = 1, dense: bool = True) -> OracleResult:
    C, P = params.n_cells, params.mito_length
    counts = np.zeros((C, P, 4, 2), np.uint32) if dense else None
    tn5 = np.zeros((C, P, 2), np.uint32) if dense else None
    cov = np.zeros((C, P), np.uint32) if dense else None
    qc = np.zeros(C, CELL_QC_DTYPE)
    stats = np.zeros(8, np.uint64)
    bt = np.zeros((P, 4), np.int64)
    keep = np.zeros(max(batch.n_records, 1), np.uint8)
    out = OracleOutputsC(counts.ctypes.data if dense else None, tn5.ctypes.data if dense else None,
                         cov.ctypes.data if dense else None, qc.ctypes.data, stats.ctypes.data,
                         bt.ctypes.data, keep.ctypes.data)
    cb = batch.as_c()
    rc = lib().mgatk_oracle_run(ctypes.addressof(params), ctypes.addressof(cb), ctypes.addressof(out), int(n_threads))
    if rc != 0:
        raise RuntimeError(f"oracle returned {rc}")
    return OracleResult(counts, tn5, cov, qc, {k: int(v) for k, v in zip(STATS_FIELDS, stats)}, bt,
                        keep[: batch.n_records])
