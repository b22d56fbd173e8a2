"""The numbers behind the reference's QC report (src/analysis/report.py:84-130,164-262), without the plots: per-position
Tn5 cut sites summed over cells, read-start / coverage totals, the dinucleotide context of Tn5 insertions and the
depth-versus-coverage points. Inputs are what the reference reads them from - the saturated uint16 datasets of
counts.h5 / metadata.h5 - taken either from a `PileupResult` (device planes) or from the files themselves (`h5lite`).
matplotlib is not in this image and the HTML page is out of scope; these reductions are its data layer."""
from __future__ import annotations

import numpy as np

from .engine import PileupResult

DINUCLEOTIDES = [a + b for a in "ACGT" for b in "ACGT"]     # sorted, as report.py:196 sorts its keys


def _tables(source):
    """(tn5_fwd, tn5_rev, coverage) summed over cells as int64[P], reference alleles as str array, mean depth and genome
    coverage per cell."""
    if isinstance(source, PileupResult):
        P = source.mito_length
        alive = source.alive()
        pl = source.planes[alive]
        s = lambda k: pl[:, k, :P].sum(axis=0, dtype=np.int64)
        qc = source.cell_qc[alive]
        with np.errstate(divide="ignore", invalid="ignore"):
            mean = (qc["sum_depth"] / np.maximum(qc["covered"], 1)).astype(np.float32)
        return s(8), s(9), s(10), source.reference_alleles().astype(str), mean, (qc["covered"] / P * 100).astype(np.float32)
    counts_file, metadata_file = source
    from .h5lite import H5Reader
    c, m = H5Reader(counts_file), H5Reader(metadata_file)
    s = lambda f, name: f.objects[name].read().sum(axis=1, dtype=np.int64)
    ref = np.array([x.decode() for x in m.objects["reference"].read().tolist()])
    return (s(c, "tn5_cuts_fwd"), s(c, "tn5_cuts_rev"), s(m, "coverage"), ref, m.objects["mean_depth"].read(),
            m.objects["genome_coverage"].read())


def report_inputs(source) -> dict:
    """`source`: a PileupResult, or (counts.h5 path, metadata.h5 path)."""
    tn5_fwd, tn5_rev, coverage, ref, mean_depth, genome_cov = _tables(source)
    total = tn5_fwd + tn5_rev                                            # report.py:180
    code = np.full(len(ref), -1, np.int64)
    for k, b in enumerate("ACGT"):
        code[ref == b] = k
    pair = code[:-1] * 4 + code[1:]                                      # report.py:190-195: context = (ref[pos], ref[pos + 1])
    ok = (code[:-1] >= 0) & (code[1:] >= 0) & (total[:-1] > 0)
    dinuc = np.bincount(pair[ok], weights=total[:-1][ok].astype(np.float64), minlength=16).astype(np.int64)
    n = int(dinuc.sum())
    keep = (mean_depth > 0) & (genome_cov > 0)                           # report.py:268-270
    return {"positions": np.arange(1, len(tn5_fwd) + 1), "tn5_cuts_fwd": tn5_fwd, "tn5_cuts_rev": tn5_rev,
            "read_start_sites": coverage,
            "dinucleotide_counts": dict(zip(DINUCLEOTIDES, dinuc.tolist())),
            "dinucleotide_percent": {d: (c / n * 100 if n else 0.0) for d, c in zip(DINUCLEOTIDES, dinuc.tolist())},
            "mean_depth": mean_depth[keep], "genome_coverage": genome_cov[keep]}
