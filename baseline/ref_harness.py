"""Runs the UNMODIFIED reference (ollieeknight/mgatk2, pip-installed into baseline/_ref) on a synthetic batch.

Timed path = what `MtDNAPipeline.run()` does between opening the BAM and handing results to a writer
(`core/pipeline.py:76-103`): `BAMReader.collect_reads_by_barcode()` (filter + dedup, one Python loop) and
`CellProcessor.process_cells_progressive()` (pileup, strand-bias filter, per-cell QC; the reference itself decides
between its process pool and its sequential loop, `processing/processors.py:88-111`). `pysam`, `h5py` and
`matplotlib` are not in this image: `baseline/stubs/` provides import stand-ins, the fake `pysam.AlignmentFile.fetch()`
hands out read objects decoded before the clock starts. Used by `bench.py --impl reference` and by the `cpu_baseline`
leg of the bench line; nothing in the product path imports this file.
"""
from __future__ import annotations

import logging
import os
import subprocess
import sys
import tempfile
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref")
STUBS = os.path.join(HERE, "stubs")
REFERENCE_SRC = "/root/reference"


def install(force: bool = False) -> bool:
    """pip-install the reference into baseline/_ref (build container only; the directory travels to the GPU box)."""
    if os.path.isdir(os.path.join(REF_DIR, "processing")) and not force:
        return True
    if not os.path.isdir(REFERENCE_SRC):
        return False
    with tempfile.TemporaryDirectory() as tmp:     # the build writes egg-info into the source tree: use a copy
        src = os.path.join(tmp, "reference")
        subprocess.run(["cp", "-r", REFERENCE_SRC, src], check=True)
        subprocess.run([sys.executable, "-m", "pip", "install", "--no-index", "--no-build-isolation", "--no-deps", "-q",
                        "--find-links", "/opt/wheelhouse", "--target", REF_DIR, "--upgrade", src], check=True)
    return os.path.isdir(os.path.join(REF_DIR, "processing"))


def available() -> bool:
    return os.path.isdir(os.path.join(REF_DIR, "processing"))


def _activate():
    os.environ.setdefault("TQDM_DISABLE", "1")     # the reference wraps its loops in tqdm progress bars
    # Workers of the reference's process pool are spawned: they re-run the parent's main module (bench.py) and then
    # unpickle their first task, which imports `processing` before `core` - the reference's circular import fails in that
    # order (its own CLI imports `core` first). bench.py sees this variable in a worker and activates the same way.
    os.environ["MGATK_REF_ACTIVATE"] = "1"
    for p in (STUBS, REF_DIR):
        if p not in sys.path:
            sys.path.insert(0, p)       # spawned pool workers inherit sys.path (multiprocessing preparation data)
    for name in ("pysam", "h5py", "matplotlib"):
        __import__(name)
    import core  # noqa: F401  (before `processing`: the reference has a circular import otherwise)
    logging.getLogger().setLevel(logging.WARNING)


def prepare(batch, barcodes):
    """Decode the batch into pysam-like read objects (untimed) and register them as a fake BAM. Returns its path."""
    _activate()
    import pysam
    reads = []
    for i in range(batch.n_records):
        rec = batch.record(i)
        b = rec["bc_idx"]
        reads.append(pysam.AlignedSegment(rec, barcodes[b] if b >= 0 else (None if b == -1 else "NOTINWHITELIST-1")))
    fd, path = tempfile.mkstemp(suffix=".bam")
    os.close(fd)
    pysam.REGISTRY[path] = (reads, ("chr1", "chrM"))
    return path


def release(path):
    import pysam
    pysam.REGISTRY.pop(path, None)
    try:
        os.unlink(path)
    except OSError:
        pass


def run_once(path, barcodes, *, min_baseq=20, min_mapq=30, min_distance_from_end=5, dedup_mode=0, max_strand_bias=1.0,
             min_reads_per_cell=1, n_cores=None):
    """One pass of the reference's path over the registered reads. Returns (seconds, stats dict, n cells, mode)."""
    _activate()
    from core.config import PipelineConfig
    from processing.processors import CellProcessor
    from processing.readers import BAMReader
    cfg = PipelineConfig(min_baseq=min_baseq, min_mapq=min_mapq, max_strand_bias=max_strand_bias,
                         skip_deduplication=(dedup_mode == 2), use_fragment_length_dedup=(dedup_mode == 0),
                         min_reads_per_cell=min_reads_per_cell, n_cores=n_cores or (os.cpu_count() or 1))
    cfg.quality.min_distance_from_end = min_distance_from_end
    with tempfile.TemporaryDirectory() as out:
        t0 = time.perf_counter()
        reader = BAMReader(path, cfg, set(barcodes))
        reads_by_barcode, stats = reader.collect_reads_by_barcode()
        n_in = len(reads_by_barcode)
        avg = sum(len(r) for r in reads_by_barcode.values()) / max(n_in, 1)
        mode = "sequential (the reference's own rule above 2500 reads/cell)" if avg > 2500 else f"process pool, {cfg.performance.n_cores} workers"
        results = CellProcessor(cfg, out).process_cells_progressive(reads_by_barcode, None)
        dt = time.perf_counter() - t0
    return dt, stats, len(results), mode
