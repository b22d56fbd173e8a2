#!/usr/bin/env python
"""The whole drop-in on one synthetic BAM: whitelist file + BAM -> native ingest -> GPU pileup through the host ABI ->
text outputs (MtDNAPipeline.run, the mirror of core/pipeline.py:76-181), with the seconds each phase took.
Needs a GPU. Usage: bench_pipeline.py [cells] [records] [gzip_level]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pathlib import Path
from mgatk2_b200.bamio import write_bam
from mgatk2_b200.config import PipelineConfig
from mgatk2_b200.pipeline import MtDNAPipeline
from mgatk2_b200.synth import synth_batch

cells = int(sys.argv[1]) if len(sys.argv) > 1 else 500
n = int(sys.argv[2]) if len(sys.argv) > 2 else 2_000_000
if len(sys.argv) > 3:
    os.environ["MGATK_TXT_GZIP_LEVEL"] = sys.argv[3]
batch = synth_batch(cells, n, "atac50", seed=2)
barcodes = [f"ACGTACGTACGT{i:04d}-1" for i in range(cells)]
with tempfile.TemporaryDirectory() as d:
    bam = os.path.join(d, "s.bam")
    t0 = time.perf_counter(); write_bam(bam, batch, barcodes); t_bam = time.perf_counter() - t0
    for rep in range(2):                                   # the first run loads the libraries and creates the CUDA context
        out = Path(d) / f"out{rep}"
        p = MtDNAPipeline(bam, barcodes, out, PipelineConfig())
        t0 = time.perf_counter(); res = p.run(); t_run = time.perf_counter() - t0
    size = sum(f.stat().st_size for f in (out / "output").iterdir())
    bam_mb = os.path.getsize(bam) / 1e6
print(f"{cells} cells x {n} records ({bam_mb:.0f} MB BAM written by the Python test writer in {t_bam:.0f} s): "
      f"run {t_run:.2f} s = ingest {p.timings['ingest_s']:.2f} + GPU through the host ABI {p.timings['gpu_host_abi_s']:.2f} "
      f"+ text outputs {p.timings['write_s']:.2f} (gzip level {os.environ.get('MGATK_TXT_GZIP_LEVEL', '9')}, {size / 1e6:.0f} MB); "
      f"{res['cells_passed_qc']} cells passed, {n / t_run / 1e6:.2f} M records/s end to end")
