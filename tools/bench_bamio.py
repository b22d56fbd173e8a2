#!/usr/bin/env python
"""Host-side throughput of the native BAM ingest (csrc/bamio.cpp): records/s and MB/s of compressed BAM for a synthetic
chrM BAM. CPU only. Usage: bench_bamio.py [records] [threads]"""
import os, sys, tempfile, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from mgatk2_b200.bamio import read_bam_chrM, write_bam
from mgatk2_b200.config import PipelineConfig
from mgatk2_b200.synth import synth_batch

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1_000_000
threads = int(sys.argv[2]) if len(sys.argv) > 2 else min(16, os.cpu_count() or 1)
cells = 500
batch = synth_batch(cells, n, "atac50", seed=1)
barcodes = [f"BC{i:06d}-1" for i in range(cells)]
with tempfile.TemporaryDirectory() as d:
    p = os.path.join(d, "s.bam")
    t0 = time.perf_counter(); write_bam(p, batch, barcodes); tw = time.perf_counter() - t0
    size = os.path.getsize(p)
    wl = {b: i for i, b in enumerate(barcodes)}
    read_bam_chrM(p, PipelineConfig(), wl, threads=threads)
    t0 = time.perf_counter(); got, _ = read_bam_chrM(p, PipelineConfig(), wl, threads=threads); tr = time.perf_counter() - t0
print(f"{n} records, BAM {size / 1e6:.1f} MB: native read {n / tr / 1e6:.2f} M records/s ({size / tr / 1e6:.0f} MB/s compressed, "
      f"{threads} inflate threads); python writer {n / tw / 1e3:.0f} k records/s")
