#!/usr/bin/env python
"""Stress-scale run of the streamed seam (north_star configs[4] shape: 2x150 bp with soft clips and indels, strand-bias
0.8, distance-from-end 10): `--records` records of `--cells` cells go through `BAMReader.collect_reads_by_barcode` in parts
of `--part` records (device-resident accumulation, prefetch thread, `--devices` GPUs) and every per-cell QC integer, the
counters and the base totals are compared with the oracle on the whole input. Prints one JSON line."""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

ap = argparse.ArgumentParser()
ap.add_argument("--cells", type=int, default=2000)
ap.add_argument("--records", type=int, default=100_000_000)
ap.add_argument("--part", type=int, default=20_000_000)
ap.add_argument("--profile", default="stress150")
ap.add_argument("--devices", default="0")
ap.add_argument("--runs", default="", help="several 'part:devices' runs on the same input, e.g. '20000000:0,1;1000000000:0'")
a = ap.parse_args()

from mgatk2_b200 import BAMReader, PipelineConfig
from mgatk2_b200.synth import make_whitelist, synth_batch
from oracle.oracle import make_params, run_oracle

t0 = time.perf_counter()
batch = synth_batch(a.cells, a.records, a.profile, seed=77)
t_synth = time.perf_counter() - t0
barcodes = make_whitelist(a.cells)
cfg = PipelineConfig(min_baseq=20, min_mapq=30, max_strand_bias=0.8, min_reads_per_cell=1)
cfg.quality.min_distance_from_end = 10
t0 = time.perf_counter()
ora = run_oracle(batch, make_params(a.cells, 20, 30, 10, 0, 0.8, 1, max_read_extent=batch.max_read_extent()),
                 n_threads=os.cpu_count() or 1, dense=False)
t_ora = time.perf_counter() - t0
runs = [r.split(":") for r in a.runs.split(";")] if a.runs else [[str(a.part), a.devices]]
all_ok = True
for part, devs in runs:
    a.part = int(part)
    devices = [int(x) for x in devs.split(",")]
    t0 = time.perf_counter()
    reader = BAMReader("in-memory.bam", cfg, set(barcodes), barcode_list=barcodes, batch=batch, devices=devices,
                       max_batch_records=a.part)
    rbb, stats = reader.collect_reads_by_barcode()
    t_run = time.perf_counter() - t0
    res = rbb.result
    cols = res.columns if res.columns is not None else np.arange(a.cells)
    bad = [f for f in ("n_reads", "n_paired", "sum_depth", "covered", "max_depth", "median_lo", "median_hi")
           if not np.array_equal(res.cell_qc[f], ora.cell_qc[f][cols])]
    if not np.array_equal(res.base_totals, ora.base_totals):
        bad.append("base_totals")
    bad += [k for k in ("total_reads", "filtered_reads", "dup_with_length", "dup_position_only") if res.stats[k] != ora.stats[k]]
    ok = not bad
    if bad:
        print("differing:", bad, {k: (res.stats[k], ora.stats[k]) for k in ("total_reads", "filtered_reads")}, file=sys.stderr)
        for f in bad:
            if f in res.cell_qc.dtype.names:
                w = np.nonzero(res.cell_qc[f] != ora.cell_qc[f][cols])[0]
                print(f, len(w), w[:5], res.cell_qc[f][w[:5]], ora.cell_qc[f][cols][w[:5]], file=sys.stderr)
    print(json.dumps({"records": a.records, "cells": a.cells, "parts_of": a.part, "devices": devices, "streamed": reader.streamed,
                      "input_gb": batch.nbytes() / 1e9, "synth_s": round(t_synth, 1), "seam_s": round(t_run, 2),
                      "records_per_s_through_the_seam": a.records / t_run, "oracle_s": round(t_ora, 1),
                      "kept_reads": res.stats["filtered_reads"], "parity_with_oracle": "OK" if ok else "MISMATCH"}))
    all_ok &= ok
sys.exit(0 if all_ok else 1)
