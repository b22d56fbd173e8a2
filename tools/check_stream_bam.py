#!/usr/bin/env python
"""BAM read in parts (bamio.iter_bam_chrM) -> accumulating device path (PileupEngine.run_stream) against the one-shot
read of the same file through the host ABI: planes, QC rows, counters and base totals must be identical. Needs a GPU."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from mgatk2_b200.bamio import iter_bam_chrM, read_bam_chrM, write_bam
from mgatk2_b200.config import PipelineConfig
from mgatk2_b200.engine import PileupEngine
from mgatk2_b200.synth import synth_batch

cells, n, part = 60, 40_000, 7_000
batch = synth_batch(cells, n, "atac70", seed=4)
barcodes = [f"BC{i:04d}-1" for i in range(cells)]
cfg = PipelineConfig()
wl = {b: i for i, b in enumerate(barcodes)}
with tempfile.TemporaryDirectory() as d:
    bam = os.path.join(d, "s.bam")
    write_bam(bam, batch, barcodes)
    whole, _ = read_bam_chrM(bam, cfg, wl)
    parts = list(iter_bam_chrM(bam, cfg, wl, max_records=part))
eng = PileupEngine(0)
params = cfg.to_params(cells, whole.max_read_extent())
ref = eng.run_host(whole, params, overflow_capacity=1 << 16)
dout = eng.alloc_device_outputs(cells, params.mito_length, max(p.n_records for p in parts), overflow_capacity=1 << 16)
got = eng.run_stream(parts, params, dout)
assert len(parts) > 3
np.testing.assert_array_equal(got.planes, ref.planes)
for f in ref.cell_qc.dtype.names:
    np.testing.assert_array_equal(got.cell_qc[f], ref.cell_qc[f], err_msg=f)
np.testing.assert_array_equal(got.base_totals, ref.base_totals)
assert {k: got.stats[k] for k in ("total_reads", "filtered_reads", "dup_with_length", "dup_position_only")} == \
       {k: ref.stats[k] for k in ("total_reads", "filtered_reads", "dup_with_length", "dup_position_only")}
print(f"streamed BAM parts OK: {len(parts)} parts, {whole.n_records} records, {int(ref.cell_qc['sum_depth'].sum())} counted bases")
