#!/usr/bin/env python
"""torchrun --nproc-per-node N tools/check_multi_gpu.py : one synthetic input sharded by barcode over N GPUs
(mgatk2_b200.multi.run_sharded), per-cell QC rows / counters / base totals compared with the oracle on the whole input,
and every rank's plane columns compared with the oracle's columns. Prints one line per rank."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
if world > 1:
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from mgatk2_b200._lib import ParamsC
from mgatk2_b200.engine import PileupEngine
from mgatk2_b200.multi import run_sharded
from mgatk2_b200.synth import synth_batch
from oracle.oracle import make_params, run_oracle

n_cells = 300
batch = synth_batch(n_cells, 600_000, "atac50", seed=123)
p = make_params(n_cells, max_read_extent=batch.max_read_extent(), max_strand_bias=0.9)
lp = ParamsC(*[getattr(p, f) for f, _ in p._fields_])
eng = PileupEngine(local)
res, cols, combined = run_sharded(batch, lp, eng, rank, world)
ora = run_oracle(batch, p, n_threads=8)
ok = bool(np.array_equal(res.coverage(), ora.coverage[cols]) and np.array_equal(res.counts(), ora.counts[cols])
          and np.array_equal(res.tn5(), ora.tn5[cols]))
if rank == 0:
    for f in ("n_reads", "n_paired", "sum_depth", "covered", "max_depth", "median_lo", "median_hi"):
        ok &= bool(np.array_equal(combined["cell_qc"][f], ora.cell_qc[f]))
    ok &= bool(np.array_equal(combined["base_totals"], ora.base_totals))
    ok &= all(combined["stats"][k] == ora.stats[k] for k in ("total_reads", "stage1_reads", "filtered_reads", "dup_with_length", "dup_position_only"))
print(f"rank {rank}/{world}: {len(cols)} cells, {res.stats['total_reads']} records on this GPU, parity {'OK' if ok else 'MISMATCH'}", flush=True)
if world > 1:
    dist.barrier()
    dist.destroy_process_group()
sys.exit(0 if ok else 1)
