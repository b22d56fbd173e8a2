"""One input, several GPUs of one box: one process per GPU under torchrun, the whitelist partitioned by
record count, every rank running stages 1-6 on its own cells through the C ABI - no collective on the counting path.
Afterwards the only cross-cell quantity, the reference-allele base totals (`int64[16569, 4]`, writers.py:220-222,
345-349), is summed with one NCCL all-reduce over NVLink, and the per-cell QC rows / counters are gathered to rank 0.
Plane columns stay on their rank (each rank writes its own cells)."""
from __future__ import annotations

import numpy as np

from .sharding import assign_cells, combine_columns, combine_stats, shard_batch


def run_sharded(batch, params, engine, rank: int, world: int, balance: bool = True):
    """Run this rank's shard of `batch` (every rank holds or can read the whole batch; routing keeps file order).
    Returns (local PileupResult, global columns of the local cells, combined dict on rank 0 else None)."""
    import torch
    import torch.distributed as dist
    from ._lib import ParamsC
    n_cells = int(params.n_cells)
    weights = np.bincount(batch.bc_idx[(batch.bc_idx >= 0) & (batch.bc_idx < n_cells)], minlength=n_cells) if balance else None
    owner = assign_cells(n_cells, world, weights)
    sub, cols, unowned = shard_batch(batch, owner, rank)
    local = ParamsC.from_buffer_copy(params)
    local.n_cells = len(cols)
    local.max_read_extent = max(int(params.max_read_extent), 1)
    res = engine.run_host(sub, local, overflow_capacity=1 << 16)
    totals = torch.from_numpy(res.base_totals.copy()).to(torch.device("cuda", engine.device))
    if world > 1:
        dist.all_reduce(totals)                           # NCCL: the only reduction of the path
    gathered = [None] * world
    payload = (cols, res.cell_qc.copy(), dict(res.stats), unowned)
    if world > 1:
        dist.all_gather_object(gathered, payload)
    else:
        gathered = [payload]
    combined = None
    if rank == 0:
        combined = {"cell_qc": combine_columns(n_cells, [g[0] for g in gathered], [g[1] for g in gathered]),
                    "stats": combine_stats([g[2] for g in gathered], sum(g[3] for g in gathered)),
                    "base_totals": totals.cpu().numpy(), "owner": owner}
    return res, cols, combined
