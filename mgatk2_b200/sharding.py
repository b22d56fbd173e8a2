"""Barcode sharding across the GPUs of one box: every quantity on the counting path is per
cell, so the whitelist is partitioned and each rank runs stages 1-6 on its own records; no collective on
the data path. Only the global counters, the per-cell QC rows and the reference-allele totals are combined
afterwards (`combine_*`, tiny all-reduce / gather)."""
from __future__ import annotations

import numpy as np

from .batch import ReadBatch


def assign_cells(n_cells: int, world: int, weights: np.ndarray | None = None) -> np.ndarray:
    """rank of every whitelist index. Without weights: index % world. With per-cell record counts: greedy
    longest-processing-time balancing."""
    if weights is None:
        return (np.arange(n_cells) % world).astype(np.int32)
    order = np.argsort(-np.asarray(weights), kind="stable")
    load = np.zeros(world, np.int64)
    owner = np.zeros(n_cells, np.int32)
    for c in order.tolist():
        r = int(np.argmin(load))
        owner[c] = r
        load[r] += int(weights[c])
    return owner


def shard_batch(batch: ReadBatch, owner: np.ndarray, rank: int) -> tuple[ReadBatch, np.ndarray, int]:
    """Records of this rank's cells (file order kept) with bc_idx renumbered to local columns.
    Returns (local batch, global index of every local column, records dropped before stage 1 that rank 0
    must still count in total_reads)."""
    n_cells = len(owner)
    mine = np.nonzero(owner == rank)[0]
    local_of = np.full(n_cells, -1, np.int32)
    local_of[mine] = np.arange(len(mine), dtype=np.int32)
    valid = (batch.bc_idx >= 0) & (batch.bc_idx < n_cells)
    loc = np.where(valid, local_of[np.clip(batch.bc_idx, 0, max(n_cells - 1, 0))], -1) if n_cells else np.full(batch.n_records, -1)
    keep = np.nonzero(loc >= 0)[0]
    sub = batch.take(keep)
    sub.bc_idx = loc[keep].astype(np.int32)
    unowned = int((~valid).sum()) if rank == 0 else 0       # records without a usable barcode belong to nobody
    return sub, mine, unowned


def combine_stats(per_rank: list[dict], unowned_records: int) -> dict:
    out = {k: sum(s[k] for s in per_rank) for k in per_rank[0]}
    out["total_reads"] += unowned_records
    return out


def combine_columns(n_cells: int, columns: list[np.ndarray], per_rank_rows: list[np.ndarray]) -> np.ndarray:
    """Scatter per-rank per-cell rows (e.g. QC rows) back to whitelist order."""
    out = np.zeros(n_cells, dtype=per_rank_rows[0].dtype)
    for cols, rows in zip(columns, per_rank_rows):
        out[cols] = rows
    return out
