"""How the seam drives the C ABI: which whitelist barcodes get a column, which GPU owns which columns, one batch or a
stream of batches (reference contract: src/core/pipeline.py:76-113, one BAM in, one result out).

* **Columns** (`ColumnPlan`): the reference only does set membership on the whitelist and materialises the barcodes it
  actually saw (readers.py:75-76,104-111). Planes are therefore allocated for the whitelist barcodes OBSERVED on the
  contig, not for the whole list (a full 10x whitelist has 737 k entries = 269 GB of planes); `columns[k]` is the
  whitelist index of local column k.
* **Devices**: every quantity of the path is per cell, so the local columns are cut into contiguous ranges of about
  equal record count, one per GPU; each GPU runs stages 1-6 on the records of its range through its own handle
  (`mgatk_pileup_host` on one host thread per GPU; the calls release the GIL) and writes straight into its slice of the
  shared host result. No data-path collective; the base totals (`int64[P, 4]`) and the counters are added on the host.
* **Streams**: inputs above `max_batch_records` arrive in parts cut on reference_start borders
  (`bamio.iter_bam_chrM`), decoded on a prefetch thread while the GPUs count the part before; the parts add up in
  device-resident planes (`MGATK_FLAG_ACCUMULATE`) and the finish pass applies what needs the totals.
"""
from __future__ import annotations

import os
import queue
import threading
from concurrent.futures import ThreadPoolExecutor
from dataclasses import dataclass

import numpy as np

from . import _lib
from .batch import ReadBatch
from .engine import CELL_QC_DTYPE, OVERFLOW_DTYPE, STATS_FIELDS, PileupResult, pos_pad
from .pileup import get_engine

DEFAULT_MAX_BATCH_RECORDS = int(os.environ.get("MGATK_MAX_BATCH_RECORDS", 64_000_000))
# streamed inputs: above this many bytes of planes per GPU for the whole whitelist, the barcodes are counted first and
# only observed ones get a column
MAX_STREAM_PLANE_BYTES = int(os.environ.get("MGATK_MAX_STREAM_PLANE_BYTES", 48 << 30))


@dataclass
class ColumnPlan:
    columns: np.ndarray        # whitelist index of local column k, ascending
    local_of: np.ndarray       # whitelist index -> local column, -1 for barcodes without a column
    cuts: list                 # device d owns local columns [cuts[d], cuts[d + 1])

    @property
    def n_columns(self) -> int:
        return len(self.columns)

    def device_of(self, local: np.ndarray) -> np.ndarray:
        return np.searchsorted(np.asarray(self.cuts[1:-1]), local, side="right").astype(np.int32)


def _engine(devices: list, d: int):
    """The engine of entry d of `devices`; a device listed twice gets two handles (threads never share one)."""
    return get_engine(devices[d], instance=devices[:d].count(devices[d]))


def usable(batch: ReadBatch, n_whitelist: int) -> np.ndarray:
    """Records whose barcode has a whitelist index (the flag filter is left to the device)."""
    return (batch.bc_idx >= 0) & (batch.bc_idx < n_whitelist)


def plan_columns(n_whitelist: int, records_per_barcode: np.ndarray | None, n_devices: int) -> ColumnPlan:
    """Columns for the barcodes with records (all of the whitelist when the counts are not known in advance), cut into
    `n_devices` contiguous ranges of about equal record count."""
    if records_per_barcode is None:
        columns = np.arange(n_whitelist, dtype=np.int64)
        weights = np.ones(n_whitelist, np.int64)
    else:
        columns = np.nonzero(records_per_barcode > 0)[0].astype(np.int64)
        weights = records_per_barcode[columns].astype(np.int64)
    local_of = np.full(max(n_whitelist, 1), -1, np.int32)
    local_of[columns] = np.arange(len(columns), dtype=np.int32)
    n_devices = max(1, min(n_devices, max(len(columns), 1)))
    cuts = [0]
    if len(columns):
        csum = np.cumsum(weights)
        for d in range(1, n_devices):
            k = int(np.searchsorted(csum, csum[-1] * d / n_devices, side="left")) + 1
            cuts.append(min(max(k, cuts[-1]), len(columns)))
    cuts.append(len(columns))
    while len(cuts) < n_devices + 1:
        cuts.append(len(columns))
    return ColumnPlan(columns, local_of, cuts)


def route(batch: ReadBatch, plan: ColumnPlan, n_whitelist: int):
    """Per device: the records of its columns in file order, bc_idx renumbered to the device's own columns. Returns
    ([ReadBatch per device], records that belong to no column)."""
    ok = usable(batch, n_whitelist)
    local = np.where(ok, plan.local_of[np.clip(batch.bc_idx, 0, len(plan.local_of) - 1)], -1).astype(np.int32)
    n_dev = len(plan.cuts) - 1
    unowned = int((local < 0).sum())
    if n_dev == 1:
        sub = ReadBatch(pos=batch.pos, tlen=batch.tlen, flag=batch.flag, mapq=batch.mapq, bc_idx=local, l_seq=batch.l_seq,
                        n_cigar=batch.n_cigar, blob_off=batch.blob_off, blob=batch.blob)   # views; only bc_idx is new
        return [sub], 0        # the device sees (and counts) every record itself
    dev = plan.device_of(local)
    subs = []
    for d in range(n_dev):
        idx = np.nonzero((local >= 0) & (dev == d))[0]
        sub = batch.take(idx)
        sub.bc_idx = (local[idx] - plan.cuts[d]).astype(np.int32)
        subs.append(sub)
    return subs, unowned


class _SharedOutputs:
    """One pinned host result for all local columns; every device gets views of its column range."""

    def __init__(self, n_columns: int, mito_length: int, n_devices: int, overflow_capacity: int):
        eng = get_engine_for_alloc()
        self.full = eng.alloc_host_outputs(n_columns, mito_length, overflow_capacity=0)
        self.per_device = []
        for _ in range(n_devices):
            small = eng.alloc_host_outputs(0, mito_length, overflow_capacity=overflow_capacity)
            self.per_device.append(small)

    def view(self, d: int, c0: int, c1: int) -> dict:
        small = self.per_device[d]
        return {"planes": self.full["planes"][c0:c1], "cell_qc": self.full["cell_qc"][c0:c1], "stats": small["stats"],
                "base_totals": small["base_totals"], "overflow": small["overflow"]}


_alloc_engine = None


def get_engine_for_alloc():
    global _alloc_engine
    if _alloc_engine is None:
        _alloc_engine = get_engine(0)
    return _alloc_engine


def _combine(parts: list, plan: ColumnPlan, shared_planes, shared_qc, mito_length: int, min_reads: int, unowned: int) -> PileupResult:
    stats = {k: 0 for k in STATS_FIELDS}
    totals = np.zeros((mito_length, 4), np.int64)
    ovf, stage_ms, launches = [], {}, 0
    for d, r in enumerate(parts):
        if r is None:
            continue
        for k in STATS_FIELDS:
            stats[k] += int(r.stats[k])
        totals += r.base_totals
        if len(r.overflow):
            o = np.array(r.overflow, dtype=OVERFLOW_DTYPE)
            o["cell"] += plan.cuts[d]
            ovf.append(o)
        for k, v in r.stage_ms.items():
            stage_ms[k] = max(stage_ms.get(k, 0.0), v)
        launches += r.launches
    stats["total_reads"] += unowned
    stats["error_bits"] = 0
    res = PileupResult(shared_planes, shared_qc, stats, totals, np.concatenate(ovf) if ovf else np.zeros(0, OVERFLOW_DTYPE),
                       mito_length, min_reads, stage_ms, launches)
    res.columns = plan.columns
    return res


def run_one_batch(batch: ReadBatch, config, n_whitelist: int, devices: list, overflow_capacity: int = 1 << 16) -> PileupResult:
    """Stages 1-6 of one in-memory batch on `devices`. Returns a PileupResult over the observed columns (`.columns`)."""
    ok = usable(batch, n_whitelist)
    counts = np.bincount(batch.bc_idx[ok], minlength=max(n_whitelist, 1))
    plan = plan_columns(n_whitelist, counts, len(devices))
    n_dev = len(plan.cuts) - 1
    subs, unowned = route(batch, plan, n_whitelist)
    extent = batch.max_read_extent()
    P = int(config.mito_length)
    if n_dev == 1:
        params = config.to_params(plan.n_columns, extent)
        res = get_engine(devices[0]).run_host(subs[0], params, overflow_capacity=overflow_capacity)
        res.columns = plan.columns
        return res
    shared = _SharedOutputs(plan.n_columns, P, n_dev, overflow_capacity)

    def work(d):
        c0, c1 = plan.cuts[d], plan.cuts[d + 1]
        params = config.to_params(c1 - c0, max(extent, 1))
        return _engine(devices, d).run_host(subs[d], params, out=shared.view(d, c0, c1))

    with ThreadPoolExecutor(max_workers=n_dev) as pool:
        parts = list(pool.map(work, range(n_dev)))
    res = _combine(parts, plan, shared.full["planes"], shared.full["cell_qc"], P, int(config.min_reads_per_cell), unowned)
    res._keep = shared
    return res


class _Prefetch:
    """Runs an iterator one item ahead on a thread (BAM parts are decoded while the GPUs count the part before)."""

    _END = object()

    def __init__(self, it, depth: int = 1):
        self.q = queue.Queue(maxsize=depth)
        self.err = None
        self.t = threading.Thread(target=self._run, args=(it,), daemon=True)
        self.t.start()

    def _run(self, it):
        try:
            for x in it:
                self.q.put(x)
        except BaseException as e:          # handed to the consumer
            self.err = e
        self.q.put(self._END)

    def __iter__(self):
        while True:
            x = self.q.get()
            if x is self._END:
                if self.err is not None:
                    raise self.err
                return
            yield x


def run_stream(parts, config, n_whitelist: int, devices: list, records_per_barcode: np.ndarray | None = None,
               overflow_capacity: int = 1 << 16, first_seen: np.ndarray | None = None) -> PileupResult:
    """Stages 1-6 over batches cut on reference_start borders, in file order, accumulating in device-resident planes.
    `records_per_barcode` (when known, e.g. from a barcode scan) restricts the columns to observed barcodes; otherwise
    every whitelist entry gets a column. `first_seen` (int64[n_whitelist], filled with the global index of the first
    usable record of every barcode) is updated when given."""
    import torch
    plan = plan_columns(n_whitelist, records_per_barcode, len(devices))
    n_dev = len(plan.cuts) - 1
    P = int(config.mito_length)
    engines = [_engine(devices, d) for d in range(n_dev)]
    state = [None] * n_dev          # per device: (params, DeviceOutputs), created with the first part
    unowned_total, seen = 0, 0

    def begin(d, n_records, extent):
        eng = engines[d]
        c0, c1 = plan.cuts[d], plan.cuts[d + 1]
        params = config.to_params(c1 - c0, max(extent, 1))
        with torch.cuda.device(eng.device):
            dout = eng.alloc_device_outputs(c1 - c0, P, n_records, overflow_capacity=overflow_capacity, max_read_extent=extent)
            eng.stream_begin(params, dout)
        return params, dout

    def step(args):
        d, sub = args
        eng = engines[d]
        with torch.cuda.device(eng.device):
            if state[d] is None:
                state[d] = begin(d, max(sub.n_records, 1), sub.max_read_extent())
            params, dout = state[d]
            eng.stream_add(sub, params, dout)
        return None

    pool = ThreadPoolExecutor(max_workers=n_dev)
    try:
        for batch in _Prefetch(parts):
            if first_seen is not None:
                ok = usable(batch, n_whitelist) & ((batch.flag & 0x904) == 0)
                cells, first = np.unique(batch.bc_idx[ok], return_index=True)
                idx = np.nonzero(ok)[0][first] + seen
                first_seen[cells] = np.minimum(first_seen[cells], idx)
            seen += batch.n_records
            subs, unowned = route(batch, plan, n_whitelist)
            unowned_total += unowned
            list(pool.map(step, list(enumerate(subs))))

        def finish(d):
            eng = engines[d]
            with torch.cuda.device(eng.device):
                if state[d] is None:
                    state[d] = begin(d, 1, 1)
                params, dout = state[d]
                return eng.stream_finish(params, dout)

        parts_res = list(pool.map(finish, range(n_dev)))
    finally:
        pool.shutdown(wait=True)
    if n_dev == 1:
        res = parts_res[0]
        res.columns = plan.columns
        return res
    planes = np.concatenate([r.planes for r in parts_res]) if plan.n_columns else np.zeros((0, _lib.N_PLANES, pos_pad(P)), np.uint16)
    qc = np.concatenate([r.cell_qc for r in parts_res]) if plan.n_columns else np.zeros(0, CELL_QC_DTYPE)
    return _combine(parts_res, plan, planes, qc, P, int(config.min_reads_per_cell), unowned_total)
