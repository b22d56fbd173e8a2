"""Host-side driver of the C-ABI pileup path.

`PileupEngine.run_host()` is the call the drop-in classes make (readers.py / processors.py mirrors):
host structure-of-arrays in, dense planes and QC rows out, through `mgatk_pileup_host`.
`upload()` + `run_device()` keep everything in HBM (torch tensors own the memory, the library only
enqueues kernels on the current stream) — the resident-input form the throughput metric is quoted on.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass, field

import numpy as np

from . import _lib
from ._lib import N_PLANES, OutputsC, ParamsC
from .batch import FIELDS, ReadBatch
from .exceptions import PileupKernelError

CELL_QC_DTYPE = np.dtype([("n_reads", "<u4"), ("n_paired", "<u4"), ("sum_depth", "<u8"), ("covered", "<u4"),
                          ("max_depth", "<u4"), ("median_lo", "<u4"), ("median_hi", "<u4")])
OVERFLOW_DTYPE = np.dtype([("cell", "<i4"), ("plane_pos", "<u4"), ("value", "<u4")])
STATS_FIELDS = ("total_reads", "stage1_reads", "filtered_reads", "dup_with_length", "dup_position_only",
                "n_empty_seq", "n_overflow", "error_bits")
PLANE_NAMES = ("A_fwd", "A_rev", "C_fwd", "C_rev", "G_fwd", "G_rev", "T_fwd", "T_rev",
               "tn5_cuts_fwd", "tn5_cuts_rev", "coverage")


def pos_pad(p: int) -> int:
    return (int(p) + 63) // 64 * 64


@dataclass
class PileupResult:
    """Dense result of one batch. `planes` is uint16 [n_cells, 11, pos_pad], saturated at 65535 like the
    reference's HDF5 datasets; `overflow` lists the exact values of saturated entries."""

    planes: np.ndarray
    cell_qc: np.ndarray
    stats: dict
    base_totals: np.ndarray
    overflow: np.ndarray
    mito_length: int
    min_reads_per_cell: int = 1
    stage_ms: dict = field(default_factory=dict)
    launches: int = 0
    columns: np.ndarray | None = None      # whitelist index of every row of `planes` / `cell_qc` (None: row = whitelist index)

    def plane(self, k: int, exact: bool = True) -> np.ndarray:
        """Plane k as uint32 [n_cells, P]; exact=True patches saturated entries from the overflow list."""
        out = self.planes[:, k, : self.mito_length].astype(np.uint32)
        if exact and len(self.overflow):
            o = self.overflow[(self.overflow["plane_pos"] >> 24) == k]
            out[o["cell"], o["plane_pos"] & 0xFFFFFF] = o["value"]
        return out

    def counts(self) -> np.ndarray:
        """uint32 [n_cells, P, 4, 2] — the reference's base_counts layout (pileup.py:24) per cell."""
        return np.stack([self.plane(k) for k in range(8)], axis=-1).reshape(len(self.planes), self.mito_length, 4, 2)

    def tn5(self) -> np.ndarray:
        return np.stack([self.plane(8), self.plane(9)], axis=-1)

    def coverage(self) -> np.ndarray:
        return self.plane(10)

    def alive(self) -> np.ndarray:
        """Cells that yield a result in process_barcode_worker (processors.py:22-31)."""
        n = self.cell_qc["n_reads"].astype(np.int64)
        return (n >= max(1, self.min_reads_per_cell)) & (self.cell_qc["sum_depth"] > 0)

    def reference_alleles(self) -> np.ndarray:
        """Per-position argmax over A,C,G,T of the cross-cell totals, first wins, 'N' if all zero
        (writers.py:345-349,493-500)."""
        best = np.argmax(self.base_totals, axis=1)
        ref = np.array(list("ACGT"))[best]
        ref[self.base_totals.max(axis=1) <= 0] = "N"
        return ref


class PileupEngine:
    """One engine per GPU (one `mgatk_handle`)."""

    def __init__(self, device: int = 0):
        self.lib = _lib.load()                      # raises ExtensionMissingError: no fallback
        h = ctypes.c_void_p()
        rc = self.lib.mgatk_create(ctypes.byref(h), int(device))
        if rc:
            raise PileupKernelError(rc, self.lib.mgatk_status_string(rc).decode())
        self.handle, self.device = h, int(device)

    def close(self):
        if getattr(self, "handle", None):
            self.lib.mgatk_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _raise(self, rc: int):
        msg = self.lib.mgatk_last_error(self.handle).decode() or self.lib.mgatk_status_string(rc).decode()
        raise PileupKernelError(rc, msg)

    def stage_times(self) -> dict:
        names = (ctypes.c_char_p * 16)()
        ms = (ctypes.c_float * 16)()
        n = ctypes.c_int(0)
        self.lib.mgatk_last_stage_times(self.handle, names, ms, ctypes.byref(n))
        return {names[i].decode(): float(ms[i]) for i in range(n.value)}

    def launch_count(self) -> int:
        return int(self.lib.mgatk_last_launch_count(self.handle))

    # ------------------------------------------------------------------ host buffers in, host buffers out
    def alloc_host_outputs(self, n_cells: int, mito_length: int, overflow_capacity: int = 4096, pinned: bool = True):
        import torch
        pin = pinned and torch.cuda.is_available()

        def buf(shape, dtype):
            nbytes = int(np.prod(shape)) * np.dtype(dtype).itemsize
            t = torch.empty(max(nbytes, 1), dtype=torch.uint8, pin_memory=pin)
            return t, t.numpy()[:nbytes].view(dtype).reshape(shape)

        keep = []
        out = {}
        for name, shape, dt in (("planes", (n_cells, N_PLANES, pos_pad(mito_length)), np.uint16),
                                ("cell_qc", (n_cells,), CELL_QC_DTYPE), ("stats", (8,), np.uint64),
                                ("base_totals", (mito_length, 4), np.int64),
                                ("overflow", (overflow_capacity,), OVERFLOW_DTYPE)):
            t, a = buf(shape, dt)
            keep.append(t)
            out[name] = a
        out["_keep"] = keep
        return out

    def run_host(self, batch: ReadBatch, params: ParamsC, out: dict | None = None,
                 overflow_capacity: int = 4096) -> PileupResult:
        if out is None:
            out = self.alloc_host_outputs(params.n_cells, params.mito_length, overflow_capacity)
        oc = OutputsC(out["planes"].ctypes.data, out["cell_qc"].ctypes.data, out["stats"].ctypes.data,
                      out["base_totals"].ctypes.data, out["overflow"].ctypes.data if len(out["overflow"]) else None,
                      len(out["overflow"]))
        bc = batch.as_c()
        rc = self.lib.mgatk_pileup_host(self.handle, ctypes.byref(params), ctypes.byref(bc), ctypes.byref(oc))
        if rc:
            self._raise(rc)
        stats = {k: int(v) for k, v in zip(STATS_FIELDS, out["stats"])}
        return PileupResult(out["planes"], out["cell_qc"], stats, out["base_totals"],
                            out["overflow"][: stats["n_overflow"]], params.mito_length, params.min_reads_per_cell,
                            self.stage_times(), self.launch_count())

    def submit_host(self, batch: ReadBatch, params: ParamsC, out: dict) -> int:
        """First half of run_host (mgatk_pileup_host_submit): enqueue upload, kernels and download, return a ticket.
        Two tickets may be in flight; `batch` and `out` must stay untouched until wait_host(ticket)."""
        oc = OutputsC(out["planes"].ctypes.data, out["cell_qc"].ctypes.data, out["stats"].ctypes.data,
                      out["base_totals"].ctypes.data, out["overflow"].ctypes.data if len(out["overflow"]) else None,
                      len(out["overflow"]))
        bc = batch.as_c()
        ticket = ctypes.c_int64(-1)
        rc = self.lib.mgatk_pileup_host_submit(self.handle, ctypes.byref(params), ctypes.byref(bc), ctypes.byref(oc),
                                               ctypes.byref(ticket))
        if rc:
            self._raise(rc)
        if not hasattr(self, "_in_flight"):
            self._in_flight = {}
        self._in_flight[ticket.value] = (batch, out, params.mito_length, params.min_reads_per_cell, self.launch_count())
        return ticket.value

    def wait_host(self, ticket: int) -> PileupResult:
        batch, out, mito_length, min_reads, launches = getattr(self, "_in_flight", {}).pop(ticket, (None,) * 5)
        rc = self.lib.mgatk_pileup_host_wait(self.handle, int(ticket))
        if rc:
            self._raise(rc)
        stats = {k: int(v) for k, v in zip(STATS_FIELDS, out["stats"])}
        return PileupResult(out["planes"], out["cell_qc"], stats, out["base_totals"], out["overflow"][: stats["n_overflow"]],
                            mito_length, min_reads, {}, launches)

    def run_host_many(self, batches, params_of, outs):
        """Pipelined run_host over an iterable of batches: yields (index, PileupResult) in order. `params_of(batch)`
        gives the parameters of a batch, `outs` is a list of TWO host output sets (alloc_host_outputs) that are reused
        alternately, so a result must be consumed before the next but one is yielded."""
        pending = []
        for i, batch in enumerate(batches):
            if len(pending) == 2:
                j, t = pending.pop(0)
                yield j, self.wait_host(t)
            pending.append((i, self.submit_host(batch, params_of(batch), outs[i % 2])))
        for j, t in pending:
            yield j, self.wait_host(t)

    # ------------------------------------------------------------------ device-resident path
    def upload(self, batch: ReadBatch, pinned_src: dict | None = None):
        """Copy a batch into HBM (torch tensors own the memory). Returns a DeviceBatch."""
        import torch
        dev = torch.device("cuda", self.device)
        tensors = {}
        for name, _ in FIELDS:
            src = pinned_src[name] if pinned_src else torch.from_numpy(getattr(batch, name))
            tensors[name] = src.to(dev, non_blocking=True) if src.numel() else torch.empty(0, dtype=src.dtype, device=dev)
        return DeviceBatch(batch.n_records, int(len(batch.blob)), tensors)

    def alloc_device_outputs(self, n_cells: int, mito_length: int, n_records: int, overflow_capacity: int = 4096,
                             max_read_extent: int | None = None):
        """Device outputs + workspace for batches of up to `n_records` records. `max_read_extent` selects the slot
        layout the workspace is sized for (32 bytes per read up to an extent of 56, more beyond); without it the
        workspace covers any extent (144 bytes per read)."""
        import torch
        dev = torch.device("cuda", self.device)
        extent = 1 << 20 if max_read_extent is None else max(int(max_read_extent), 1)
        ws_bytes = int(self.lib.mgatk_workspace_bytes(int(n_records), int(n_cells), extent))
        if ws_bytes < 0:
            raise PileupKernelError(8, "n_records / n_cells outside limits")
        return DeviceOutputs(
            planes=torch.empty((max(n_cells, 1), N_PLANES, pos_pad(mito_length)), dtype=torch.uint16, device=dev),
            cell_qc=torch.empty(max(n_cells, 1) * CELL_QC_DTYPE.itemsize, dtype=torch.uint8, device=dev),
            stats=torch.empty(8, dtype=torch.int64, device=dev),
            base_totals=torch.empty((mito_length, 4), dtype=torch.int64, device=dev),
            overflow=torch.empty(max(overflow_capacity, 1) * OVERFLOW_DTYPE.itemsize, dtype=torch.uint8, device=dev),
            overflow_capacity=overflow_capacity,
            workspace=torch.empty(max(ws_bytes, 256), dtype=torch.uint8, device=dev), n_cells=n_cells)

    def run_device(self, dbatch: "DeviceBatch", params: ParamsC, dout: "DeviceOutputs") -> None:
        """Enqueue stages 1-6 on torch's current stream. No host synchronisation."""
        import torch
        bc = dbatch.as_c()
        oc = OutputsC(dout.planes.data_ptr(), dout.cell_qc.data_ptr(), dout.stats.data_ptr(),
                      dout.base_totals.data_ptr(), dout.overflow.data_ptr() if dout.overflow_capacity else None,
                      dout.overflow_capacity)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.lib.mgatk_pileup_device(self.handle, ctypes.byref(params), ctypes.byref(bc), ctypes.byref(oc),
                                          dout.workspace.data_ptr(), dout.workspace.numel(), ctypes.c_void_p(stream))
        if rc:
            self._raise(rc)

    # ------------------------------------------------------------------ streaming (inputs larger than HBM)
    def _outputs_c(self, dout: "DeviceOutputs") -> OutputsC:
        return OutputsC(dout.planes.data_ptr(), dout.cell_qc.data_ptr(), dout.stats.data_ptr(), dout.base_totals.data_ptr(),
                        dout.overflow.data_ptr() if dout.overflow_capacity else None, dout.overflow_capacity)

    def stream_begin(self, params: ParamsC, dout: "DeviceOutputs") -> None:
        """Zero the resident outputs of a stream of batches (mgatk_stream_begin_device)."""
        import torch
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        oc = self._outputs_c(dout)
        rc = self.lib.mgatk_stream_begin_device(self.handle, ctypes.byref(params), ctypes.byref(oc), stream)
        if rc:
            self._raise(rc)
        dout.stream_launches = 0

    def stream_add(self, batch: ReadBatch, params: ParamsC, dout: "DeviceOutputs") -> None:
        """One batch of a stream (cut on reference_start borders, file order): upload, stages 1-6 with
        MGATK_FLAG_ACCUMULATE - the batch's raw counts are added to the resident planes."""
        import torch
        if batch.n_records == 0:
            return
        acc = ParamsC.from_buffer_copy(params)
        acc.flags = int(params.flags) | _lib.FLAG_ACCUMULATE
        acc.max_read_extent = max(int(batch.max_read_extent()), 1)
        need = int(self.lib.mgatk_workspace_bytes(int(batch.n_records), int(params.n_cells), int(acc.max_read_extent)))
        if need < 0:
            raise PileupKernelError(8, "n_records / n_cells outside limits")
        if need > dout.workspace.numel():                # a later part may be larger (or longer reads) than the first
            dout.workspace = None
            dout.workspace = torch.empty(need, dtype=torch.uint8, device=torch.device("cuda", self.device))
        db = self.upload(batch)
        self.run_device(db, acc, dout)
        dout.stream_launches = getattr(dout, "stream_launches", 0) + self.launch_count()
        torch.cuda.synchronize(self.device)              # the batch's device buffers are released before the next upload

    def stream_finish(self, params: ParamsC, dout: "DeviceOutputs") -> PileupResult:
        """What needs the totals of all batches (mgatk_stream_finish_device): cell gate, strand-bias filter, coverage /
        Tn5 gating, depth statistics, base totals, medians; then the download."""
        import torch
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        oc = self._outputs_c(dout)
        rc = self.lib.mgatk_stream_finish_device(self.handle, ctypes.byref(params), ctypes.byref(oc), stream)
        if rc:
            self._raise(rc)
        launches = getattr(dout, "stream_launches", 0) + self.launch_count()
        res = self.download(dout, params)
        res.launches = launches
        return res

    def run_stream(self, batches, params: ParamsC, dout: "DeviceOutputs") -> PileupResult:
        """Batches cut on reference_start borders (`ReadBatch.split_on_start_borders`), in file order: every batch adds its
        raw counts to the resident planes (MGATK_FLAG_ACCUMULATE), the finish pass applies the cell gate, the strand-bias
        filter, coverage / Tn5 gating and the depth statistics. Equals the one-batch result for any depth (cells whose
        entries pass 65535 on the way get 32-bit carry planes inside the handle)."""
        self.stream_begin(params, dout)
        for batch in batches:
            self.stream_add(batch, params, dout)
        return self.stream_finish(params, dout)

    def download(self, dout: "DeviceOutputs", params: ParamsC) -> PileupResult:
        """Synchronise and bring a device result back as a PileupResult (checks the device error bits)."""
        import torch
        torch.cuda.synchronize(self.device)
        stats_np = dout.stats.cpu().numpy().view(np.uint64)
        rc = self.lib.mgatk_check_stats(stats_np.ctypes.data)
        if rc:
            raise PileupKernelError(rc, self.lib.mgatk_status_string(rc).decode())
        stats = {k: int(v) for k, v in zip(STATS_FIELDS, stats_np)}
        n = dout.n_cells
        qc = dout.cell_qc.cpu().numpy().view(CELL_QC_DTYPE)[:n]
        ovf = dout.overflow.cpu().numpy().view(OVERFLOW_DTYPE)[: min(stats["n_overflow"], dout.overflow_capacity)]
        return PileupResult(dout.planes[:n].cpu().numpy(), qc, stats, dout.base_totals.cpu().numpy(), ovf,
                            params.mito_length, params.min_reads_per_cell, self.stage_times(), self.launch_count())


@dataclass
class DeviceBatch:
    n_records: int
    blob_bytes: int
    tensors: dict

    def as_c(self):
        from .batch import MgatkBatchC
        c = MgatkBatchC()
        c.n_records, c.blob_bytes = self.n_records, self.blob_bytes
        for name, _ in FIELDS:
            setattr(c, name, self.tensors[name].data_ptr())
        return c

    def nbytes(self) -> int:
        return int(sum(t.numel() * t.element_size() for t in self.tensors.values()))


@dataclass
class DeviceOutputs:
    planes: "object"
    cell_qc: "object"
    stats: "object"
    base_totals: "object"
    overflow: "object"
    overflow_capacity: int
    workspace: "object"
    n_cells: int
