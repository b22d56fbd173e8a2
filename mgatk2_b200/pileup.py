"""`PileupGenerator` — same constructor and methods as the reference's (src/processing/pileup.py:10-154),
computed by the CUDA path. `generate_pileup` takes the reference's per-read objects (`SimpleRead`-like:
reference_start, is_reverse, mapping_quality, query_sequence, query_qualities, cigar, is_paired,
template_length) or a `ReadBatch`, and returns the reference's dict-of-dicts."""
from __future__ import annotations

import numpy as np

from . import _lib
from .batch import ReadBatch
from .engine import PLANE_NAMES, PileupEngine, PileupResult

_ENGINES: dict[tuple, PileupEngine] = {}


def get_engine(device: int = 0, instance: int = 0) -> PileupEngine:
    """One engine (= one C-ABI handle) per device; `instance` separates the handles when a device is listed more than
    once (two host threads must not share a handle)."""
    if (device, instance) not in _ENGINES:
        _ENGINES[(device, instance)] = PileupEngine(device)
    return _ENGINES[(device, instance)]


def reads_to_batch(reads, bc_idx: int = 0) -> ReadBatch:
    """Pack the reference's SimpleRead objects (src/core/config.py:37-49) of ONE cell, in list order."""
    recs = []
    for r in reads:
        seq = r.query_sequence
        if isinstance(seq, (bytes, bytearray)):
            seq = seq.decode("ascii")
        flag = (0x10 if r.is_reverse else 0) | (0x1 if getattr(r, "is_paired", False) else 0) | \
               (0x2 if getattr(r, "is_proper_pair", False) else 0)
        recs.append(dict(pos=int(r.reference_start), flag=flag, mapq=int(r.mapping_quality), seq=seq,
                         qual=np.asarray(r.query_qualities).astype(np.uint8).tolist(),
                         cigar=list(r.cigar or []), tlen=int(getattr(r, "template_length", 0)), bc_idx=bc_idx))
    return ReadBatch.from_records(recs)


def planes_to_dict(planes: np.ndarray, keep: np.ndarray) -> dict[int, dict[str, int]]:
    """Rows of [11, P] exact planes -> the per-position dicts of pileup.py:100-124 for the positions in `keep`."""
    out = {}
    for pos in np.nonzero(keep)[0].tolist():
        v = planes[:, pos]
        d = {"depth": int(v[10]), "tn5_cuts_fwd": int(v[8]), "tn5_cuts_rev": int(v[9])}
        for bi, base in enumerate("ACGT"):
            f, r = int(v[2 * bi]), int(v[2 * bi + 1])
            d[base] = f + r
            d[f"{base}_fwd"] = f
            d[f"{base}_rev"] = r
        out[pos] = d
    return out


def cell_pileup_dict(res: PileupResult, cell: int, raw: bool = False) -> dict[int, dict[str, int]]:
    """The per-cell pileup dict of pileup.py:100-124 (keys 0-based positions) from the dense planes."""
    planes = np.stack([res.plane(k)[cell] for k in range(11)])            # exact (overflow list applied), [11, P]
    keep = planes[10] > 0
    if raw:
        keep |= (planes[8] > 0) | (planes[9] > 0)
    return planes_to_dict(planes, keep)


class PileupGenerator:
    def __init__(self, config, device: int = 0):
        self.config = config
        self.bases = ["A", "C", "G", "T"]
        self.base_to_idx = {"A": 0, "C": 1, "G": 2, "T": 3}
        self.device = device

    def _run(self, batch: ReadBatch, flags: int) -> PileupResult:
        # reads handed over here are already filtered and de-duplicated (readers.py), so no dedup and no gate
        cfg = self.config
        p = _lib.ParamsC(int(cfg.quality.min_baseq), int(cfg.quality.min_mapq), int(cfg.quality.min_distance_from_end),
                         2, float(cfg.quality.max_strand_bias), 0, int(cfg.mito_length), 1,
                         batch.max_read_extent(), flags)
        return get_engine(self.device).run_host(batch, p, overflow_capacity=1 << 16)

    def generate_pileup(self, reads) -> dict[int, dict[str, int]]:
        """pileup.py:18-126. Reads must be one cell's reads sorted by reference_start (BAM order)."""
        if not len(reads):
            return {}
        batch = reads if isinstance(reads, ReadBatch) else reads_to_batch(reads)
        return cell_pileup_dict(self._run(batch, _lib.FLAG_RAW_PILEUP), 0, raw=True)

    def filter_strand_bias(self, pileup: dict[int, dict[str, int]]) -> dict[int, dict[str, int]]:
        """pileup.py:128-154 on the device for a dict generate_pileup returned: the dict's plain integers go into
        32-bit planes (`mgatk_filter_strand_bias_u32_device`), so any depth takes the same device path."""
        import ctypes

        import torch
        if not pileup:
            return {}
        eng = get_engine(self.device)
        P = int(self.config.mito_length)
        ppad = (P + 63) // 64 * 64
        host = np.zeros((1, 11, ppad), np.uint32)
        for pos, d in pileup.items():
            host[0, :, pos] = [d["A_fwd"], d["A_rev"], d["C_fwd"], d["C_rev"], d["G_fwd"], d["G_rev"], d["T_fwd"], d["T_rev"],
                               d["tn5_cuts_fwd"], d["tn5_cuts_rev"], d["depth"]]
        dev = torch.from_numpy(host.view(np.int32)).cuda(self.device)
        rc = eng.lib.mgatk_filter_strand_bias_u32_device(eng.handle, dev.data_ptr(), 1, P,
                                                         float(self.config.quality.max_strand_bias),
                                                         ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream))
        if rc:
            eng._raise(rc)
        planes = dev.cpu().numpy().view(np.uint32)
        return planes_to_dict(planes[0, :, :P], planes[0, 10, :P] > 0)


__all__ = ["PileupGenerator", "cell_pileup_dict", "reads_to_batch", "get_engine", "PLANE_NAMES"]
