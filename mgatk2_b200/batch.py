"""Structure-of-arrays read batch: the host side of the C-ABI boundary.

One `ReadBatch` holds every record `fetch(chrM)` would yield, in BAM (coordinate) order —
the loop variable of reference src/processing/readers.py:87-93 — decoded into the flat
arrays `include/mgatk2_b200.h: mgatk_batch` describes. What the reference materialises per
kept read as a `SimpleRead` object (src/core/config.py:37-49, readers.py:153-163) is here
one row across nine arrays plus a 16-byte-aligned cigar|seq|qual blob (the same contiguous
region a BAM record carries), so the host parser is a memcpy per record.
"""
from __future__ import annotations

import ctypes
from dataclasses import dataclass

import numpy as np

SEQ_NT16 = "=ACMGRSVTWYHKDBN"  # BAM 4-bit code -> IUPAC letter (SAM spec; what pysam.query_sequence decodes)
_NT16_CODE = {c: i for i, c in enumerate(SEQ_NT16)}
_NT16_CODE.update({c.lower(): i for i, c in enumerate(SEQ_NT16)})
CIGAR_REF_OPS = (0, 2, 3, 7, 8)  # M D N = X consume the reference


class MgatkBatchC(ctypes.Structure):
    """ctypes mirror of `mgatk_batch` (include/mgatk2_b200.h)."""

    _fields_ = [
        ("n_records", ctypes.c_int64),
        ("pos", ctypes.c_void_p),
        ("tlen", ctypes.c_void_p),
        ("flag", ctypes.c_void_p),
        ("mapq", ctypes.c_void_p),
        ("bc_idx", ctypes.c_void_p),
        ("l_seq", ctypes.c_void_p),
        ("n_cigar", ctypes.c_void_p),
        ("blob_off", ctypes.c_void_p),
        ("blob", ctypes.c_void_p),
        ("blob_bytes", ctypes.c_int64),
    ]


FIELDS = (
    ("pos", np.int32),
    ("tlen", np.int32),
    ("flag", np.uint16),
    ("mapq", np.uint8),
    ("bc_idx", np.int32),
    ("l_seq", np.uint16),
    ("n_cigar", np.uint16),
    ("blob_off", np.uint32),
    ("blob", np.uint8),
)


@dataclass
class ReadBatch:
    pos: np.ndarray
    tlen: np.ndarray
    flag: np.ndarray
    mapq: np.ndarray
    bc_idx: np.ndarray
    l_seq: np.ndarray
    n_cigar: np.ndarray
    blob_off: np.ndarray
    blob: np.ndarray

    def __post_init__(self):
        n = len(self.pos)
        for name, dt in FIELDS:
            a = np.ascontiguousarray(getattr(self, name), dtype=dt)
            setattr(self, name, a)
            if name != "blob" and len(a) != n:
                raise ValueError(f"ReadBatch.{name} has {len(a)} entries, expected {n}")
        if len(self.blob) % 16:
            raise ValueError("ReadBatch.blob must be a multiple of 16 bytes")

    # ------------------------------------------------------------------ basic views
    @property
    def n_records(self) -> int:
        return int(len(self.pos))

    def nbytes(self) -> int:
        return int(sum(getattr(self, f).nbytes for f, _ in FIELDS))

    def blob_units(self) -> np.ndarray:
        """Size of every record's blob in 16-byte units."""
        nb = 4 * self.n_cigar.astype(np.int64) + (self.l_seq.astype(np.int64) + 1) // 2 + self.l_seq
        return (nb + 15) // 16

    def cigar_words(self, k: int) -> tuple[np.ndarray, np.ndarray]:
        """(record indices having > k cigar ops, their k-th BAM cigar word)."""
        idx = np.nonzero(self.n_cigar > k)[0]
        w32 = self.blob.view(np.uint32)
        return idx, w32[self.blob_off[idx].astype(np.int64) * 4 + k]

    def reference_span(self) -> np.ndarray:
        span = np.zeros(self.n_records, dtype=np.int64)
        for k in range(int(self.n_cigar.max()) if self.n_records else 0):
            idx, w = self.cigar_words(k)
            op = w & 0xF
            span[idx] += np.where(np.isin(op, CIGAR_REF_OPS), (w >> 4).astype(np.int64), 0)
        return span

    def max_read_extent(self) -> int:
        """Upper bound the kernels size their position ring with: max(reference span, l_seq)."""
        if self.n_records == 0:
            return 1
        return int(max(int(self.reference_span().max()), int(self.l_seq.max()), 1))

    def is_sorted(self) -> bool:
        return bool(np.all(self.pos[1:] >= self.pos[:-1]))

    # ------------------------------------------------------------------ C view
    def as_c(self, ptr=None) -> MgatkBatchC:
        """ctypes struct over these arrays (or over `ptr(name)` device pointers)."""
        c = MgatkBatchC()
        c.n_records = self.n_records
        c.blob_bytes = int(len(self.blob))
        for name, _ in FIELDS:
            a = getattr(self, name)
            setattr(c, name, ptr(name) if ptr else a.ctypes.data)
        return c

    # ------------------------------------------------------------------ subsetting (barcode sharding)
    def split_on_start_borders(self, n_parts: int) -> list:
        """Cut a coordinate-sorted batch into about `n_parts` consecutive batches whose borders fall between different
        reference_start values: all candidates for a duplicate of a read share its start (readers.py:118-150), so dedup
        never needs state across such batches."""
        n = self.n_records
        if n == 0 or n_parts <= 1:
            return [self]
        cuts = [0]
        for k in range(1, n_parts):
            i = max(cuts[-1], n * k // n_parts)
            while 0 < i < n and self.pos[i] == self.pos[i - 1]:
                i += 1
            if cuts[-1] < i < n:
                cuts.append(i)
        cuts.append(n)
        return [self.slice(a, b) for a, b in zip(cuts[:-1], cuts[1:])]      # views when the blobs lie in record order

    def slice(self, a: int, b: int) -> "ReadBatch":
        """Records [a, b) of a batch whose blobs lie in record order (every ingest batch): views, no gather."""
        a, b = max(0, int(a)), min(self.n_records, int(b))
        if b <= a:
            return self.take(np.zeros(0, np.int64))
        units = self.blob_units()
        off = self.blob_off.astype(np.int64)
        if not np.array_equal(off[a + 1: b], off[a: b - 1] + units[a: b - 1]):
            return self.take(np.arange(a, b))
        lo, hi = int(off[a]) * 16, (int(off[b - 1]) + int(units[b - 1])) * 16
        return ReadBatch(pos=self.pos[a:b], tlen=self.tlen[a:b], flag=self.flag[a:b], mapq=self.mapq[a:b],
                         bc_idx=self.bc_idx[a:b], l_seq=self.l_seq[a:b], n_cigar=self.n_cigar[a:b],
                         blob_off=(off[a:b] - off[a]).astype(np.uint32), blob=self.blob[lo:hi])

    @staticmethod
    def concat(parts: list) -> "ReadBatch":
        """Batches one after the other (blobs packed in record order in every part)."""
        parts = [p for p in parts if p.n_records]
        if not parts:
            return ReadBatch.from_records([])
        if len(parts) == 1:
            return parts[0]
        packed = []
        for p in parts:
            units = p.blob_units()
            off = p.blob_off.astype(np.int64)
            ok = int(off[0]) == 0 and np.array_equal(off[1:], off[:-1] + units[:-1]) and len(p.blob) == int(units.sum()) * 16
            packed.append(p if ok else p.take(np.arange(p.n_records)))
        shift, offs = 0, []
        for p in packed:
            offs.append(p.blob_off.astype(np.int64) + shift)
            shift += len(p.blob) // 16
        cat = lambda f: np.concatenate([getattr(p, f) for p in packed])
        return ReadBatch(pos=cat("pos"), tlen=cat("tlen"), flag=cat("flag"), mapq=cat("mapq"), bc_idx=cat("bc_idx"),
                         l_seq=cat("l_seq"), n_cigar=cat("n_cigar"), blob_off=np.concatenate(offs).astype(np.uint32),
                         blob=cat("blob"))

    def take(self, index: np.ndarray) -> "ReadBatch":
        """Records `index` (kept in the given order) with a freshly packed blob."""
        index = np.asarray(index, dtype=np.int64)
        units = self.blob_units()[index]
        new_off = np.zeros(len(index), dtype=np.int64)
        if len(index):
            np.cumsum(units[:-1], out=new_off[1:])
        total = int(units.sum())
        src16 = self.blob.view("V16")
        dst16 = np.zeros(total, dtype="V16")
        old_off = self.blob_off[index].astype(np.int64)
        for u in range(int(units.max()) if len(index) else 0):
            m = units > u
            dst16[new_off[m] + u] = src16[old_off[m] + u]
        return ReadBatch(
            pos=self.pos[index], tlen=self.tlen[index], flag=self.flag[index], mapq=self.mapq[index],
            bc_idx=self.bc_idx[index], l_seq=self.l_seq[index], n_cigar=self.n_cigar[index],
            blob_off=new_off.astype(np.uint32), blob=dst16.view(np.uint8),
        )

    # ------------------------------------------------------------------ record-level construction / decoding
    @staticmethod
    def from_records(records: list[dict]) -> "ReadBatch":
        """Pack a list of dict records (tests, small inputs).

        Each record: pos, flag, mapq, tlen, bc_idx, seq (str, IUPAC), qual (sequence of ints or
        None -> 30s), cigar (list of (op, len) tuples, pysam `cigartuples` convention).
        """
        n = len(records)
        pos = np.zeros(n, np.int32); tlen = np.zeros(n, np.int32); flag = np.zeros(n, np.uint16)
        mapq = np.zeros(n, np.uint8); bc = np.zeros(n, np.int32); l_seq = np.zeros(n, np.uint16)
        n_cig = np.zeros(n, np.uint16); off = np.zeros(n, np.uint32)
        chunks = []
        cursor = 0
        for i, r in enumerate(records):
            seq = r["seq"]
            L = len(seq)
            qual = r.get("qual")
            if qual is None:
                qual = [30] * L
            if len(qual) != L:
                raise ValueError("qual/seq length mismatch")
            cig = r.get("cigar") or []
            words = np.array([(int(ln) << 4) | int(op) for op, ln in cig], dtype=np.uint32)
            codes = np.array([_NT16_CODE[ch] for ch in seq] + ([0] if L % 2 else []), dtype=np.uint8)
            packed = (codes[0::2] << 4) | codes[1::2]
            raw = words.tobytes() + packed.tobytes() + np.asarray(qual, dtype=np.uint8).tobytes()
            raw += b"\0" * (-len(raw) % 16)
            pos[i], tlen[i], flag[i], mapq[i] = r["pos"], r.get("tlen", 0), r["flag"], r.get("mapq", 60)
            bc[i], l_seq[i], n_cig[i], off[i] = r.get("bc_idx", 0), L, len(cig), cursor // 16
            chunks.append(raw)
            cursor += len(raw)
        blob = np.frombuffer(b"".join(chunks), dtype=np.uint8).copy() if chunks else np.zeros(0, np.uint8)
        return ReadBatch(pos, tlen, flag, mapq, bc, l_seq, n_cig, off, blob)

    def record(self, i: int) -> dict:
        """Decode record i back to the dict form `from_records` takes."""
        L = int(self.l_seq[i]); nc = int(self.n_cigar[i])
        base = int(self.blob_off[i]) * 16
        words = self.blob[base: base + 4 * nc].view(np.uint32)
        packed = self.blob[base + 4 * nc: base + 4 * nc + (L + 1) // 2]
        codes = np.empty(2 * len(packed), np.uint8)
        codes[0::2] = packed >> 4
        codes[1::2] = packed & 0xF
        qual = self.blob[base + 4 * nc + (L + 1) // 2: base + 4 * nc + (L + 1) // 2 + L]
        return {
            "pos": int(self.pos[i]), "flag": int(self.flag[i]), "mapq": int(self.mapq[i]),
            "tlen": int(self.tlen[i]), "bc_idx": int(self.bc_idx[i]),
            "seq": "".join(SEQ_NT16[c] for c in codes[:L]),
            "qual": qual.tolist(),
            "cigar": [(int(w & 0xF), int(w >> 4)) for w in words],
        }

    # ------------------------------------------------------------------ (de)serialisation for fixtures
    def to_npz_dict(self, prefix: str = "in_") -> dict:
        return {prefix + f: getattr(self, f) for f, _ in FIELDS}

    @staticmethod
    def from_npz_dict(d, prefix: str = "in_") -> "ReadBatch":
        return ReadBatch(**{f: np.asarray(d[prefix + f]) for f, _ in FIELDS})
