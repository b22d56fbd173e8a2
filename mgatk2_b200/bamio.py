"""BAM ingest for the drop-in seam: replaces pysam at readers.py:37-61,85-165.

`read_bam_chrM()` decodes every record placed on the mitochondrial contig into the structure-of-arrays `ReadBatch`
through the native reader `csrc/bamio.cpp` (BGZF inflate on host threads, records parsed straight into the batch, the
cigar|seq|qual region of each record copied verbatim). Barcode strings are mapped to whitelist indices here, on the
distinct values only. `write_bam()` is the inverse (pure Python, zlib): it turns a batch into a BGZF BAM plus a minimal
.bai, which is how the synthetic generator feeds the reader in the tests and how a batch can be handed to other tools.
"""
from __future__ import annotations

import ctypes
import weakref
import os
import struct
import subprocess
import zlib

import numpy as np

from .batch import ReadBatch
from .exceptions import BAMFormatError, BAMReadError, NoBarcodeTagsError, NoChrMReadsError

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgatk2_bamio.so")
SRC_PATH = os.path.join(HERE, "csrc", "bamio.cpp")
MITO_NAMES = ("chrM", "MT", "M", "chrMT")              # readers.py:43 order
EXPORTS = ("mgatk_bam_open", "mgatk_bam_close", "mgatk_bam_error", "mgatk_bam_n_refs", "mgatk_bam_ref_name",
           "mgatk_bam_ref_len", "mgatk_bam_coordinate_sorted", "mgatk_bam_fetch", "mgatk_bam_n_records",
           "mgatk_bam_blob_bytes", "mgatk_bam_n_barcodes", "mgatk_bam_barcode_bytes", "mgatk_bam_export", "mgatk_bam_detach",
           "mgatk_bam_free", "mgatk_bam_fetch_more", "mgatk_bam_align_parts", "mgatk_bai_inspect")
_lib = None


def build_bamio(force: bool = False) -> str:
    """g++ -O2 -shared -fPIC csrc/bamio.cpp -lz -pthread -> libmgatk2_bamio.so (in-tree, host only)."""
    deps = (SRC_PATH, os.path.join(HERE, "csrc", "fast_inflate.h"), os.path.join(HERE, "..", "include", "mgatk2_bamio.h"))
    stale = not os.path.exists(LIB_PATH) or any(os.path.getmtime(d) > os.path.getmtime(LIB_PATH) for d in deps)
    if force or stale:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB_PATH, SRC_PATH, "-lz", "-pthread"], check=True)
    return LIB_PATH


def inspect_bai(bai_path: str, ref_id: int) -> dict | None:
    """Entry of reference `ref_id` in a BAM index, through the parser the native fetch uses for its start offset:
    counts of references / bins / chunks / linear-index intervals, the smallest chunk virtual offset and — when the
    index carries htslib's metadata pseudo-bin — the file range of the reference and its mapped / unmapped-placed
    record counts (what `fetch(contig)` will return, readers.py:85-93). None when the file is not a BAI covering it."""
    out = (ctypes.c_int64 * 10)()
    if load().mgatk_bai_inspect(str(bai_path).encode(), int(ref_id), out):
        return None
    keys = ("n_ref", "n_bin", "n_chunk", "n_intv", "min_voff", "has_meta", "ref_beg", "ref_end", "n_mapped", "n_unmapped")
    return dict(zip(keys, (int(v) for v in out)))


def load():
    global _lib
    if _lib is None:
        build_bamio()
        lib = ctypes.CDLL(LIB_PATH)
        lib.mgatk_bam_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.mgatk_bam_close.argtypes = [ctypes.c_void_p]
        lib.mgatk_bam_error.argtypes = [ctypes.c_void_p]
        lib.mgatk_bam_error.restype = ctypes.c_char_p
        lib.mgatk_bam_n_refs.argtypes = [ctypes.c_void_p]
        lib.mgatk_bam_ref_name.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.mgatk_bam_ref_name.restype = ctypes.c_char_p
        lib.mgatk_bam_ref_len.argtypes = [ctypes.c_void_p, ctypes.c_int]
        lib.mgatk_bam_ref_len.restype = ctypes.c_int64
        lib.mgatk_bam_coordinate_sorted.argtypes = [ctypes.c_void_p]
        lib.mgatk_bam_fetch.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_int64]
        lib.mgatk_bam_fetch_more.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int64]
        lib.mgatk_bam_align_parts.argtypes = [ctypes.c_void_p, ctypes.c_int]
        for f in ("mgatk_bam_n_records", "mgatk_bam_blob_bytes", "mgatk_bam_n_barcodes", "mgatk_bam_barcode_bytes"):
            getattr(lib, f).argtypes = [ctypes.c_void_p]
            getattr(lib, f).restype = ctypes.c_int64
        lib.mgatk_bam_export.argtypes = [ctypes.c_void_p] + [ctypes.c_void_p] * 12
        lib.mgatk_bam_detach.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p * 10), ctypes.c_void_p, ctypes.c_void_p]
        lib.mgatk_bam_free.argtypes = [ctypes.c_void_p]
        lib.mgatk_bam_free.restype = None
        lib.mgatk_bai_inspect.argtypes = [ctypes.c_char_p, ctypes.c_int, ctypes.POINTER(ctypes.c_int64)]
        _lib = lib
    return _lib


class BamFile:
    """Thin handle on the native reader: references and `fetch(contig)` into arrays."""

    def __init__(self, path: str):
        self.lib = load()
        self.path = str(path)
        h = ctypes.c_void_p()
        rc = self.lib.mgatk_bam_open(self.path.encode(), ctypes.byref(h))
        self.h = h
        if rc:
            msg = self.lib.mgatk_bam_error(h).decode() if h else "cannot open"
            self.close()
            raise BAMFormatError(self.path, f"Cannot open: {msg}")       # readers.py:38-39
        n = self.lib.mgatk_bam_n_refs(self.h)
        self.references = [self.lib.mgatk_bam_ref_name(self.h, i).decode() for i in range(n)]
        self.lengths = [int(self.lib.mgatk_bam_ref_len(self.h, i)) for i in range(n)]
        self.coordinate_sorted = bool(self.lib.mgatk_bam_coordinate_sorted(self.h))

    def close(self):
        if getattr(self, "h", None):
            self.lib.mgatk_bam_close(self.h)
            self.h = None

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def fetch(self, contig: str, tag: str = "CB", threads: int | None = None, max_records: int = -1):
        """Every record placed on `contig`, in file order: (ReadBatch with bc_idx = index into `barcodes` or -1/-2,
        list of distinct barcode strings in order of first appearance, qual_missing mask)."""
        if contig not in self.references:
            raise BAMReadError(self.path, f"contig {contig!r} not in the header")
        if len(tag) != 2:
            raise BAMReadError(self.path, f"barcode tag {tag!r} is not two characters")
        threads = threads or min(16, os.cpu_count() or 1)
        rc = self.lib.mgatk_bam_fetch(self.h, self.references.index(contig), tag.encode(), int(threads), int(max_records))
        return self._collect(rc)

    def fetch_more(self, threads: int | None = None, max_records: int = -1):
        """The next part of a fetch that stopped at max_records (same return value; no records when the contig is done).
        Barcode indices of all parts refer to one growing table, returned in full every time."""
        threads = threads or min(16, os.cpu_count() or 1)
        return self._collect(self.lib.mgatk_bam_fetch_more(self.h, int(threads), int(max_records)))

    def _collect(self, rc: int):
        if rc:
            raise BAMReadError(self.path, f"Read error: {self.lib.mgatk_bam_error(self.h).decode()}")
        n = int(self.lib.mgatk_bam_n_records(self.h))
        nb, nbc = int(self.lib.mgatk_bam_blob_bytes(self.h)), int(self.lib.mgatk_bam_n_barcodes(self.h))
        chars = np.empty(max(int(self.lib.mgatk_bam_barcode_bytes(self.h)), 1), np.uint8)
        ends = np.empty(max(nbc, 1), np.int64)
        p = lambda x: x.ctypes.data_as(ctypes.c_void_p)
        # the arrays change hands without a copy (mgatk_bam_detach): numpy views over the library's malloc blocks, freed
        # when the last view of a block is gone
        ptrs = (ctypes.c_void_p * 10)()
        self.lib.mgatk_bam_detach(self.h, ctypes.byref(ptrs), p(chars), p(ends))
        layout = (("pos", np.int32, n), ("tlen", np.int32, n), ("flag", np.uint16, n), ("mapq", np.uint8, n),
                  ("bc_idx", np.int32, n), ("l_seq", np.uint16, n), ("n_cigar", np.uint16, n), ("blob_off", np.uint32, n),
                  ("blob", np.uint8, nb), ("qm", np.uint8, n))
        a = {}
        for (name, dt, count), ptr in zip(layout, ptrs):
            a[name] = _adopt(self.lib, ptr, np.dtype(dt), count)
        qm = a.pop("qm")
        raw = chars.tobytes()
        barcodes, o = [], 0
        for i in range(nbc):
            barcodes.append(raw[o:int(ends[i])].decode("ascii", "replace"))
            o = int(ends[i])
        return ReadBatch(**a), barcodes, qm.astype(bool)


def _adopt(lib, ptr, dtype, count):
    """numpy array over a malloc block handed over by the library; the block is freed with the last view of it."""
    if not ptr or count == 0:
        if ptr:
            lib.mgatk_bam_free(ptr)
        return np.empty(0, dtype)
    buf = (ctypes.c_uint8 * (count * dtype.itemsize)).from_address(ptr)
    weakref.finalize(buf, lib.mgatk_bam_free, ptr)
    return np.frombuffer(buf, dtype=dtype)


def pick_mito_contig(references) -> str | None:
    """readers.py:42-51: the first of chrM, MT, M, chrMT present in the header."""
    for name in MITO_NAMES:
        if name in references:
            return name
    return None


def _finish_part(path, config, wl_index, batch, barcodes, qual_missing, first_part: bool):
    """Whitelist mapping and the per-record refusals of read_bam_chrM for one part of a fetch."""
    has_tag = batch.bc_idx != -1
    if first_part and batch.n_records > 1000 and not has_tag[:1001].any():     # readers.py:54-59
        raise NoBarcodeTagsError(str(path), config.barcode_tag, 1000)
    if set(wl_index) == {"bulk"}:
        bc = np.zeros(batch.n_records, np.int32)
    else:
        table = np.array([wl_index.get(b, -1) for b in barcodes] + [-1, -1], dtype=np.int32)   # ids -2 / -1 -> -1
        bc = table[batch.bc_idx] if batch.n_records else np.zeros(0, np.int32)
    batch.bc_idx = np.ascontiguousarray(bc, dtype=np.int32)
    ok = ((batch.flag & 0x904) == 0) & (batch.bc_idx >= 0)
    for i in np.nonzero(qual_missing & ok)[0].tolist():
        # readers.py:146-158: np.array(None) raises when the SimpleRead is built, i.e. only for a record that survived
        # dedup. The few QUAL-less survivors of stage 1 are checked here the way the reference's sets would see them
        # (an earlier stage-1 record of the same barcode, start and strand [and |tlen|] makes this one a duplicate).
        dup = False
        if not config.dedup.skip:
            j = i - 1
            while j >= 0 and batch.pos[j] == batch.pos[i] and not dup:
                dup = bool(ok[j] and batch.bc_idx[j] == batch.bc_idx[i] and (batch.flag[j] & 0x10) == (batch.flag[i] & 0x10)
                           and (not config.dedup.use_fragment_length
                                or abs(int(batch.tlen[j])) == abs(int(batch.tlen[i]))))
                j -= 1
        if not dup:
            raise BAMReadError(str(path), "Read error: record without base qualities")
    return batch


def iter_bam_chrM(path: str, config, wl_index: dict, max_records: int, threads: int | None = None):
    """read_bam_chrM in parts of about `max_records` records for inputs that do not fit host memory (BASELINE
    configs[4]): yields ReadBatches in file order, cut between different reference_start values - all candidates for a
    duplicate of a read share its start (readers.py:118-150) - which is what `PileupEngine.run_stream` takes. A part
    holds at least `max_records` records (it runs on to the next start border) unless it is the last."""
    if max_records < 1:
        raise ValueError("max_records must be positive")
    with BamFile(path) as bam:
        mito = pick_mito_contig(bam.references)
        if mito is None:
            raise NoChrMReadsError(str(path), bam.references)
        bam.lib.mgatk_bam_align_parts(bam.h, 1)        # the reader itself runs every part on to the next start border
        part = bam.fetch(mito, config.barcode_tag, threads, max_records)
        first, last_pos = True, None
        while True:
            batch, barcodes, qm = part
            if batch.n_records == 0:
                return
            batch = _finish_part(path, config, wl_index, batch, barcodes, qm, first)
            first = False
            if not batch.is_sorted() or (last_pos is not None and int(batch.pos[0]) <= last_pos):
                raise BAMReadError(str(path), "Read error: records are not sorted by reference_start")
            last_pos = int(batch.pos[-1])
            got = batch.n_records
            yield batch
            if got < max_records:                      # the contig is exhausted
                return
            part = bam.fetch_more(threads, max_records)


def read_bam_chrM(path: str, config, wl_index: dict, threads: int | None = None):
    """BAMReader._validate_bam_file + the fetch loop's record decode (readers.py:35-61,85-111,153-165).

    Returns (ReadBatch with bc_idx = whitelist index or -1, mito contig name). `wl_index` maps barcode -> column;
    the reference's bulk mode (barcodes == {"bulk"}, readers.py:72,100-102) puts every record in column 0.
    """
    with BamFile(path) as bam:
        mito = pick_mito_contig(bam.references)
        if mito is None:
            raise NoChrMReadsError(str(path), bam.references)
        batch, barcodes, qual_missing = bam.fetch(mito, config.barcode_tag, threads)
    batch = _finish_part(path, config, wl_index, batch, barcodes, qual_missing, True)
    return batch, mito


# ------------------------------------------------------------------------------------------- writer (tests, tools)
def _reg2bin(beg: int, end: int) -> int:
    end -= 1
    if beg >> 14 == end >> 14:
        return ((1 << 15) - 1) // 7 + (beg >> 14)
    if beg >> 17 == end >> 17:
        return ((1 << 12) - 1) // 7 + (beg >> 17)
    if beg >> 20 == end >> 20:
        return ((1 << 9) - 1) // 7 + (beg >> 20)
    if beg >> 23 == end >> 23:
        return ((1 << 6) - 1) // 7 + (beg >> 23)
    if beg >> 26 == end >> 26:
        return ((1 << 3) - 1) // 7 + (beg >> 26)
    return 0


def _bgzf_block(data: bytes, level: int = 1) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, -15)
    cdata = c.compress(data) + c.flush()
    bsize = len(cdata) + 25
    return (struct.pack("<BBBBIBBHBBHH", 31, 139, 8, 4, 0, 0, 255, 6, 66, 67, 2, bsize) + cdata +
            struct.pack("<II", zlib.crc32(data) & 0xFFFFFFFF, len(data)))


def write_bam(path: str, batch: ReadBatch, barcodes, *, ref_names=("chr1", "chrM", "chrX"), ref_lens=(248956422, 16569, 156040895),
              mito="chrM", tag="CB", cb_strings=None, extra=None, sorted_header=True, write_index=True, block_bytes=60000,
              other_tags=True):
    """Write `batch` as the records of contig `mito` of a BGZF BAM (+ `path.bai` with one bin per contig).

    cb_strings: per-record barcode string or None (no tag); default: whitelist entry for bc_idx >= 0, a non-whitelisted
    string for odd records with bc_idx < 0 and no tag for even ones. extra: records of other contigs as
    (ref_name | None, pos) pairs, written before / after the mito records according to the header order."""
    refs = list(ref_names)
    mi = refs.index(mito)
    text = ("@HD\tVN:1.6\tSO:%s\n" % ("coordinate" if sorted_header else "unsorted")) + \
           "".join(f"@SQ\tSN:{n}\tLN:{l}\n" for n, l in zip(refs, ref_lens))
    hdr = b"BAM\1" + struct.pack("<I", len(text)) + text.encode() + struct.pack("<I", len(refs))
    for n, l in zip(refs, ref_lens):
        hdr += struct.pack("<I", len(n) + 1) + n.encode() + b"\0" + struct.pack("<I", l)

    def record(ref, pos, flag, mapq, tlen, cig_words, seq_packed, qual, l_seq, name, cb):
        span = sum((w >> 4) for w in cig_words if (w & 15) in (0, 2, 3, 7, 8)) or 1
        aux = b""
        if other_tags:                  # tags of several types before the barcode: the scanner has to skip them
            aux += b"NMC\x03" + b"ASs" + struct.pack("<h", -7) + b"XBBS" + struct.pack("<IHH", 2, 5, 6) + b"RGZgrp1\0"
        if cb is not None:
            aux += tag.encode() + b"Z" + cb.encode() + b"\0"
        body = struct.pack("<iiBBHHHIiii", ref, pos, len(name) + 1, mapq, _reg2bin(max(pos, 0), max(pos, 0) + span),
                           len(cig_words), flag, l_seq, -1, -1, tlen)
        body += name.encode() + b"\0" + b"".join(struct.pack("<I", w) for w in cig_words) + seq_packed + qual + aux
        return struct.pack("<I", len(body)) + body

    recs = []                            # (ref index or big number for unplaced, payload)
    for ref_name, pos in (extra or []):
        ri = refs.index(ref_name) if ref_name is not None else -1
        recs.append((ri if ri >= 0 else 1 << 30, pos, record(ri, pos, 0 if ri >= 0 else 4, 60, 0, [(20 << 4)], b"\x11" * 10,
                                                             b"\x1e" * 20, 20, "x", "ZZZZ-1")))
    for i in range(batch.n_records):
        L, nc = int(batch.l_seq[i]), int(batch.n_cigar[i])
        base = int(batch.blob_off[i]) * 16
        words = batch.blob[base: base + 4 * nc].view(np.uint32).tolist()
        seqp = batch.blob[base + 4 * nc: base + 4 * nc + (L + 1) // 2].tobytes()
        qual = batch.blob[base + 4 * nc + (L + 1) // 2: base + 4 * nc + (L + 1) // 2 + L].tobytes()
        if cb_strings is not None:
            cb = cb_strings[i]
        else:
            b = int(batch.bc_idx[i])
            cb = barcodes[b] if b >= 0 else ("NOTINLIST-1" if i % 2 else None)
        recs.append((mi, i, record(mi, int(batch.pos[i]), int(batch.flag[i]), int(batch.mapq[i]), int(batch.tlen[i]),
                                   words, seqp, qual, L, f"r{i}", cb)))
    if sorted_header:
        recs.sort(key=lambda r: r[0])    # stable: contig order of the header, file order inside a contig
    out = bytearray()
    ustream = bytearray(hdr)
    first_voff, last_voff = {}, {}
    coff = 0

    def flush():
        nonlocal coff, ustream
        if ustream:
            blk = _bgzf_block(bytes(ustream))
            out.extend(blk)
            coff += len(blk)
            ustream = bytearray()

    for ri, _, payload in recs:
        if len(ustream) + len(payload) > block_bytes and len(ustream) > 0 and len(payload) <= block_bytes:
            flush()
        first_voff.setdefault(ri, (coff << 16) | len(ustream))
        # long records may span blocks: cut the stream at block_bytes
        ustream.extend(payload)
        while len(ustream) > 65000:
            rest = ustream[block_bytes:]
            ustream = ustream[:block_bytes]
            flush()
            ustream = bytearray(rest)
        last_voff[ri] = (coff << 16) | len(ustream)
    flush()
    out.extend(_bgzf_block(b""))                                    # EOF marker
    with open(path, "wb") as f:
        f.write(bytes(out))
    if write_index:
        bai = bytearray(b"BAI\1" + struct.pack("<i", len(refs)))
        for r in range(len(refs)):
            if r in first_voff:
                bai += struct.pack("<i", 2)
                bai += struct.pack("<Ii", 0, 1) + struct.pack("<QQ", first_voff[r], last_voff[r])       # everything in bin 0
                bai += struct.pack("<Ii", 37450, 2) + struct.pack("<QQQQ", first_voff[r], last_voff[r], 0, 0)
                bai += struct.pack("<i", 1) + struct.pack("<Q", first_voff[r])
            else:
                bai += struct.pack("<i", 0) + struct.pack("<i", 0)
        bai += struct.pack("<Q", 0)
        with open(str(path) + ".bai", "wb") as f:
            f.write(bytes(bai))
    return path
