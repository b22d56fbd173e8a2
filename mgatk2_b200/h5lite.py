"""A small HDF5 writer (and the reader its tests need) for the two files `mgatk2 run` writes by default:
`output/counts.h5` and `output/metadata.h5` (reference src/file_io/writers.py:60-134,325-395; read back by
R/mgatk2_functions.R:18-61). h5py / libhdf5 are not in this image, so the format is written directly, in the same
variant the reference's files use (h5py `libver="latest"`): version-3 superblock, version-2 object headers with
Jenkins lookup3 checksums, links stored compactly in the group's header, chunked datasets (gzip, level 4,
chunks (1000, 100)) indexed by a Fixed Array of filtered-chunk entries, contiguous small datasets, scalar attributes.

The reader understands exactly what such files contain - including what libhdf5 itself writes that this writer avoids
(dense link storage in a fractal heap, variable-length string attributes in the global heap, NIL padding) - and verifies
every checksum it meets. `tests/test_h5lite.py` reads the reference's own committed files
(`tests/run_hdf5_output/output/*.h5`, copied to tests/golden/) with it, then reads what the writer wrote through the same
code: structure and checksums of the written files are held to what libhdf5 produced.
"""
from __future__ import annotations

import struct
import zlib
from concurrent.futures import ThreadPoolExecutor

import numpy as np

UNDEF = 0xFFFFFFFFFFFFFFFF
SIGNATURE = b"\x89HDF\r\n\x1a\n"


# --------------------------------------------------------------------------------------------- checksum
def _rot(x, k):
    return ((x << k) | (x >> (32 - k))) & 0xFFFFFFFF


def lookup3(data: bytes, initval: int = 0) -> int:
    """Bob Jenkins' lookup3 `hashlittle` (the metadata checksum of HDF5 1.8+, H5_checksum_lookup3)."""
    length = len(data)
    a = b = c = (0xDEADBEEF + length + initval) & 0xFFFFFFFF
    off = 0
    M = 0xFFFFFFFF
    while length > 12:
        a = (a + int.from_bytes(data[off:off + 4], "little")) & M
        b = (b + int.from_bytes(data[off + 4:off + 8], "little")) & M
        c = (c + int.from_bytes(data[off + 8:off + 12], "little")) & M
        a = (a - c) & M; a ^= _rot(c, 4); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 6); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 8); b = (b + a) & M
        a = (a - c) & M; a ^= _rot(c, 16); c = (c + b) & M
        b = (b - a) & M; b ^= _rot(a, 19); a = (a + c) & M
        c = (c - b) & M; c ^= _rot(b, 4); b = (b + a) & M
        off += 12
        length -= 12
    if length == 0:
        return c
    tail = data[off:off + length] + b"\0" * (12 - length)
    a = (a + int.from_bytes(tail[0:4], "little")) & M
    b = (b + int.from_bytes(tail[4:8], "little")) & M
    c = (c + int.from_bytes(tail[8:12], "little")) & M
    c ^= b; c = (c - _rot(b, 14)) & M
    a ^= c; a = (a - _rot(c, 11)) & M
    b ^= a; b = (b - _rot(a, 25)) & M
    c ^= b; c = (c - _rot(b, 16)) & M
    a ^= c; a = (a - _rot(c, 4)) & M
    b ^= a; b = (b - _rot(a, 14)) & M
    c ^= b; c = (c - _rot(b, 24)) & M
    return c


# --------------------------------------------------------------------------------------------- datatypes
def encode_datatype(dt: np.dtype) -> bytes:
    """Datatype message body (version 1) for the types the files hold: unsigned / signed integers, IEEE floats,
    fixed-length byte strings (numpy 'S': null-padded ASCII, as h5py stores them)."""
    dt = np.dtype(dt)
    if dt.kind in "ui":
        bits = 0x08 if dt.kind == "i" else 0x00            # little-endian, sign bit 3
        return struct.pack("<BBBBIHH", 0x10 | 0, bits, 0, 0, dt.itemsize, 0, dt.itemsize * 8)
    if dt.kind == "f":
        if dt.itemsize == 4:
            return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, 0x20, 31, 0, 4, 0, 32, 23, 8, 0, 23, 127)
        return struct.pack("<BBBBIHHBBBBI", 0x10 | 1, 0x20, 63, 0, 8, 0, 64, 52, 11, 0, 52, 1023)
    if dt.kind == "S":
        return struct.pack("<BBBBI", 0x10 | 3, 0x01, 0, 0, max(dt.itemsize, 1))
    raise TypeError(f"unsupported dtype {dt}")


def decode_datatype(b: bytes):
    """-> (numpy dtype or ('vlen_str',), bytes consumed)."""
    cls, ver = b[0] & 15, b[0] >> 4
    bits = b[1] | (b[2] << 8) | (b[3] << 16)
    size = struct.unpack_from("<I", b, 4)[0]
    if cls == 0:
        if bits & 1:
            raise ValueError("big-endian integers not supported")
        return np.dtype(("<i" if bits & 8 else "<u") + str(size)), 12
    if cls == 1:
        return np.dtype("<f" + str(size)), 20
    if cls == 3:
        return np.dtype("S" + str(size)), 8
    if cls == 9:                                             # variable length: base type follows
        base, used = decode_datatype(b[8:])
        if (bits & 15) == 1:
            return ("vlen_str",), 8 + used
        return ("vlen", base), 8 + used
    raise ValueError(f"datatype class {cls} (version {ver}) not supported")


# --------------------------------------------------------------------------------------------- writer
class _Out:
    """Append-only output with a running offset; written straight to the file (chunks never pile up in memory)."""

    def __init__(self, path):
        self.f = open(path, "wb")
        self.pos = 0

    def tell(self):
        return self.pos

    def align(self, a=8):
        self.put(b"\0" * (-self.pos % a))

    def put(self, b: bytes) -> int:
        at = self.pos
        if b:
            self.f.write(b)
            self.pos += len(b)
        return at

    def patch(self, at: int, b: bytes):
        self.f.seek(at)
        self.f.write(b)
        self.f.seek(self.pos)

    def close(self):
        self.f.close()


def _msg(mtype: int, body: bytes, flags: int = 0) -> bytes:
    return struct.pack("<BHB", mtype, len(body), flags) + body


def _dataspace(shape) -> bytes:
    if len(shape) == 0:
        return struct.pack("<BBBB", 2, 0, 0, 0)              # version 2, scalar
    return struct.pack("<BBBB", 2, len(shape), 0, 1) + b"".join(struct.pack("<Q", int(s)) for s in shape)


VLEN_STR_TYPE = struct.pack("<BBBBI", 0x10 | 9, 0x01, 0x01, 0, 16) + struct.pack("<BBBBIHH", 0x10, 0, 0, 0, 1, 0, 8)   # UTF-8, base: 1-byte integer


def _attribute(name: str, value, heap=None) -> bytes:
    """Attribute message (version 3). A Python str becomes a variable-length UTF-8 string in the global heap, which is
    how h5py stores `attrs[name] = "text"` (and reads it back as str); integers become int64 scalars."""
    nm = name.encode("utf-8") + b"\0"
    if isinstance(value, str):
        raw = value.encode("utf-8")
        dt, ds = VLEN_STR_TYPE, _dataspace(())
        data = struct.pack("<IQI", len(raw), heap.address, heap.add(raw))
    else:
        arr = np.asarray(value)
        if arr.dtype.kind in "iu":
            arr = arr.astype(np.int64)
        dt, ds, data = encode_datatype(arr.dtype), _dataspace(arr.shape), arr.tobytes()
    body = struct.pack("<BBHHHB", 3, 0, len(nm), len(dt), len(ds), 0) + nm + dt + ds + data
    return _msg(0x0C, body)


class _GlobalHeap:
    """One global heap collection (4096 bytes at a fixed address) for the variable-length strings of attributes."""

    SIZE = 4096

    def __init__(self, address: int):
        self.address, self.objects = address, []

    def add(self, raw: bytes) -> int:
        self.objects.append(raw)
        return len(self.objects)

    def encode(self) -> bytes:
        body = b""
        for k, raw in enumerate(self.objects, start=1):
            body += struct.pack("<HHIQ", k, 0, 0, len(raw)) + raw + b"\0" * (-len(raw) % 8)
        free = self.SIZE - 16 - len(body)
        if free < 16:
            raise ValueError("too many string attributes for one heap collection")
        body += struct.pack("<HHIQ", 0, 0, 0, free)          # object 0: the free space
        blob = b"GCOL" + struct.pack("<BBBBQ", 1, 0, 0, 0, self.SIZE) + body
        return blob + b"\0" * (self.SIZE - len(blob))


def _object_header(messages: list) -> bytes:
    body = b"".join(messages)
    head = b"OHDR" + struct.pack("<BB", 2, 0x02) + struct.pack("<I", len(body))    # flags: 4-byte size of chunk 0
    blob = head + body
    return blob + struct.pack("<I", lookup3(blob))


def _chunk_size_bytes(chunk_nbytes: int) -> int:
    """Bytes of the 'chunk size' field of a filtered-chunk index entry (H5D layout: 1 + (log2(size) + 8) / 8, at most 8)."""
    return min(8, 1 + ((int(chunk_nbytes).bit_length() - 1) + 8) // 8)


class H5Writer:
    """with H5Writer(path) as f: f.attr(name, value); f.dataset(name, array, chunks=..., gzip=4); f.group(name) ..."""

    def __init__(self, path, threads: int = 8):
        self.path = path
        self.out = _Out(path)
        self.out.put(b"\0" * 64)                             # superblock (48 bytes), written last
        self.heap = _GlobalHeap(self.out.put(b"\0" * _GlobalHeap.SIZE))
        self.root = {"links": [], "attrs": []}
        self.threads = threads

    def __enter__(self):
        return self

    def __exit__(self, et, ev, tb):
        if et is None:
            self.close()
        else:
            self.out.close()
        return False

    def attr(self, name: str, value, group=None):
        (group or self.root)["attrs"].append(_attribute(name, value, self.heap))

    def group(self, name: str, parent=None) -> dict:
        g = {"links": [], "attrs": [], "name": name}
        (parent or self.root)["links"].append((name, g))
        return g

    def dataset(self, name: str, data: np.ndarray, chunks=None, gzip: int | None = None, parent=None, chunk_source=None,
                shape=None, dtype=None):
        """A contiguous dataset (`chunks` None) or a chunked one. `chunk_source(i0, i1, j0, j1) -> array` (2-D datasets
        only) supplies chunk contents on demand instead of `data` (the planes are transposed chunk by chunk, never as a
        whole)."""
        if data is not None:
            data = np.ascontiguousarray(data)
            shape, dtype = data.shape, data.dtype
        if chunks is not None and (int(np.prod(shape)) == 0):
            chunks = None                                    # nothing to chunk: an empty contiguous dataset
            data = np.zeros(shape, dtype)
        dtype = np.dtype(dtype)
        msgs = [_msg(0x01, _dataspace(shape)), _msg(0x03, encode_datatype(dtype), 1)]
        o = self.out
        if chunks is None:
            msgs.append(_msg(0x05, struct.pack("<BB", 3, 0x0A), 1))       # fill value v3: late allocation, write if set, default
            raw = data.tobytes()
            o.align()
            addr = o.put(raw) if raw else UNDEF
            msgs.append(_msg(0x08, struct.pack("<BBQQ", 4, 1, addr, len(raw))))
        else:
            chunks = tuple(int(min(c, s)) if s else int(c) for c, s in zip(chunks, shape))
            msgs.append(_msg(0x05, struct.pack("<BB", 3, 0x0B), 1))       # incremental allocation, write if set
            level = 4 if gzip is None else int(gzip)
            msgs.append(_msg(0x0B, struct.pack("<BBHHHI", 2, 1, 1, 1, 1, level), 1))   # filter pipeline v2: deflate, optional, one value
            grid = [(-(-s // c)) if s else 0 for s, c in zip(shape, chunks)]
            n_chunks = int(np.prod(grid)) if len(grid) else 0
            chunk_nbytes = int(np.prod(chunks)) * dtype.itemsize

            def build(k):
                idx = np.unravel_index(k, grid)
                lo = [i * c for i, c in zip(idx, chunks)]
                hi = [min(l + c, s) for l, c, s in zip(lo, chunks, shape)]
                if chunk_source is not None:
                    part = chunk_source(lo[0], hi[0], lo[1], hi[1])
                    if part is None:
                        return None                          # never written: no storage, reads as the fill value
                    part = np.asarray(part, dtype=dtype)
                else:
                    part = data[tuple(slice(l, h) for l, h in zip(lo, hi))]
                if part.shape != chunks:                     # edge chunks are stored whole, padded with the fill value
                    full = np.zeros(chunks, dtype)
                    full[tuple(slice(0, n) for n in part.shape)] = part
                    part = full
                return zlib.compress(np.ascontiguousarray(part).tobytes(), level)

            with ThreadPoolExecutor(max_workers=self.threads) as pool:     # zlib releases the GIL
                packed = list(pool.map(build, range(n_chunks)))
            entries = [(o.put(z), len(z)) if z is not None else (UNDEF, 0) for z in packed]
            enc = max(1, max((int(c).bit_length() + 7) // 8 for c in list(chunks) + [dtype.itemsize]))
            dims = b"".join(int(c).to_bytes(enc, "little") for c in list(chunks) + [dtype.itemsize])
            if n_chunks == 1 and entries[0][0] != UNDEF:     # "single chunk" index: the chunk's address sits in the layout message
                body = struct.pack("<BBBBB", 4, 2, 0x02, len(chunks) + 1, enc) + dims + struct.pack("<BQI", 1, entries[0][1], 0) + \
                    struct.pack("<Q", entries[0][0])
            else:
                page_bits = 10
                szb = _chunk_size_bytes(chunk_nbytes)
                esize = 8 + szb + 4
                o.align()
                fahd_addr = o.tell()
                dblk_addr = fahd_addr + 28
                head = b"FAHD" + struct.pack("<BBBBQQ", 0, 1, esize, page_bits, n_chunks, dblk_addr)
                o.put(head + struct.pack("<I", lookup3(head)))
                elem = [struct.pack("<Q", a) + int(n).to_bytes(szb, "little") + struct.pack("<I", 0) for a, n in entries]
                page = 1 << page_bits
                pre = b"FADB" + struct.pack("<BBQ", 0, 1, fahd_addr)
                if n_chunks <= page:
                    blk = pre + b"".join(elem)
                    o.put(blk + struct.pack("<I", lookup3(blk)))
                else:                                        # paged data block: bitmap of initialised pages, checksum, then pages
                    n_pages = -(-n_chunks // page)
                    bitmap = bytearray((n_pages + 7) // 8)
                    for pg in range(n_pages):
                        bitmap[pg // 8] |= 0x80 >> (pg % 8)
                    blk = pre + bytes(bitmap)
                    o.put(blk + struct.pack("<I", lookup3(blk)))
                    for pg in range(n_pages):
                        part = elem[pg * page:(pg + 1) * page]
                        part += [struct.pack("<Q", UNDEF) + b"\0" * (szb + 4)] * (page - len(part) if pg < n_pages - 1 else 0)
                        pb = b"".join(part)
                        o.put(pb + struct.pack("<I", lookup3(pb)))
                body = struct.pack("<BBBBB", 4, 2, 0x00, len(chunks) + 1, enc) + dims + struct.pack("<BB", 3, page_bits) + \
                    struct.pack("<Q", fahd_addr)
            msgs.append(_msg(0x08, body))
        o.align()
        addr = o.put(_object_header(msgs))
        (parent or self.root)["links"].append((name, addr))
        return addr

    def _write_group(self, g: dict) -> int:
        links = []
        for name, target in g["links"]:
            addr = self._write_group(target) if isinstance(target, dict) else target
            nm = name.encode("utf-8")
            links.append(_msg(0x06, struct.pack("<BBB", 1, 0x00, len(nm)) + nm + struct.pack("<Q", addr)))
        msgs = [_msg(0x02, struct.pack("<BBQQ", 0, 0, UNDEF, UNDEF)),                        # link info: compact storage
                _msg(0x0A, struct.pack("<BBHH", 0, 0x01, 0xFFFF, 0xFFFF), 1)]                  # group info: links stay compact
        msgs += links + g["attrs"]
        self.out.align()
        return self.out.put(_object_header(msgs))

    def close(self):
        root_addr = self._write_group(self.root)
        eof = self.out.tell()
        sb = SIGNATURE + struct.pack("<BBBBQQQQ", 3, 8, 8, 0, 0, UNDEF, eof, root_addr)
        self.out.patch(0, sb + struct.pack("<I", lookup3(sb)))
        self.out.patch(self.heap.address, self.heap.encode())
        self.out.close()


# --------------------------------------------------------------------------------------------- reader
class H5Dataset:
    def __init__(self, f, name, shape, dtype, layout, filters, attrs):
        self.file, self.name, self.shape, self.dtype, self.layout, self.filters, self.attrs = f, name, shape, dtype, layout, filters, attrs

    def read(self) -> np.ndarray:
        return self.file._read_dataset(self)


class H5Reader:
    """Read-only view of a file in the variant described in the module docstring. `objects` maps paths ("A_fwd",
    "barcode_metadata/is_cell") to H5Dataset; `attrs` are the root group's attributes. Every metadata checksum on the
    way is verified (ValueError otherwise)."""

    def __init__(self, path):
        self.d = open(path, "rb").read()
        d = self.d
        if d[:8] != SIGNATURE:
            raise ValueError("not an HDF5 file")
        self.version = d[8]
        if self.version not in (2, 3) or d[9] != 8 or d[10] != 8:
            raise ValueError("superblock version / offset size not supported")
        base, ext, self.eof, root = struct.unpack_from("<QQQQ", d, 12)
        self._check(0, 44)
        self.structures = {"superblock": self.version}
        self.objects: dict[str, H5Dataset] = {}
        self.groups: dict[str, dict] = {}
        self.attrs = self._walk_group(root, "")

    def _check(self, start, length):
        if start + length + 4 > len(self.d):
            raise ValueError(f"metadata block at {start:#x} runs past the end of the file")
        want = struct.unpack_from("<I", self.d, start + length)[0]
        if lookup3(self.d[start:start + length]) != want:
            raise ValueError(f"checksum mismatch at {start:#x}")

    def _count(self, what):
        self.structures[what] = self.structures.get(what, 0) + 1

    # ---- object headers
    def _messages(self, addr):
        d = self.d
        if d[addr:addr + 4] != b"OHDR" or d[addr + 4] != 2:
            raise ValueError(f"no version-2 object header at {addr:#x}")
        flags = d[addr + 5]
        p = addr + 6
        if flags & 0x20:
            p += 16
        if flags & 0x10:
            p += 4
        nsz = 1 << (flags & 3)
        size0 = int.from_bytes(d[p:p + nsz], "little")
        p += nsz
        self._check(addr, p - addr + size0)
        self._count("OHDR")
        out, blocks = [], [(p, p + size0)]
        while blocks:
            p, end = blocks.pop(0)
            while p + 4 <= end:
                mtype, msize, mflags = d[p], int.from_bytes(d[p + 1:p + 3], "little"), d[p + 3]
                p += 4 + (2 if flags & 0x04 else 0)
                body = d[p:p + msize]
                p += msize
                if mtype == 0x10:                            # continuation
                    caddr, clen = struct.unpack("<QQ", body[:16])
                    if d[caddr:caddr + 4] != b"OCHK":
                        raise ValueError("bad continuation block")
                    self._check(caddr, clen - 4)
                    blocks.append((caddr + 4, caddr + clen - 4))
                elif mtype != 0:
                    out.append((mtype, mflags, body))
        return out

    def _attr(self, body):
        ver = body[0]
        nlen, tlen, slen = struct.unpack_from("<HHH", body, 2)
        p = 8
        if ver == 3:
            p = 9
        pad = (lambda n: (n + 7) & ~7) if ver == 1 else (lambda n: n)
        name = body[p:p + nlen].split(b"\0")[0].decode()
        p += pad(nlen)
        dt, _ = decode_datatype(body[p:p + tlen])
        p += pad(tlen)
        shape = self._space(body[p:p + slen])
        p += pad(slen)
        n = int(np.prod(shape)) if shape else 1
        if isinstance(dt, tuple):
            if dt[0] != "vlen_str":
                return name, None
            vals = []
            for k in range(n):
                ln, gaddr, gidx = struct.unpack_from("<IQI", body, p + 16 * k)
                vals.append(self._global_heap(gaddr, gidx)[:ln].decode("utf-8"))
            return name, vals[0] if not shape else vals
        arr = np.frombuffer(body, dtype=dt, count=n, offset=p).reshape(shape)
        return name, (arr[()] if not shape else arr)

    def _global_heap(self, addr, idx):
        d = self.d
        if d[addr:addr + 4] != b"GCOL":
            raise ValueError("bad global heap collection")
        self._count("GCOL")
        size = struct.unpack_from("<Q", d, addr + 8)[0]
        p = addr + 16
        while p < addr + size:
            oid, _, _, osz = struct.unpack_from("<HHIQ", d, p)
            if oid == idx:
                return d[p + 16:p + 16 + osz]
            if oid == 0:
                break
            p += 16 + ((osz + 7) & ~7)
        raise ValueError("global heap object not found")

    @staticmethod
    def _space(b):
        ver, rank, flags = b[0], b[1], b[2]
        off = 4 if ver == 2 else 8
        return tuple(struct.unpack_from("<Q", b, off + 8 * k)[0] for k in range(rank))

    # ---- groups
    def _dense_links(self, heap_addr):
        """Link messages of a group in dense storage: the managed objects of the fractal heap, read from its direct
        blocks in order (a group of this size has a single root direct block)."""
        d = self.d
        if d[heap_addr:heap_addr + 4] != b"FRHP":
            raise ValueError("bad fractal heap header")
        self._count("FRHP")
        (heap_id_len, io_filter_len, flags) = struct.unpack_from("<HHB", d, heap_addr + 5)
        p = heap_addr + 10
        p += 4                                               # max size of managed objects
        p += 8 + 8                                           # next huge id, huge btree addr
        p += 8 + 8                                           # free space, free space manager addr
        p += 8 + 8 + 8 + 8                                   # managed space, allocated, iterator offset, n managed objects
        p += 8 + 8 + 8 + 8                                   # huge size / count, tiny size / count
        table_width, = struct.unpack_from("<H", d, p)
        start_block, max_direct = struct.unpack_from("<QQ", d, p + 2)
        max_heap_bits, start_rows, = struct.unpack_from("<HH", d, p + 18)
        root_addr, = struct.unpack_from("<Q", d, p + 22)
        cur_rows, = struct.unpack_from("<H", d, p + 30)
        hdr_len = p + 32 - heap_addr + (12 if io_filter_len else 0)
        self._check(heap_addr, hdr_len)
        off_bytes = (max_heap_bits + 7) // 8
        blocks = []                                          # (address, size) of the direct blocks, in heap order
        if cur_rows == 0:
            blocks.append((root_addr, start_block))
        else:
            if d[root_addr:root_addr + 4] != b"FHIB":
                raise ValueError("bad fractal heap indirect block")
            self._count("FHIB")
            max_direct_rows = (max_direct.bit_length() - start_block.bit_length()) + 2
            q = root_addr + 5 + 8 + off_bytes
            n_direct = min(cur_rows, max_direct_rows) * table_width
            n_indirect = max(cur_rows - max_direct_rows, 0) * table_width
            self._check(root_addr, q - root_addr + n_direct * (8 + (12 if io_filter_len else 0)) + n_indirect * 8)
            for k in range(n_direct):
                a, = struct.unpack_from("<Q", d, q)
                q += 8 + (12 if io_filter_len else 0)
                row = k // table_width
                if a != UNDEF:
                    blocks.append((a, start_block << max(row - 1, 0)))
            for k in range(n_indirect):
                a, = struct.unpack_from("<Q", d, q + 8 * k)
                if a != UNDEF:
                    raise ValueError("nested indirect blocks not supported")
        links = []
        for baddr, bsize in blocks:
            if d[baddr:baddr + 4] != b"FHDB":
                raise ValueError("bad fractal heap direct block")
            self._count("FHDB")
            head = 5 + 8 + off_bytes
            if flags & 2:                                    # checksummed direct blocks: computed with the checksum field zeroed
                blk = bytearray(d[baddr:baddr + bsize])
                want, = struct.unpack_from("<I", blk, head)
                blk[head:head + 4] = b"\0\0\0\0"
                if lookup3(bytes(blk)) != want:
                    raise ValueError(f"checksum mismatch in direct block at {baddr:#x}")
                head += 4
            q, end = baddr + head, baddr + bsize
            while q < end and d[q] == 1:                     # link message version 1
                name, addr, used = self._link(d[q:end])
                links.append((name, addr))
                q += used
        return links

    @staticmethod
    def _link(b):
        flags = b[1]
        p = 2
        ltype = 0
        if flags & 0x08:
            ltype = b[p]; p += 1
        if flags & 0x04:
            p += 8
        if flags & 0x10:
            p += 1
        n = 1 << (flags & 3)
        ln = int.from_bytes(b[p:p + n], "little"); p += n
        name = b[p:p + ln].decode(); p += ln
        if ltype != 0:
            raise ValueError("only hard links supported")
        addr, = struct.unpack_from("<Q", b, p)
        return name, addr, p + 8

    def _walk_group(self, addr, prefix):
        msgs = self._messages(addr)
        attrs, links, is_group = {}, [], False
        for mtype, _, body in msgs:
            if mtype == 0x0C:
                k, v = self._attr(body)
                attrs[k] = v
            elif mtype == 0x06:
                name, a, _ = self._link(body)
                links.append((name, a))
            elif mtype == 0x02:
                is_group = True
                flags = body[1]
                p = 2 + (8 if flags & 1 else 0)
                heap, _ = struct.unpack_from("<QQ", body, p)
                if heap != UNDEF:
                    self._count("dense link storage")
                    links += self._dense_links(heap)
        if not is_group:
            self._dataset(prefix.rstrip("/"), msgs, attrs)
            return attrs
        self.groups[prefix.rstrip("/")] = attrs
        for name, a in links:
            self._walk_group(a, prefix + name + "/")
        return attrs

    # ---- datasets
    def _dataset(self, name, msgs, attrs):
        shape = dtype = layout = None
        filters = []
        for mtype, _, body in msgs:
            if mtype == 0x01:
                shape = self._space(body)
            elif mtype == 0x03:
                dtype, _ = decode_datatype(body)
            elif mtype == 0x0B:
                ver, n = body[0], body[1]
                p = 2 if ver == 2 else 8
                for _ in range(n):
                    fid, = struct.unpack_from("<H", body, p)
                    p += 2
                    nlen = 0
                    if ver == 1 or fid >= 256:
                        nlen, = struct.unpack_from("<H", body, p)
                        p += 2
                    fl, ncd = struct.unpack_from("<HH", body, p)
                    p += 4 + ((nlen + 7) & ~7 if ver == 1 else nlen)
                    cd = struct.unpack_from("<" + "I" * ncd, body, p)
                    p += 4 * ncd + (4 if ver == 1 and ncd % 2 else 0)
                    filters.append((fid, cd))
            elif mtype == 0x08:
                layout = body
        self.objects[name] = H5Dataset(self, name, shape, dtype, layout, filters, attrs)

    def _read_dataset(self, ds: H5Dataset) -> np.ndarray:
        d, b = self.d, ds.layout
        ver, cls = b[0], b[1]
        if ver not in (3, 4):
            raise ValueError("data layout version not supported")
        n = int(np.prod(ds.shape)) if ds.shape else 1
        if cls == 1:                                         # contiguous
            addr, size = struct.unpack_from("<QQ", b, 2)
            if addr == UNDEF:
                return np.zeros(ds.shape, ds.dtype)
            return np.frombuffer(d, dtype=ds.dtype, count=n, offset=addr).reshape(ds.shape).copy()
        if cls == 0:                                         # compact
            size, = struct.unpack_from("<H", b, 2)
            return np.frombuffer(b, dtype=ds.dtype, count=n, offset=4).reshape(ds.shape).copy()
        if cls != 2 or ver != 4:
            raise ValueError("chunked layout: only version 4 supported")
        flags, ndim, enc = b[2], b[3], b[4]
        dims = [int.from_bytes(b[5 + k * enc:5 + (k + 1) * enc], "little") for k in range(ndim)]
        chunks = dims[:-1]
        p = 5 + ndim * enc
        itype = b[p]; p += 1
        grid = [-(-s // c) for s, c in zip(ds.shape, chunks)]
        n_chunks = int(np.prod(grid))
        filtered = bool(ds.filters)
        chunk_nbytes = int(np.prod(chunks)) * ds.dtype.itemsize
        entries = []
        if itype == 1:                                       # single chunk
            size, mask = chunk_nbytes, 0
            if flags & 0x02:
                size, mask = struct.unpack_from("<QI", b, p); p += 12
            addr, = struct.unpack_from("<Q", b, p)
            entries = [(addr, size, mask)]
            self._count("single-chunk index")
        elif itype == 2:                                     # implicit: chunks back to back
            addr, = struct.unpack_from("<Q", b, p)
            entries = [(addr + k * chunk_nbytes, chunk_nbytes, 0) for k in range(n_chunks)]
        elif itype == 3:                                     # fixed array
            page_bits = b[p]; p += 1
            fa, = struct.unpack_from("<Q", b, p)
            if d[fa:fa + 4] != b"FAHD":
                raise ValueError("bad fixed array header")
            self._check(fa, 24)
            self._count("FAHD")
            client, esize, pbits = d[fa + 5], d[fa + 6], d[fa + 7]
            nelem, dblk = struct.unpack_from("<QQ", d, fa + 8)
            if d[dblk:dblk + 4] != b"FADB":
                raise ValueError("bad fixed array data block")
            self._count("FADB")
            q = dblk + 14
            page = 1 << pbits

            def elements(q, count):
                out = []
                for k in range(count):
                    a, = struct.unpack_from("<Q", d, q)
                    if client == 1:
                        sz = int.from_bytes(d[q + 8:q + esize - 4], "little")
                        mask, = struct.unpack_from("<I", d, q + esize - 4)
                    else:
                        sz, mask = chunk_nbytes, 0
                    out.append((a, sz, mask))
                    q += esize
                return out
            if nelem <= page:
                self._check(dblk, 14 + nelem * esize)
                entries = elements(q, nelem)
            else:
                n_pages = -(-nelem // page)
                bm = (n_pages + 7) // 8
                self._check(dblk, 14 + bm)
                self._count("paged FADB")
                q += bm + 4
                left = nelem
                for pg in range(n_pages):
                    cnt = page if pg < n_pages - 1 else left
                    self._check(q, cnt * esize)
                    entries += elements(q, min(cnt, left))
                    q += cnt * esize + 4
                    left -= cnt
        else:
            raise ValueError(f"chunk index type {itype} not supported")
        out = np.zeros(ds.shape, ds.dtype)
        for k, (addr, size, mask) in enumerate(entries[:n_chunks]):
            if addr == UNDEF:
                continue
            raw = d[addr:addr + size]
            if filtered and not (mask & 1):
                for fid, _ in ds.filters:
                    if fid != 1:
                        raise ValueError(f"filter {fid} not supported")
                raw = zlib.decompress(raw)
            part = np.frombuffer(raw, dtype=ds.dtype, count=int(np.prod(chunks))).reshape(chunks)
            idx = np.unravel_index(k, grid)
            sl = tuple(slice(i * c, min((i + 1) * c, s)) for i, c, s in zip(idx, chunks, ds.shape))
            out[sl] = part[tuple(slice(0, s.stop - s.start) for s in sl)]
        return out
