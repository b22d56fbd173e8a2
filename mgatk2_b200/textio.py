"""Native dense-plane text writer: ctypes binding of csrc/textio.cpp (include/mgatk2_textio.h).

`write_plane_file()` writes one `output.{A,C,G,T,coverage}.txt.gz` of the reference's IncrementalTextWriter
(src/file_io/writers.py:440-486) from the uint16 planes of a `PileupResult`, rows formatted and gzip-compressed on all
host cores (one gzip member per group of cells)."""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libmgatk2_textio.so")
SRC_PATH = os.path.join(HERE, "csrc", "textio.cpp")
EXPORTS = ("mgatk_text_write_plane_file", "mgatk_text_error")
_lib = None


def build_textio(force: bool = False) -> str:
    """g++ -O2 -shared -fPIC csrc/textio.cpp -lz -pthread -> libmgatk2_textio.so (in-tree, host only)."""
    stale = not os.path.exists(LIB_PATH) or os.path.getmtime(SRC_PATH) > os.path.getmtime(LIB_PATH)
    if force or stale:
        subprocess.run(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-o", LIB_PATH, SRC_PATH, "-lz", "-pthread"], check=True)
    return LIB_PATH


def load():
    global _lib
    if _lib is None:
        build_textio()
        lib = ctypes.CDLL(LIB_PATH)
        lib.mgatk_text_write_plane_file.argtypes = [
            ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int32, ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64,
            ctypes.c_int32, ctypes.c_int32, ctypes.c_void_p, ctypes.c_int64, ctypes.c_char_p, ctypes.c_void_p, ctypes.c_int32,
            ctypes.c_int32, ctypes.POINTER(ctypes.c_int64)]
        lib.mgatk_text_error.restype = ctypes.c_char_p
        _lib = lib
    return _lib


def write_plane_file(path, planes: np.ndarray, mito_length: int, overflow: np.ndarray | None, plane_a: int, plane_b: int,
                     cells, barcodes: list[str], level: int = 9, threads: int = 0) -> int:
    """Rows "pos,barcode,a[,b]" of the listed cells (indices `cells`, names `barcodes`, in that order) where a or b is
    non-zero; plane_b = -1 writes three columns. Returns the number of rows."""
    lib = load()
    planes = np.ascontiguousarray(planes, dtype=np.uint16)
    if planes.ndim != 3:
        raise ValueError("planes must be [n_cells, n_planes, pos_pad]")
    cells = np.ascontiguousarray(cells, dtype=np.int32)
    if len(cells) != len(barcodes):
        raise ValueError("one barcode per listed cell")
    enc = [b.encode() for b in barcodes]
    ends = np.cumsum([len(b) for b in enc], dtype=np.int64) if enc else np.zeros(0, np.int64)
    names = b"".join(enc)
    ovf = None if overflow is None or len(overflow) == 0 else np.ascontiguousarray(overflow)
    rows = ctypes.c_int64(0)
    rc = lib.mgatk_text_write_plane_file(
        os.fsencode(str(path)), planes.ctypes.data, planes.shape[0], planes.shape[2], int(mito_length),
        ovf.ctypes.data if ovf is not None else None, 0 if ovf is None else len(ovf), int(plane_a), int(plane_b),
        cells.ctypes.data if len(cells) else None, len(cells), names, ends.ctypes.data if len(ends) else None,
        int(level), int(threads), ctypes.byref(rows))
    if rc:
        raise OSError(lib.mgatk_text_error().decode() or "mgatk_text_write_plane_file failed")
    return rows.value
