"""Build the CUDA library in-tree: nvcc, sm_100a only, -lineinfo (so ncu source pages map back)."""
from __future__ import annotations

import os
import shutil
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libmgatk2_b200.so")
SOURCES = ("api.cu",)
def _deps():
    """Everything the CUDA library is compiled from: every .cu / .cuh under csrc/ and the public header."""
    names = [f for f in os.listdir(CSRC) if f.endswith((".cu", ".cuh"))]
    return [os.path.join(CSRC, f) for f in names] + [os.path.join(HERE, "..", "include", "mgatk2_b200.h")]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
              "-shared", "-Xcompiler", "-fPIC", "-cudart", "static"]


def nvcc_path() -> str:
    for cand in (shutil.which("nvcc"), "/usr/local/cuda/bin/nvcc"):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError("nvcc not found: cannot build libmgatk2_b200.so")


def is_stale() -> bool:
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(d) > t for d in _deps())


def build_bamio(force: bool = False) -> str:
    """Host-only BAM ingest library (g++, zlib): csrc/bamio.cpp -> libmgatk2_bamio.so."""
    from .bamio import build_bamio as _b
    return _b(force)


def build_textio(force: bool = False) -> str:
    """Host-only dense-plane text writer (g++, zlib): csrc/textio.cpp -> libmgatk2_textio.so."""
    from .textio import build_textio as _b
    return _b(force)


def build_extension(force: bool = False, verbose: bool = False) -> str:
    if force or is_stale():
        cmd = [nvcc_path(), *NVCC_FLAGS, "-o", LIB, *SOURCES]
        if verbose:
            cmd.insert(1, "-Xptxas=-v")
        subprocess.run(cmd, cwd=CSRC, check=True)
    return LIB
