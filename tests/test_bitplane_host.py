"""Host checks of the word-parallel helpers of the bit-plane pileup (mgatk2_b200/csrc/bitplane.cuh).

The helpers are plain integer arithmetic and compile for the host: two small C++ programs compare them with the
per-base rule of the reference (pileup.py:67-86) — every int8 quality against every threshold, random SEQ words,
the 32x32 warp transpose (lanes simulated), whole reads through build_query_masks / query_window, and whole phase-B
rounds (one, two or four transposes for the four bit matrices of up to 32 candidate reads) against the obvious loop.
inflate_check does the same for the DEFLATE decoder of the BAM ingest (csrc/fast_inflate.h) against zlib."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("prog", ["bitplane_check", "qmask_check", "refplane_check", "round_check", "inflate_check"])
def test_bitplane_helpers_on_host(prog, tmp_path):
    gxx = shutil.which("g++")
    assert gxx, "g++ is part of the image"
    exe = tmp_path / prog
    subprocess.run([gxx, "-O2", "-std=c++17", "-o", str(exe), os.path.join(ROOT, "tests", "native", prog + ".cpp"), "-lz"], check=True)
    out = subprocess.run([str(exe)], capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout + out.stderr
    assert "ok" in out.stdout
