"""CPU-side checks of the C-ABI library: it loads, exports every symbol include/mgatk2_b200.h declares,
and answers the calls that need no GPU."""
import ctypes
import os
import re

import pytest

from tests.helpers import ROOT


@pytest.fixture(scope="module")
def lib():
    from mgatk2_b200.build import build_extension
    build_extension()
    from mgatk2_b200 import _lib
    return _lib.load()


def test_exports_match_header(lib):
    header = open(os.path.join(ROOT, "include", "mgatk2_b200.h")).read()
    declared = set(re.findall(r"\b(mgatk_[a-z0-9_]+)\s*\(", header))
    declared -= {"mgatk_status"}
    from mgatk2_b200 import _lib
    assert declared == set(_lib.EXPORTS)
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_bamio_exports_match_header():
    header = open(os.path.join(ROOT, "include", "mgatk2_bamio.h")).read()
    declared = set(re.findall(r"\b(mgatk_ba[mi]_[a-z_]+)\s*\(", header))
    from mgatk2_b200 import bamio
    assert declared == set(bamio.EXPORTS)
    lib = bamio.load()
    for sym in declared:
        assert hasattr(lib, sym), sym


def test_struct_sizes():
    from mgatk2_b200._lib import OutputsC, ParamsC
    from mgatk2_b200.batch import MgatkBatchC
    from mgatk2_b200.engine import CELL_QC_DTYPE, OVERFLOW_DTYPE
    assert ctypes.sizeof(ParamsC) == 48 and ctypes.sizeof(MgatkBatchC) == 88 and ctypes.sizeof(OutputsC) == 48
    assert CELL_QC_DTYPE.itemsize == 32 and OVERFLOW_DTYPE.itemsize == 12


def test_no_gpu_calls(lib):
    assert lib.mgatk_abi_version() == 2
    assert lib.mgatk_status_string(4) == b"records are not sorted by reference_start"
    assert lib.mgatk_workspace_bytes(1_000_000, 2000, 50) > 64_000_000
    assert lib.mgatk_workspace_bytes(1_000_000, 2000, 150) > 2 * 96 * 1_000_000
    assert lib.mgatk_workspace_bytes(-1, 5, 50) == -1
    import torch
    if not torch.cuda.is_available():
        h = ctypes.c_void_p()
        assert lib.mgatk_create(ctypes.byref(h), 0) == 7      # MGATK_ERR_NO_DEVICE: fails loudly, no fallback
        from mgatk2_b200.engine import PileupEngine
        from mgatk2_b200.exceptions import PileupKernelError
        with pytest.raises(PileupKernelError):
            PileupEngine(0)
