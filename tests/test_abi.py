"""CPU-side checks of the drop-in boundary: the C-ABI library loads, exports every symbol the
header declares, and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import re
from pathlib import Path

import pytest

from fastore_b200 import _native as N
from fastore_b200 import build

ROOT = Path(__file__).resolve().parent.parent


@pytest.fixture(scope="module")
def lib():
    build.build_cuda()
    return N.cuda_lib()


def test_header_symbols_are_exported(lib):
    header = (ROOT / "include" / "fastore_b200.h").read_text()
    declared = set(re.findall(r"\b(fsb_[a-z_]+)\s*\(", header))
    assert declared == set(N.C_ABI_SYMBOLS), declared ^ set(N.C_ABI_SYMBOLS)
    for name in declared:
        assert hasattr(lib, name), f"{name} declared in the header but not exported"


def test_struct_layouts_match_header():
    assert C.sizeof(N.FsbParams) == 16
    assert N.RECORD_DTYPE.itemsize == 16
    assert N.BIN_DESC_DTYPE.itemsize == 64
    assert C.sizeof(N.FsbChunk) == 56
    assert C.sizeof(N.FsbBlock) == 15 * 8


def test_no_cpu_fallback(lib):
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    assert lib.fsb_device_count() == 0
    ctx = C.c_void_p()
    p = N.make_params()
    rc = lib.fsb_create(C.byref(p), 0, None, C.byref(ctx))
    assert rc == N.FSB_ERR_CUDA and not ctx.value
    assert b"no CPU fallback" in lib.fsb_last_error(None)


def test_parameter_validation_precedes_device_use(lib):
    ctx = C.c_void_p()
    for kw in (dict(signature_len=2), dict(signature_len=16), dict(quality_method=7)):
        p = N.make_params(**kw)
        assert lib.fsb_create(C.byref(p), 0, None, C.byref(ctx)) == N.FSB_ERR_PARAM
    p = N.make_params()
    p.dna_symbol_order = b"ACGNT"
    assert lib.fsb_create(C.byref(p), 0, None, C.byref(ctx)) == N.FSB_ERR_PARAM
