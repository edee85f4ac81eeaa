"""Loaders for the CPU checkers under oracle/ (TEST INFRASTRUCTURE -- only tests, smoke() and
bench.py's cpu_baseline / --impl reference legs may import this)."""
from __future__ import annotations

import ctypes as C
from pathlib import Path

import numpy as np

from fastore_b200 import _native as N

ROOT = Path(__file__).resolve().parent.parent
PORT_LIB = ROOT / "oracle" / "_build" / "liboracle.so"
REF_LIB = ROOT / "oracle" / "_ref" / "libfastore_ref.so"
REF_BIN = ROOT / "oracle" / "_ref" / "fastore_bin"

_libs = {}


def _load(kind: str):
    if kind not in _libs:
        path = PORT_LIB if kind == "orc" else REF_LIB
        if not path.exists():
            return None
        lib = C.CDLL(str(path))
        f = getattr(lib, f"{kind}_bin_chunk")
        f.restype = C.c_int
        f.argtypes = [C.POINTER(N.FsbParams), C.POINTER(N.FsbChunk), C.POINTER(N.OrcBlock)]
        g = getattr(lib, f"{kind}_block_free")
        g.restype = None
        g.argtypes = [C.POINTER(N.OrcBlock)]
        h = getattr(lib, f"{kind}_find_minimizer")
        h.restype = None
        h.argtypes = [C.POINTER(N.FsbParams), C.c_void_p, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        t = getattr(lib, f"{kind}_time_bin_chunk")
        t.restype = C.c_double
        t.argtypes = [C.POINTER(N.FsbParams), C.POINTER(N.FsbChunk), C.c_int, C.c_int]
        if kind == "orc":
            lib.orc_signature_valid.restype = C.c_int
            lib.orc_signature_valid.argtypes = [C.POINTER(N.FsbParams), C.c_uint32]
        _libs[kind] = lib
    return _libs[kind]


def have_reference() -> bool:
    return REF_LIB.exists()


def bin_chunk(kind: str, params: N.FsbParams, chunk: N.FsbChunk) -> dict:
    """kind: 'orc' (C port) or 'ref' (compiled reference).  Returns numpy copies of the block."""
    lib = _load(kind)
    if lib is None:
        raise RuntimeError(f"oracle library for '{kind}' is not built")
    blk = N.OrcBlock()
    rc = getattr(lib, f"{kind}_bin_chunk")(C.byref(params), C.byref(chunk), C.byref(blk))
    if rc != N.FSB_OK:
        raise RuntimeError(f"{kind}_bin_chunk failed: {rc}")
    d = N.block_to_dict(blk)
    getattr(lib, f"{kind}_block_free")(C.byref(blk))
    return d


def find_minimizer(kind: str, params: N.FsbParams, seq: bytes):
    lib = _load(kind)
    s, p = C.c_uint32(), C.c_uint32()
    buf = np.frombuffer(seq, dtype=np.uint8).copy()
    getattr(lib, f"{kind}_find_minimizer")(C.byref(params), N.np_ptr(buf), len(seq), C.byref(s), C.byref(p))
    return s.value, p.value


def time_bin_chunk(kind: str, params: N.FsbParams, chunk: N.FsbChunk, threads: int, reps: int = 1) -> float:
    lib = _load(kind)
    return float(getattr(lib, f"{kind}_time_bin_chunk")(C.byref(params), C.byref(chunk), threads, reps))


def assert_blocks_equal(a: dict, b: dict, what: str = "", per_read: bool = True):
    for s in ("meta", "dna", "qua", "head"):
        x, y = a[s], b[s]
        if x.size != y.size or not np.array_equal(x, y):
            n = min(x.size, y.size)
            diff = np.nonzero(x[:n] != y[:n])[0]
            first = int(diff[0]) if diff.size else n
            raise AssertionError(f"{what}: stream '{s}' differs: sizes {x.size} vs {y.size}, first differing byte {first}")
    assert a["bins"].shape == b["bins"].shape, f"{what}: bin count {a['bins'].shape} vs {b['bins'].shape}"
    for f in a["bins"].dtype.names:
        if not np.array_equal(a["bins"][f], b["bins"][f]):
            i = int(np.nonzero(a["bins"][f] != b["bins"][f])[0][0])
            raise AssertionError(f"{what}: descriptor field '{f}' differs at bin {i}: {a['bins'][i]} vs {b['bins'][i]}")
    assert a["raw_dna_size"] == b["raw_dna_size"] and a["raw_head_size"] == b["raw_head_size"], what
    assert a["n_records"] == b["n_records"], what
    if per_read and a.get("read_signature") is not None and b.get("read_signature") is not None:
        if not np.array_equal(a["read_signature"], b["read_signature"]):
            i = int(np.nonzero(a["read_signature"] != b["read_signature"])[0][0])
            raise AssertionError(f"{what}: signature of read {i}: {a['read_signature'][i]:#x} vs {b['read_signature'][i]:#x}")
        if not np.array_equal(a["read_info"], b["read_info"]):
            i = int(np.nonzero(a["read_info"] != b["read_info"])[0][0])
            raise AssertionError(f"{what}: info of read {i}: {a['read_info'][i]:#x} vs {b['read_info'][i]:#x}")


REBIN_REF_LIB = ROOT / "oracle" / "_ref" / "libfastore_ref_rebin.so"
_rebin = {}


def find_new_minimizer(kind: str, params: N.FsbParams, seq: bytes, cur: int, divisor: int):
    """DnaRebalancer::FindNewMinimizer of one read: kind 'orc' (C port) or 'ref' (the reference's member function)."""
    if kind not in _rebin:
        lib = C.CDLL(str(PORT_LIB if kind == "orc" else REBIN_REF_LIB))
        f = getattr(lib, "orc_find_new_minimizer" if kind == "orc" else "refrebin_find_new_minimizer")
        f.restype = None
        f.argtypes = [C.POINTER(N.FsbParams), C.c_void_p, C.c_uint32, C.c_uint32, C.c_uint32, C.POINTER(C.c_uint32), C.POINTER(C.c_uint32), C.POINTER(C.c_uint32)]
        _rebin[kind] = f
    s, p, r = C.c_uint32(), C.c_uint32(), C.c_uint32()
    buf = np.frombuffer(seq, dtype=np.uint8).copy()
    _rebin[kind](C.byref(params), N.np_ptr(buf), len(seq), cur, divisor, C.byref(s), C.byref(p), C.byref(r))
    return s.value, p.value, r.value


def new_minimizers_port(params: N.FsbParams, text: np.ndarray, recs: np.ndarray, cur: int, divisor: int):
    """(signature, info) arrays of the port for a record table, in the C ABI's format (info = pos | FSB_INFO_REVERSE)."""
    sig = np.zeros(len(recs), dtype=np.uint32)
    info = np.zeros(len(recs), dtype=np.uint32)
    raw = text.tobytes()
    for i, r in enumerate(recs):
        s, p, rev = find_new_minimizer("orc", params, raw[int(r["seq_off"]): int(r["seq_off"]) + int(r["seq_len"])], cur, divisor)
        sig[i] = s
        info[i] = p | (N.FSB_INFO_REVERSE if rev else 0)
    return sig, info
