"""The per-thread routines of the CUDA kernels (csrc/*_core.cuh) are plain integer code that also
compiles for the host.  tests/emul/emul.cpp runs them "thread" by "thread" over a chunk; here their
output must equal the oracle's.  This checks kernel logic in the CPU tier; the product never runs
this code on the host (the GPU parity tests call the real kernels through the C ABI)."""
import ctypes as C
import subprocess
from pathlib import Path

import numpy as np
import pytest

import oracle_helpers as O
from cases import CASES, make_case
from fastore_b200 import _native as N
from fastore_b200 import synth

ROOT = Path(__file__).resolve().parent.parent
EMUL_SRC = ROOT / "tests" / "emul" / "emul.cpp"
EMUL_LIB = ROOT / "tests" / "emul" / "_build" / "libemul.so"


@pytest.fixture(scope="module")
def emul():
    deps = [EMUL_SRC] + list((ROOT / "fastore_b200" / "csrc").glob("*.cuh")) + [ROOT / "include" / "fastore_b200.h"]
    if not EMUL_LIB.exists() or any(d.stat().st_mtime > EMUL_LIB.stat().st_mtime for d in deps):
        EMUL_LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-x", "c++", str(EMUL_SRC), "-o", str(EMUL_LIB)], check=True)
    lib = C.CDLL(str(EMUL_LIB))
    lib.emul_signatures.restype = C.c_int
    lib.emul_signatures.argtypes = [C.POINTER(N.FsbParams), C.POINTER(N.FsbChunk), C.c_void_p, C.c_void_p]
    return lib


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_signature_core_matches_oracle(emul, name):
    params, chunk, keep = make_case(name)
    want = O.bin_chunk("orc", params, chunk)
    n = int(chunk.n_records)
    sig = np.zeros(n, dtype=np.uint32)
    info = np.zeros(n, dtype=np.uint32)
    assert emul.emul_signatures(C.byref(params), C.byref(chunk), N.np_ptr(sig), N.np_ptr(info)) == N.FSB_OK
    bad = np.nonzero(sig != want["read_signature"])[0]
    assert bad.size == 0, f"{name}: signature of read {bad[0]}: {sig[bad[0]]:#x} vs {want['read_signature'][bad[0]]:#x} ({bad.size} differ)"
    bad = np.nonzero(info != want["read_info"])[0]
    assert bad.size == 0, f"{name}: info of read {bad[0]}: {info[bad[0]]:#x} vs {want['read_info'][bad[0]]:#x} ({bad.size} differ)"


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_pack_core_matches_oracle(emul, name):
    params, chunk, keep = make_case(name)
    want = O.bin_chunk("orc", params, chunk)
    emul.emul_pack.restype = C.c_int
    emul.emul_pack.argtypes = [C.POINTER(N.FsbParams), C.POINTER(N.FsbChunk)] + [C.c_void_p] * 5
    bufs = [np.zeros(want[s].size + 64, dtype=np.uint8) for s in ("meta", "dna", "qua", "head")]
    sizes = np.zeros(4, dtype=np.uint64)
    assert emul.emul_pack(C.byref(params), C.byref(chunk), *[N.np_ptr(b) for b in bufs], N.np_ptr(sizes)) == N.FSB_OK
    for k, s in enumerate(("meta", "dna", "qua", "head")):
        assert int(sizes[k]) == want[s].size, f"{name}: stream {s} size {int(sizes[k])} vs {want[s].size}"
        got = bufs[k][: want[s].size]
        bad = np.nonzero(got != want[s])[0]
        assert bad.size == 0, f"{name}: stream {s} differs at byte {bad[0]} of {want[s].size} ({bad.size} bytes differ)"


@pytest.mark.parametrize("name", [c[0] for c in CASES])
@pytest.mark.parametrize("per_thread,threads", [(32, 128), (3, 5), (1, 7), (32, 4)])
def test_one_scan_layout_matches_the_sequential_walk(emul, name, per_thread, threads):
    """layout_core.cuh: the framing rules as an associative scan (run states, block states, exclusive scan, absolute
    walk) give every record the bit positions and every bin the descriptor of the plain sequential walk, for any
    decomposition into threads and blocks.  Cases whose reads differ in length are outside this form (they take the
    general layout kernels) and are skipped here."""
    params, chunk, keep = make_case(name)
    emul.emul_layout_fused.restype = C.c_long
    emul.emul_layout_fused.argtypes = [C.POINTER(N.FsbParams), C.POINTER(N.FsbChunk), C.c_uint32, C.c_uint32]
    bad = emul.emul_layout_fused(C.byref(params), C.byref(chunk), per_thread, threads)
    if bad == -1:
        pytest.skip("reads of different lengths")
    assert bad == 0, f"{name}: {bad} mismatches"


def test_line_end_masks_follow_skipline(emul):
    """parse_core.cuh finds line ends four bytes at a time; here against the byte-by-byte rule of SkipLine
    (FastqParser.cpp:46-68): LF ends a line, CR ends a line unless an LF follows it."""
    emul.emul_line_end_masks.restype = None
    emul.emul_line_end_masks.argtypes = [C.c_void_p, C.c_uint64, C.c_void_p]
    rng = np.random.default_rng(5)
    alphabet = np.frombuffer(b"\n\r\n\rACGTN@+I!~\x00\x0b\x0c\x0e\x8a\x8d\xff", dtype=np.uint8)      # line ends often, and their near misses
    for size in (0, 1, 15, 16, 17, 31, 32, 33, 4096, 100003):
        text = alphabet[rng.integers(0, alphabet.size, size)].copy()
        masks = np.zeros((size + 15) // 16 + 1, dtype=np.uint16)
        emul.emul_line_end_masks(N.np_ptr(text), size, N.np_ptr(masks))
        nxt = np.append(text[1:], 0) if size else text
        ends = (text == 10) | ((text == 13) & (nxt != 10))
        padded = np.zeros(((size + 15) // 16) * 16, dtype=bool)
        padded[:size] = ends
        want = (padded.reshape(-1, 16) * (1 << np.arange(16))).sum(axis=1).astype(np.uint16) if size else np.zeros(0, np.uint16)
        assert np.array_equal(masks[: want.size], want), f"size {size}"


@pytest.mark.parametrize("keep_comments", [True, False])
def test_device_parse_rules_equal_the_host_parser(emul, keep_comments):
    """The three passes of the device-side parse (parse_core.cuh, run here on the host) against the host parser -- itself
    pinned to SingleFastqRecordParser::ReadNextRecord by the whole-file tests -- for every kind of line end and for texts
    that stop short.  tests/test_gpu_parse.py repeats this through the kernels."""
    from test_gpu_parse import variants
    emul.emul_parse_text.restype = C.c_uint64
    emul.emul_parse_text.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_uint64)]
    cfg = synth.synth_config(1500, 100, seed=51, header_comments=True, nrich=0.03)
    t1, _, _, _ = synth.generate(cfg, threads=2)
    for name, txt in variants(t1.tobytes()).items():
        text = np.frombuffer(txt, dtype=np.uint8).copy()
        want, _ = synth.parse_chunk(text, keep_headers=True, keep_comments=keep_comments, strict=True)
        got = np.zeros(max(1, len(txt) // 8), dtype=N.RECORD_DTYPE)
        reason, bad = C.c_uint32(), C.c_uint64()
        n = emul.emul_parse_text(N.np_ptr(text), text.size, 1, int(keep_comments), N.np_ptr(got), got.shape[0], C.byref(reason), C.byref(bad))
        assert n == want.shape[0], f"{name}: {n} records vs {want.shape[0]}"
        assert bad.value == 2**64 - 1, name
        for f in N.RECORD_DTYPE.names:
            assert np.array_equal(got[f][:n], want[f]), f"{name}: field {f}"
