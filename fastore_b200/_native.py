"""ctypes declarations for the two native libraries (include/fastore_b200.h, csrc/host/host_api.h).

Python is plumbing here: it loads the C ABI exactly as a reference-side binding would and never
computes any part of the path itself.
"""
from __future__ import annotations

import ctypes as C
import os
from pathlib import Path

import numpy as np

PKG = Path(__file__).resolve().parent
ROOT = PKG.parent

# ---- status codes / constants (include/fastore_b200.h) --------------------------------------------
FSB_OK, FSB_ERR_PARAM, FSB_ERR_INPUT, FSB_ERR_CUDA, FSB_ERR_NOMEM, FSB_ERR_STATE = range(6)
FSB_QUA_NONE, FSB_QUA_BINARY, FSB_QUA_8BIN, FSB_QUA_QVZ = range(4)
FSB_INFO_POS_MASK = 0x0000FFFF
FSB_INFO_REVERSE = 0x00010000
FSB_INFO_SWAPPED = 0x00020000
FSB_INFO_PLAIN_A = 0x00040000
FSB_INFO_PLAIN_B = 0x00080000
FSB_OPT_PER_READ, FSB_OPT_PROFILE, FSB_OPT_VALIDATE, FSB_OPT_SUBBATCH_RECORDS, FSB_OPT_RUN_SPLIT = 1, 2, 3, 4, 5
FSB_OPT_KEEP_COMMENTS, FSB_OPT_KEEP_RECORDS = 10, 11
FSB_STAGE_NAMES = ("ingest", "sort", "layout", "place", "check")     # "check": input-check kernels of fsb_stage, reported on request


class FsbParams(C.Structure):
    _fields_ = [
        ("signature_len", C.c_uint8),
        ("skip_zone_len", C.c_uint8),
        ("signature_mask_cutoff_bits", C.c_uint8),
        ("paired_end", C.c_uint8),
        ("quality_method", C.c_uint8),
        ("quality_offset", C.c_uint8),
        ("binary_threshold", C.c_uint8),
        ("reads_have_headers", C.c_uint8),
        ("dna_symbol_order", C.c_char * 5),
        ("reserved", C.c_uint8 * 3),
    ]


class FsbChunk(C.Structure):
    _fields_ = [
        ("text", C.c_void_p * 2),
        ("text_size", C.c_uint64 * 2),
        ("records", C.c_void_p * 2),
        ("n_records", C.c_uint64),
    ]


class FsbBlock(C.Structure):
    _fields_ = [
        ("meta", C.c_void_p), ("dna", C.c_void_p), ("qua", C.c_void_p), ("head", C.c_void_p),
        ("meta_size", C.c_uint64), ("dna_size", C.c_uint64), ("qua_size", C.c_uint64), ("head_size", C.c_uint64),
        ("raw_dna_size", C.c_uint64), ("raw_head_size", C.c_uint64),
        ("bins", C.c_void_p), ("n_bins", C.c_uint64), ("n_records", C.c_uint64),
        ("read_signature", C.c_void_p), ("read_info", C.c_void_p),
    ]


class OrcBlock(C.Structure):
    _fields_ = [
        ("meta", C.c_void_p), ("dna", C.c_void_p), ("qua", C.c_void_p), ("head", C.c_void_p),
        ("meta_size", C.c_uint64), ("dna_size", C.c_uint64), ("qua_size", C.c_uint64), ("head_size", C.c_uint64),
        ("raw_dna_size", C.c_uint64), ("raw_head_size", C.c_uint64),
        ("bins", C.c_void_p), ("n_bins", C.c_uint64), ("n_records", C.c_uint64),
        ("read_signature", C.c_void_p), ("read_info", C.c_void_p),
    ]


class FsbStats(C.Structure):
    _fields_ = [
        ("kernel_launches", C.c_uint64), ("h2d_bytes", C.c_uint64), ("d2h_bytes", C.c_uint64),
        ("records", C.c_uint64), ("algorithmic_bytes", C.c_uint64),
    ]


class FshParseStats(C.Structure):
    _fields_ = [
        ("n_records", C.c_uint64), ("min_seq_len", C.c_uint32), ("max_seq_len", C.c_uint32),
        ("consumed_bytes", C.c_uint64), ("stop_reason", C.c_uint32), ("invalid_records", C.c_uint32),
    ]


class FshBinConfig(C.Structure):
    _fields_ = [
        ("params", FsbParams), ("min_block_bin_size", C.c_uint32), ("keep_comments", C.c_uint8), ("verbose", C.c_uint8),
        ("reserved", C.c_uint8 * 2), ("fastq_block_size", C.c_uint64),
    ]


class FshSynthConfig(C.Structure):
    _fields_ = [
        ("seed", C.c_uint64), ("first_index", C.c_uint64), ("n_records", C.c_uint64),
        ("read_len", C.c_uint32), ("min_len", C.c_uint32), ("paired", C.c_uint32), ("genome_len", C.c_uint32),
        ("sub_rate_ppm", C.c_uint32), ("n_rate_ppm", C.c_uint32), ("nrich_ppm", C.c_uint32),
        ("lowcomplex_ppm", C.c_uint32), ("alln_ppm", C.c_uint32), ("tie_ppm", C.c_uint32),
        ("header_comments", C.c_uint32), ("crlf", C.c_uint32), ("qual_mean_x10", C.c_uint32),
        ("qual_sd_x10", C.c_uint32), ("reserved", C.c_uint32),
    ]


# numpy views of the plain structs
RECORD_DTYPE = np.dtype([("head_off", "<u4"), ("seq_off", "<u4"), ("qua_off", "<u4"),
                         ("seq_len", "<u2"), ("head_len", "u1"), ("reserved", "u1")])
BIN_DESC_DTYPE = np.dtype([(n, "<u8") for n in ("signature", "meta_size", "dna_size", "qua_size", "head_size",
                                                 "records_count", "raw_dna_size", "raw_head_size")])
assert RECORD_DTYPE.itemsize == 16 and BIN_DESC_DTYPE.itemsize == 64

# every symbol include/fastore_b200.h declares (checked by tests/test_abi.py)
C_ABI_SYMBOLS = (
    "fsb_create", "fsb_destroy", "fsb_last_error", "fsb_set_option", "fsb_bin_chunks", "fsb_stage", "fsb_run",
    "fsb_fetch", "fsb_sync", "fsb_stage_times", "fsb_get_stats", "fsb_host_alloc", "fsb_host_free", "fsb_device_count", "fsb_get_records", "fsb_find_new_minimizers",
)

_host = None
_cuda = None


def host_lib() -> C.CDLL:
    """libfastore_host.so (built on demand; pure host code)."""
    global _host
    if _host is None:
        from . import build
        lib = C.CDLL(str(build.build_host()))
        lib.fsh_parse_chunk.restype = C.c_int
        lib.fsh_parse_chunk.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int,
                                        C.c_void_p, C.c_uint64, C.POINTER(FshParseStats)]
        lib.fsh_parse_chunk_ex.restype = C.c_int
        lib.fsh_parse_chunk_ex.argtypes = [C.c_void_p, C.c_uint64, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                           C.c_void_p, C.c_uint64, C.POINTER(FshParseStats)]
        lib.fsh_titles_new.restype = C.c_void_p
        lib.fsh_titles_new.argtypes = []
        lib.fsh_titles_free.restype = None
        lib.fsh_titles_free.argtypes = [C.c_void_p]
        lib.fsh_titles_add.restype = C.c_int
        lib.fsh_titles_add.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.fsh_titles_consistent.restype = C.c_int
        lib.fsh_titles_consistent.argtypes = [C.c_void_p]
        lib.fsh_writer_merge_titles.restype = C.c_int
        lib.fsh_writer_merge_titles.argtypes = [C.c_void_p, C.c_void_p]
        lib.fsh_max_records.restype = C.c_uint64
        lib.fsh_max_records.argtypes = [C.c_void_p, C.c_uint64]
        lib.fsh_cut_position.restype = C.c_uint64
        lib.fsh_cut_position.argtypes = [C.c_void_p, C.c_uint64, C.c_uint64]
        lib.fsh_reader_open.restype = C.c_void_p
        lib.fsh_reader_open.argtypes = [C.POINTER(C.c_char_p), C.c_uint32, C.POINTER(C.c_char_p), C.c_uint32, C.c_uint64]
        lib.fsh_reader_next.restype = C.c_int
        lib.fsh_reader_next.argtypes = [C.c_void_p, C.c_void_p, C.POINTER(C.c_uint64), C.c_void_p, C.POINTER(C.c_uint64)]
        lib.fsh_reader_close.restype = None
        lib.fsh_reader_close.argtypes = [C.c_void_p]
        lib.fsh_writer_open.restype = C.c_void_p
        lib.fsh_writer_open.argtypes = [C.c_char_p, C.POINTER(FshBinConfig)]
        lib.fsh_writer_add_titles.restype = C.c_int
        lib.fsh_writer_add_titles.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64]
        lib.fsh_writer_add_block.restype = C.c_int
        lib.fsh_writer_add_block.argtypes = [C.c_void_p, C.c_void_p]
        lib.fsh_writer_close.restype = C.c_int
        lib.fsh_writer_close.argtypes = [C.c_void_p]
        lib.fsh_last_error.restype = C.c_char_p
        lib.fsh_last_error.argtypes = []
        lib.fsh_synth_size.restype = C.c_int
        lib.fsh_synth_size.argtypes = [C.POINTER(FshSynthConfig), C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]
        lib.fsh_synth_fill.restype = C.c_int
        lib.fsh_synth_fill.argtypes = [C.POINTER(FshSynthConfig), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int]
        _host = lib
    return _host


def cuda_lib() -> C.CDLL:
    """libfastore_b200.so: the C ABI.  Raises if it is not built -- there is no fallback."""
    global _cuda
    if _cuda is None:
        from . import build
        path = Path(os.environ["FSB_CUDA_LIB"]) if os.environ.get("FSB_CUDA_LIB") else build.CUDA_LIB      # (diagnostic builds)
        if not path.exists():
            raise RuntimeError(f"{path} is missing: run `python -m fastore_b200.build` (there is no CPU fallback)")
        lib = C.CDLL(str(path))
        lib.fsb_create.restype = C.c_int
        lib.fsb_create.argtypes = [C.POINTER(FsbParams), C.c_int, C.c_void_p, C.POINTER(C.c_void_p)]
        lib.fsb_destroy.restype = None
        lib.fsb_destroy.argtypes = [C.c_void_p]
        lib.fsb_last_error.restype = C.c_char_p
        lib.fsb_last_error.argtypes = [C.c_void_p]
        lib.fsb_set_option.restype = C.c_int
        lib.fsb_set_option.argtypes = [C.c_void_p, C.c_int, C.c_int64]
        lib.fsb_bin_chunks.restype = C.c_int
        lib.fsb_bin_chunks.argtypes = [C.c_void_p, C.POINTER(FsbChunk), C.c_uint32, C.POINTER(FsbBlock)]
        lib.fsb_stage.restype = C.c_int
        lib.fsb_stage.argtypes = [C.c_void_p, C.POINTER(FsbChunk), C.c_uint32]
        lib.fsb_run.restype = C.c_int
        lib.fsb_run.argtypes = [C.c_void_p]
        lib.fsb_fetch.restype = C.c_int
        lib.fsb_fetch.argtypes = [C.c_void_p, C.POINTER(FsbBlock), C.c_uint32]
        lib.fsb_sync.restype = C.c_int
        lib.fsb_sync.argtypes = [C.c_void_p]
        lib.fsb_stage_times.restype = C.c_int
        lib.fsb_stage_times.argtypes = [C.c_void_p, C.POINTER(C.c_float), C.c_uint32, C.POINTER(C.c_uint32)]
        lib.fsb_get_stats.restype = C.c_int
        lib.fsb_get_stats.argtypes = [C.c_void_p, C.POINTER(FsbStats)]
        lib.fsb_host_alloc.restype = C.c_void_p
        lib.fsb_host_alloc.argtypes = [C.c_size_t]
        lib.fsb_host_free.restype = None
        lib.fsb_host_free.argtypes = [C.c_void_p]
        lib.fsb_get_records.restype = C.c_int
        lib.fsb_get_records.argtypes = [C.c_void_p, C.c_uint32, C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint64)]
        lib.fsb_find_new_minimizers.restype = C.c_int
        lib.fsb_find_new_minimizers.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
        lib.fsb_device_count.restype = C.c_int
        lib.fsb_device_count.argtypes = []
        _cuda = lib
    return _cuda


def make_params(signature_len=8, skip_zone_len=0, paired_end=False, quality_method=FSB_QUA_NONE, quality_offset=33,
                binary_threshold=20, reads_have_headers=True, cutoff_bits=0) -> FsbParams:
    p = FsbParams()
    p.signature_len = signature_len
    p.skip_zone_len = skip_zone_len
    p.signature_mask_cutoff_bits = cutoff_bits
    p.paired_end = 1 if paired_end else 0
    p.quality_method = quality_method
    p.quality_offset = quality_offset
    p.binary_threshold = binary_threshold
    p.reads_have_headers = 1 if reads_have_headers else 0
    p.dna_symbol_order = b"ACGTN"
    return p


def np_ptr(a: np.ndarray) -> int:
    return a.ctypes.data


def make_chunk(text1: np.ndarray, rec1: np.ndarray, text2: np.ndarray | None = None, rec2: np.ndarray | None = None) -> FsbChunk:
    """fsb_chunk over numpy buffers (the caller keeps them alive).  rec1 = None: text alone, the library parses it on the device."""
    ch = FsbChunk()
    ch.text[0] = np_ptr(text1)
    ch.text_size[0] = text1.size
    if rec1 is not None:
        ch.records[0] = np_ptr(rec1)
        ch.n_records = rec1.shape[0]
    if text2 is not None:
        ch.text[1] = np_ptr(text2)
        ch.text_size[1] = text2.size
        if rec1 is not None:
            assert rec2 is not None and rec2.shape[0] == rec1.shape[0]
            ch.records[1] = np_ptr(rec2)
    return ch


def _copy(ptr, nbytes, dtype=np.uint8):
    if not ptr or nbytes == 0:
        return np.zeros(0, dtype=dtype)
    buf = (C.c_uint8 * int(nbytes)).from_address(ptr)
    return np.frombuffer(buf, dtype=dtype).copy()


def block_to_dict(b) -> dict:
    """Copy an fsb_block / orc_block into numpy arrays (the source memory may be reused afterwards)."""
    n = int(b.n_records)
    return {
        "meta": _copy(b.meta, b.meta_size), "dna": _copy(b.dna, b.dna_size),
        "qua": _copy(b.qua, b.qua_size), "head": _copy(b.head, b.head_size),
        "raw_dna_size": int(b.raw_dna_size), "raw_head_size": int(b.raw_head_size),
        "bins": _copy(b.bins, int(b.n_bins) * BIN_DESC_DTYPE.itemsize, np.uint8).view(BIN_DESC_DTYPE),
        "n_records": n,
        "read_signature": _copy(b.read_signature, 4 * n, np.uint8).view("<u4") if b.read_signature else None,
        "read_info": _copy(b.read_info, 4 * n, np.uint8).view("<u4") if b.read_info else None,
    }
