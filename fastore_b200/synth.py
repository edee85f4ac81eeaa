"""Seeded synthetic FASTQ -- thin wrapper over csrc/host/synth.cpp."""
from __future__ import annotations

import ctypes as C
import os

import numpy as np

from . import _native as N


def synth_config(n_records, read_len, paired=False, seed=1, first_index=0, min_len=0, sub_rate=0.01, n_rate=0.005,
                 nrich=0.0, lowcomplex=0.0, alln=0.0, tie=0.0, header_comments=False, crlf=False,
                 genome_len=0) -> N.FshSynthConfig:
    c = N.FshSynthConfig()
    c.seed, c.first_index, c.n_records = seed, first_index, n_records
    c.read_len, c.min_len, c.paired, c.genome_len = read_len, min_len, 1 if paired else 0, genome_len
    c.sub_rate_ppm, c.n_rate_ppm = int(round(sub_rate * 1e6)), int(round(n_rate * 1e6))
    c.nrich_ppm, c.lowcomplex_ppm = int(round(nrich * 1e6)), int(round(lowcomplex * 1e6))
    c.alln_ppm, c.tie_ppm = int(round(alln * 1e6)), int(round(tie * 1e6))
    c.header_comments, c.crlf = int(header_comments), int(crlf)
    return c


def generate(cfg: N.FshSynthConfig, threads: int | None = None, with_tables=True, out=None):
    """Returns (text1, text2|None, records1|None, records2|None) as numpy arrays.

    `out` may be a callable (nbytes) -> writable uint8 ndarray (e.g. pinned memory) used for the text."""
    lib = N.host_lib()
    b1, b2 = C.c_uint64(), C.c_uint64()
    if lib.fsh_synth_size(C.byref(cfg), C.byref(b1), C.byref(b2)) != N.FSB_OK:
        raise ValueError("bad synthetic configuration")
    alloc = out or (lambda nbytes: np.empty(nbytes, dtype=np.uint8))
    t1 = alloc(b1.value)
    t2 = alloc(b2.value) if cfg.paired else None
    r1 = np.zeros(cfg.n_records, dtype=N.RECORD_DTYPE) if with_tables else None
    r2 = np.zeros(cfg.n_records, dtype=N.RECORD_DTYPE) if (with_tables and cfg.paired) else None
    threads = threads or min(32, os.cpu_count() or 1)
    rc = lib.fsh_synth_fill(C.byref(cfg), N.np_ptr(t1), N.np_ptr(t2) if t2 is not None else None,
                            N.np_ptr(r1) if r1 is not None else None, N.np_ptr(r2) if r2 is not None else None, threads)
    if rc != N.FSB_OK:
        raise ValueError(f"fsh_synth_fill failed: {rc}")
    return t1, t2, r1, r2


def parse_chunk(text: np.ndarray, keep_headers=True, keep_comments=True, quality_offset=33, quality_method=0,
                strict=True):
    """Host parser (csrc/host/fastq_parser.cpp): chunk text -> (record table, stats)."""
    lib = N.host_lib()
    cap = int(lib.fsh_max_records(N.np_ptr(text), text.size)) if text.size else 1
    recs = np.zeros(cap, dtype=N.RECORD_DTYPE)
    st = N.FshParseStats()
    rc = lib.fsh_parse_chunk(N.np_ptr(text), text.size, int(keep_headers), int(keep_comments), quality_offset,
                             quality_method, N.np_ptr(recs), cap, C.byref(st))
    if rc != N.FSB_OK and strict:
        raise ValueError(f"fsh_parse_chunk: {st.invalid_records} records outside the input contract")
    return recs[: st.n_records].copy(), st
