"""Python mirror of the reference interface for the hot path, over the C ABI.

Reference: `categorizer.Categorize(reads, dnaBins); packer.PackToBins(dnaBins, binBins)`
(BinModule.cpp:130-133).  Here: `GpuBinner(params).bin_chunks([chunk, ...]) -> [BinBlock, ...]`.
All compute happens in libfastore_b200.so on the GPU; a missing library or GPU raises.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _native as N


class FastoreError(RuntimeError):
    """Mirrors the reference's `Exception` (Exception.h:20-40): carries the library's message."""


@dataclass
class BinBlock:
    """One BinaryBinBlock (BinBlockData.h:62-180): four packed streams + per-bin descriptors."""
    meta: np.ndarray
    dna: np.ndarray
    qua: np.ndarray
    head: np.ndarray
    bins: np.ndarray            # BIN_DESC_DTYPE, ascending signature, N-bin last
    raw_dna_size: int
    raw_head_size: int
    n_records: int
    read_signature: np.ndarray | None = None
    read_info: np.ndarray | None = None


class GpuBinner:
    def __init__(self, params: N.FsbParams, device: int = 0, stream: int | None = None, per_read: bool = False,
                 profile: bool = False, sub_batch_records: int | None = None, validate: bool | None = None,
                 run_split: int | None = None, keep_comments: bool | None = None, keep_records: bool | None = None):
        self._lib = N.cuda_lib()
        self._ctx = C.c_void_p()
        self.params = params
        rc = self._lib.fsb_create(C.byref(params), device, C.c_void_p(stream) if stream else None, C.byref(self._ctx))
        if rc != N.FSB_OK:
            msg = self._lib.fsb_last_error(None)
            raise FastoreError(f"fsb_create failed ({rc}): {msg.decode() if msg else ''}")
        if per_read:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_PER_READ, 1))
        if profile:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_PROFILE, 1))
        if validate is not None:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_VALIDATE, 1 if validate else 0))
        if keep_comments is not None:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_KEEP_COMMENTS, 1 if keep_comments else 0))
        if keep_records is not None:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_KEEP_RECORDS, 1 if keep_records else 0))
        if run_split is not None:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_RUN_SPLIT, run_split))
        if sub_batch_records is not None:
            self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_SUBBATCH_RECORDS, sub_batch_records))

    # -- lifetime ---------------------------------------------------------------------------------
    def close(self):
        if getattr(self, "_ctx", None):
            self._lib.fsb_destroy(self._ctx)
            self._ctx = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *a):
        self.close()

    def _check(self, rc):
        if rc != N.FSB_OK:
            msg = self._lib.fsb_last_error(self._ctx)
            raise FastoreError(f"fastore_b200 error {rc}: {msg.decode() if msg else ''}")

    def set_run_split(self, n: int):
        """Sub-batches fsb_run cuts the next staged batch into (overlapped on two streams)."""
        self._check(self._lib.fsb_set_option(self._ctx, N.FSB_OPT_RUN_SPLIT, n))

    # -- the path ---------------------------------------------------------------------------------
    @staticmethod
    def _chunk_array(chunks):
        arr = (N.FsbChunk * len(chunks))()
        for i, ch in enumerate(chunks):
            arr[i] = ch
        return arr

    def bin_chunks(self, chunks) -> list[BinBlock]:
        """Host buffers in, host blocks out (H2D + kernels + D2H), one BinBlock per chunk."""
        arr = self._chunk_array(chunks)
        blocks = (N.FsbBlock * len(chunks))()
        self._check(self._lib.fsb_bin_chunks(self._ctx, arr, len(chunks), blocks))
        return [self._to_block(b) for b in blocks]

    def stage(self, chunks):
        self._n_staged = len(chunks)
        self._check(self._lib.fsb_stage(self._ctx, self._chunk_array(chunks), len(chunks)))

    def run(self):
        self._check(self._lib.fsb_run(self._ctx))

    def sync(self):
        self._check(self._lib.fsb_sync(self._ctx))

    def fetch(self, copy=True):
        blocks = (N.FsbBlock * self._n_staged)()
        self._check(self._lib.fsb_fetch(self._ctx, blocks, self._n_staged))
        return [self._to_block(b) for b in blocks] if copy else blocks

    def get_records(self, chunk: int, mate: int = 0) -> np.ndarray:
        """Device-side parse: the record table the library built for a chunk of the last call."""
        n = C.c_uint64()
        rc = self._lib.fsb_get_records(self._ctx, chunk, mate, None, 0, C.byref(n))
        if rc not in (N.FSB_OK, N.FSB_ERR_PARAM):
            self._check(rc)
        out = np.zeros(int(n.value), dtype=N.RECORD_DTYPE)
        self._check(self._lib.fsb_get_records(self._ctx, chunk, mate, N.np_ptr(out) if out.size else None, out.size, C.byref(n)))
        return out

    def find_new_minimizers(self, text: np.ndarray, records: np.ndarray, cur_signature: int, signature_parity: int):
        """DnaRebalancer::FindNewMinimizer for a table of reads (fastore_rebin's scan): (signature, info) arrays."""
        n = records.shape[0]
        sig = np.zeros(n, dtype=np.uint32)
        info = np.zeros(n, dtype=np.uint32)
        self._check(self._lib.fsb_find_new_minimizers(self._ctx, N.np_ptr(text), text.size, N.np_ptr(records), n, cur_signature, signature_parity,
                                                      N.np_ptr(sig), N.np_ptr(info)))
        return sig, info

    def stage_times(self, n_stages: int = 4):
        """Per-stage device milliseconds since the last call: the four stages of fsb_run, plus "check" (the input-check
        kernels of fsb_stage) when n_stages is 5."""
        ms = (C.c_float * n_stages)()
        runs = C.c_uint32()
        self._check(self._lib.fsb_stage_times(self._ctx, ms, n_stages, C.byref(runs)))
        return {n: float(ms[i]) for i, n in enumerate(N.FSB_STAGE_NAMES[:n_stages])}, int(runs.value)

    def stats(self) -> dict:
        s = N.FsbStats()
        self._check(self._lib.fsb_get_stats(self._ctx, C.byref(s)))
        return {f: int(getattr(s, f)) for f, _ in N.FsbStats._fields_}

    @staticmethod
    def _to_block(b) -> BinBlock:
        d = N.block_to_dict(b)
        return BinBlock(d["meta"], d["dna"], d["qua"], d["head"], d["bins"], d["raw_dna_size"], d["raw_head_size"],
                        d["n_records"], d["read_signature"], d["read_info"])
