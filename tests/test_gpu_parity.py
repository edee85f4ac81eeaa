"""GPU parity (driver runs these with -m gpu on a B200): the CUDA path, called through the C ABI,
must equal the CPU checkers bit for bit -- streams, descriptors, per-read (signature, pos, flags)."""
import numpy as np
import pytest

import oracle_helpers as O
from cases import CASES, make_case
from fastore_b200 import _native as N
from fastore_b200 import synth
from fastore_b200.binner import GpuBinner, FastoreError

pytestmark = pytest.mark.gpu


def gpu_block_dict(blk):
    return {"meta": blk.meta, "dna": blk.dna, "qua": blk.qua, "head": blk.head, "bins": blk.bins,
            "raw_dna_size": blk.raw_dna_size, "raw_head_size": blk.raw_head_size, "n_records": blk.n_records,
            "read_signature": blk.read_signature, "read_info": blk.read_info}


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_case_matches_oracle(name):
    params, chunk, keep = make_case(name)
    want = O.bin_chunk("orc", params, chunk)
    with GpuBinner(params, per_read=True) as g:
        got = gpu_block_dict(g.bin_chunks([chunk])[0])
    O.assert_blocks_equal(got, want, f"{name} (gpu vs port)")
    if O.have_reference():
        O.assert_blocks_equal(got, O.bin_chunk("ref", params, chunk), f"{name} (gpu vs compiled reference)")


@pytest.mark.parametrize("paired", [False, True])
def test_multi_chunk_batch(paired):
    """Several chunks in one pass of the pipeline: every chunk must still yield exactly its own block."""
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    keep, chunks = [], []
    for ci, (n, L) in enumerate([(5000, 100), (1, 100), (7000, 151), (3000, 36), (4097, 100)]):
        cfg = synth.synth_config(n, L, paired=paired, seed=900 + ci, first_index=ci * 100000, nrich=0.05, lowcomplex=0.05, alln=0.01, tie=0.02)
        t = synth.generate(cfg, threads=2)
        keep.append(t)
        chunks.append(N.make_chunk(t[0], t[2], t[1], t[3]))
    with GpuBinner(params, per_read=True) as g:
        got = g.bin_chunks(chunks)
        # the same context again with a different batch shape (buffers are reused)
        got2 = g.bin_chunks(chunks[::-1])
    for ci, ch in enumerate(chunks):
        want = O.bin_chunk("orc", params, ch)
        O.assert_blocks_equal(gpu_block_dict(got[ci]), want, f"chunk {ci}")
        O.assert_blocks_equal(gpu_block_dict(got2[len(chunks) - 1 - ci]), want, f"chunk {ci} (reversed batch)")


@pytest.mark.parametrize("paired", [False, True])
def test_pipelined_sub_batches(paired):
    """fsb_bin_chunks with every chunk as its own sub-batch: copies and kernels of neighbouring sub-batches
    overlap on three streams over two buffer sets; every chunk must still yield exactly its own block."""
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    keep, chunks = [], []
    for ci, (n, L) in enumerate([(6000, 100), (3, 100), (9000, 151), (2500, 36), (4097, 100), (7000, 120), (1, 50)]):
        cfg = synth.synth_config(n, L, paired=paired, seed=700 + ci, first_index=ci * 100000, nrich=0.05, lowcomplex=0.05, alln=0.01, tie=0.02)
        t = synth.generate(cfg, threads=2)
        keep.append(t)
        chunks.append(N.make_chunk(t[0], t[2], t[1], t[3]))
    want = [O.bin_chunk("orc", params, ch) for ch in chunks]
    with GpuBinner(params, per_read=True, sub_batch_records=1) as g:
        for rep in range(2):                               # second call reuses every buffer
            got = g.bin_chunks(chunks)
            for ci in range(len(chunks)):
                O.assert_blocks_equal(gpu_block_dict(got[ci]), want[ci], f"chunk {ci} (rep {rep})")
        # a contract violation in a later sub-batch stops the pipeline cleanly
        t1, t2, r1, r2 = keep[4]
        bad = r1.copy()
        bad["seq_off"][17] = t1.size
        broken = list(chunks)
        broken[4] = N.make_chunk(t1, bad, t2, r2)
        with pytest.raises(FastoreError):
            g.bin_chunks(broken)
        got = g.bin_chunks(chunks[:3])
        for ci in range(3):
            O.assert_blocks_equal(gpu_block_dict(got[ci]), want[ci], f"chunk {ci} (after error)")
    with GpuBinner(params, sub_batch_records=10000) as g:   # sub-batches of several chunks
        got = g.bin_chunks(chunks)
        for ci in range(len(chunks)):
            O.assert_blocks_equal(gpu_block_dict(got[ci]), want[ci], f"chunk {ci} (grouped)", per_read=False)


def test_stage_run_fetch_repeatable():
    params, chunk, keep = make_case("c2_pe150_lossless")
    want = O.bin_chunk("orc", params, chunk)
    with GpuBinner(params, per_read=True, profile=True) as g:
        g.stage([chunk])
        for _ in range(3):
            g.run()
        got = gpu_block_dict(g.fetch()[0])
        times, runs = g.stage_times()
        st = g.stats()
    O.assert_blocks_equal(got, want, "resident re-run")
    assert runs == 3 and all(v > 0 for v in times.values())
    assert st["kernel_launches"] > 0 and st["records"] == 3 * chunk.n_records


def test_input_contract_errors():
    params, chunk, keep = make_case("pe_three_records")
    t1, t2, r1, r2 = keep
    bad = r2.copy()
    bad["seq_len"][1] -= 1            # unequal mates: the reference only ASSERTs (FastqRecord.h:87)
    with GpuBinner(params) as g:
        with pytest.raises(FastoreError):
            g.bin_chunks([N.make_chunk(t1, r1, t2, bad)])
        bad2 = r1.copy()
        bad2["seq_off"][2] = t1.size    # offset outside the chunk
        with pytest.raises(FastoreError):
            g.bin_chunks([N.make_chunk(t1, bad2, t2, r2)])
        # the context is still usable afterwards
        blk = g.bin_chunks([chunk])[0]
        assert blk.n_records == 3


@pytest.mark.parametrize("paired", [False, True])
def test_bytes_outside_the_contract_are_rejected_on_the_device(paired):
    """FSB_OPT_VALIDATE (default on): a caller that wires the reference's parser straight to the C ABI gets FSB_ERR_INPUT
    for symbols outside ACGTN, qualities outside [offset, offset + 64) and 8-bit title characters instead of silently
    mis-coded streams (the reference itself leaves these undefined: FastqRecord.h:95-96, FastqPacker.cpp:250)."""
    cfg = synth.synth_config(700, 100, paired=paired, seed=77, nrich=0.05)
    t1, t2, r1, r2 = synth.generate(cfg)
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    binary = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired, quality_method=N.FSB_QUA_BINARY)

    def poke(text, off, value):
        t = text.copy()
        t[off] = value
        return t

    last = len(r1) - 1
    cases = [("lowercase base", 0, int(r1["seq_off"][5]) + 17, ord("a")), ("IUPAC base", 0, int(r1["seq_off"][last]) + 99, ord("R")),
             ("dot", 0, int(r1["seq_off"][300]), ord(".")), ("quality below the offset", 0, int(r1["qua_off"][123]) + 50, 32),
             ("quality above offset + 63", 0, int(r1["qua_off"][0]), 33 + 64), ("8-bit title character", 0, int(r1["head_off"][17]) + 3, 0xC3)]
    if paired:
        cases += [("lowercase base in mate 2", 1, int(r2["seq_off"][last]) + 99, ord("t")), ("mate 2 quality below the offset", 1, int(r2["qua_off"][9]), 10)]
    with GpuBinner(params) as g, GpuBinner(binary) as gb:
        for what, mate, off, value in cases:
            a, b = (poke(t1, off, value), t2) if mate == 0 else (t1, poke(t2, off, value))
            with pytest.raises(FastoreError, match="outside the input contract"):
                g.bin_chunks([N.make_chunk(a, r1, b, r2)])
        # the 1-bit mode only needs q >= offset (the threshold compare takes any value above it)
        hi = poke(t1, int(r1["qua_off"][0]), 33 + 64)
        assert gb.bin_chunks([N.make_chunk(hi, r1, t2, r2)])[0].n_records == len(r1)
        with pytest.raises(FastoreError, match="outside the input contract"):
            gb.bin_chunks([N.make_chunk(poke(t1, int(r1["qua_off"][0]), 32), r1, t2, r2)])
        # bytes next to the checked spans do not matter, the clean input passes, and the check can be switched off
        # (the poked copies are named: an fsb_chunk only holds pointers)
        beside = poke(t1, int(r1["seq_off"][5]) - 1, ord("\r"))
        assert g.bin_chunks([N.make_chunk(beside, r1, t2, r2)])[0].n_records == len(r1)
        assert g.bin_chunks([N.make_chunk(t1, r1, t2, r2)])[0].n_records == len(r1)
    lower = poke(t1, int(r1["seq_off"][5]) + 17, ord("a"))
    with GpuBinner(params, validate=False) as g:
        assert g.bin_chunks([N.make_chunk(lower, r1, t2, r2)])[0].n_records == len(r1)


def test_parser_table_equals_generator_table_on_gpu_path():
    """host parser -> C ABI: the table the parser builds drives the device exactly like the generator's."""
    cfg = synth.synth_config(3000, 100, paired=False, seed=33, header_comments=True)
    t1, _, r1, _ = synth.generate(cfg)
    recs, st = synth.parse_chunk(t1, keep_headers=True, keep_comments=False)
    params = N.make_params(signature_len=8, skip_zone_len=0)
    chunk = N.make_chunk(t1, recs)
    want = O.bin_chunk("orc", params, chunk)
    with GpuBinner(params, per_read=True) as g:
        got = gpu_block_dict(g.bin_chunks([chunk])[0])
    O.assert_blocks_equal(got, want, "-C headers")
    assert int(recs["head_len"].max()) < int(r1["head_len"].min())


@pytest.mark.parametrize("paired,split", [(True, 2), (False, 3), (True, 4), (True, 9)])
def test_run_split_gives_the_same_blocks(paired, split):
    """FSB_OPT_RUN_SPLIT: fsb_run cuts the staged batch into sub-batches of whole chunks and runs them on two streams with
    their own intermediates and their own regions of the result buffers (non-persistent K1 / K4 grids).  Every chunk
    must still yield exactly its own block, whatever the split (9 > chunks: one chunk per sub-batch), run after run."""
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    keep, chunks = [], []
    for ci, (n, L) in enumerate([(9000, 100), (1, 100), (12000, 151), (5000, 36), (4097, 100), (20000, 150), (300, 150)]):
        cfg = synth.synth_config(n, L, paired=paired, seed=950 + ci, first_index=ci * 100000, nrich=0.05, lowcomplex=0.05, alln=0.01, tie=0.02)
        t = synth.generate(cfg, threads=2)
        keep.append(t)
        chunks.append(N.make_chunk(t[0], t[2], t[1], t[3]))
    want = [O.bin_chunk("orc", params, ch) for ch in chunks]
    with GpuBinner(params, per_read=True, run_split=split) as g:
        g.stage(chunks)
        for rep in range(3):
            g.run()
            got = g.fetch()
            for ci in range(len(chunks)):
                O.assert_blocks_equal(gpu_block_dict(got[ci]), want[ci], f"split {split}, run {rep}, chunk {ci}")
        # back to one pass on the same context
        g.set_run_split(1)
        g.stage(chunks)
        g.run()
        got = g.fetch()
        for ci in range(len(chunks)):
            O.assert_blocks_equal(gpu_block_dict(got[ci]), want[ci], f"unsplit after split, chunk {ci}")


@pytest.mark.parametrize("paired", [False, True])
def test_many_chunks_in_one_batch(paired):
    """More than 32 chunks in one batch: K1 then looks the chunk of a record up by binary search in global memory
    instead of the ballot over lane-resident chunk tables, and the sort key carries six chunk bits."""
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    keep, chunks = [], []
    for ci in range(41):
        n = 1 if ci % 10 == 3 else 700 + 37 * ci
        cfg = synth.synth_config(n, 100 + (ci % 3) * 25, paired=paired, seed=300 + ci, first_index=ci * 10000, nrich=0.05, lowcomplex=0.05, alln=0.01, tie=0.02)
        t = synth.generate(cfg, threads=2)
        keep.append(t)
        chunks.append(N.make_chunk(t[0], t[2], t[1], t[3]))
    with GpuBinner(params, per_read=True) as g:
        g.stage(chunks)                 # one batch, whatever the sub-batch size of fsb_bin_chunks
        g.run()
        got = g.fetch()
    for ci, ch in enumerate(chunks):
        O.assert_blocks_equal(gpu_block_dict(got[ci]), O.bin_chunk("orc", params, ch), f"chunk {ci} of 41")
