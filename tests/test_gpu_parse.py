"""Device-side FASTQ parse (-m gpu): chunks handed to the C ABI as text alone must give the record tables of the
host parser -- itself pinned to SingleFastqRecordParser::ReadNextRecord (FastqParser.cpp:118-165) by the whole-file tests -- and
therefore the same blocks, for every kind of line end, for texts that stop short, and through both interfaces."""
import numpy as np
import pytest

import oracle_helpers as O
from fastore_b200 import _native as N
from fastore_b200 import synth
from fastore_b200.binner import GpuBinner, FastoreError
from test_gpu_parity import gpu_block_dict

pytestmark = pytest.mark.gpu


def host_tables(texts, keep_comments=True, keep_headers=True, strict=True):
    recs = [synth.parse_chunk(t, keep_headers=keep_headers, keep_comments=keep_comments, strict=strict)[0] for t in texts]
    n = min(len(r) for r in recs)
    return [r[:n].copy() for r in recs]


def variants(raw: bytes):
    lines = raw.split(b"\n")
    mixed = b"".join(ln + (b"\n", b"\r\n", b"\r")[(i // 4) % 3] for i, ln in enumerate(lines[:-1]))
    cut = raw[: (len(raw) * 2) // 3]
    blank = raw.replace(b"\n@", b"\n\n@", 1) if raw.count(b"\n@") > 40 else raw
    k = raw.find(b"\n@", len(raw) // 2)
    return {"lf": raw, "crlf": raw.replace(b"\n", b"\r\n"), "cr": raw.replace(b"\n", b"\r"), "mixed": mixed, "no_final_eol": raw[:-1],
            "cut_short": cut, "blank_line_stops": raw[:k] + b"\n" + raw[k:], "title_without_at": raw[:k + 1] + b"X" + raw[k + 2:], "only_eol": b"\n",
            "empty_plus": raw[:k + 1] + raw[k + 1:].replace(b"\n+\n", b"\n\n", 1)}


@pytest.mark.parametrize("keep_comments", [True, False])
def test_device_tables_equal_the_host_parser(keep_comments):
    cfg = synth.synth_config(3000, 100, seed=51, header_comments=True, nrich=0.03)
    t1, _, _, _ = synth.generate(cfg, threads=2)
    params = N.make_params(signature_len=8, skip_zone_len=0)
    with GpuBinner(params, keep_comments=keep_comments) as g:
        for name, txt in variants(t1.tobytes()).items():
            text = np.frombuffer(txt, dtype=np.uint8).copy()
            want = host_tables([text], keep_comments=keep_comments)[0]
            g.stage([N.make_chunk(text, None)])
            got = g.get_records(0)
            assert got.shape == want.shape, f"{name}: {got.shape[0]} records vs {want.shape[0]}"
            for f in N.RECORD_DTYPE.names:
                assert np.array_equal(got[f], want[f]), f"{name}: field {f}"
            if want.shape[0]:
                g.run()
                blk = gpu_block_dict(g.fetch()[0])
                O.assert_blocks_equal(blk, O.bin_chunk("orc", params, N.make_chunk(text, want)), f"{name}: block", per_read=False)


def test_paired_chunks_keep_the_shorter_count_and_many_chunks_share_a_batch():
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=True)
    keep, chunks, wants, tables = [], [], [], []
    for ci, (n, L) in enumerate([(4000, 100), (1, 100), (2500, 151), (700, 36)]):
        cfg = synth.synth_config(n, L, paired=True, seed=960 + ci, first_index=ci * 100000, nrich=0.05)
        t1, t2, _, g2 = synth.generate(cfg, threads=2)
        if ci == 2:                                     # file 2 of this chunk stops 11 records early: the pair count is the shorter one
            t2 = t2[: int(g2["head_off"][n - 11]) - 1].copy()
        keep.append((t1, t2))
        r1, r2 = host_tables([t1, t2])
        assert len(r1) == (n - 11 if ci == 2 else n)
        wants.append(N.make_chunk(t1, r1, t2, r2)); tables.append((r1, r2))
        chunks.append(N.make_chunk(t1, None, t2))
    with GpuBinner(params, keep_records=True) as g:
        g.stage(chunks)
        g.run()
        res = g.fetch()
        for ci in range(len(chunks)):
            O.assert_blocks_equal(gpu_block_dict(res[ci]), O.bin_chunk("orc", params, wants[ci]), f"resident, chunk {ci}", per_read=False)
        # the pipelined call: one chunk per sub-batch, tables copied back
        g._check(g._lib.fsb_set_option(g._ctx, N.FSB_OPT_SUBBATCH_RECORDS, 1))
        res = g.bin_chunks(chunks)
        for ci in range(len(chunks)):
            O.assert_blocks_equal(gpu_block_dict(res[ci]), O.bin_chunk("orc", params, wants[ci]), f"pipelined, chunk {ci}", per_read=False)
            got1, got2 = g.get_records(ci, 0), g.get_records(ci, 1)
            assert got1.shape[0] == int(wants[ci].n_records) == got2.shape[0]
            assert np.array_equal(got1, tables[ci][0]) and np.array_equal(got2, tables[ci][1]), f"pipelined, tables of chunk {ci}"


def test_records_outside_the_contract_are_errors():
    cfg = synth.synth_config(200, 100, seed=52)
    t1, _, _, _ = synth.generate(cfg, threads=1)
    raw = t1.tobytes()
    k = raw.find(b"\n", raw.find(b"\n@", len(raw) // 2) + 2) + 1       # start of a sequence line in the middle
    e = raw.find(b"\n", k)
    long_read = raw[:k] + b"A" * 300 + raw[e:]
    q0 = raw.find(b"\n+\n", e) + 3
    long_read = long_read[: q0 + 200] + b"I" * 300 + long_read[long_read.find(b"\n", q0 + 200):]
    params = N.make_params(signature_len=8, skip_zone_len=0)
    with GpuBinner(params) as g:
        text = np.frombuffer(long_read, dtype=np.uint8).copy()
        recs, st = synth.parse_chunk(text, strict=False)
        if st.invalid_records:                           # the host parser calls it out: so must the device
            with pytest.raises(FastoreError):
                g.stage([N.make_chunk(text, None)])
        # mixing the two ways of handing chunks over is refused
        good = np.frombuffer(raw, dtype=np.uint8).copy()
        r, _ = synth.parse_chunk(good)
        with pytest.raises(FastoreError):
            g.stage([N.make_chunk(good, None), N.make_chunk(good, r)])
        g.stage([N.make_chunk(good, None)])
        assert g.get_records(0).shape[0] == 200


@pytest.mark.parametrize("eol", [b"\r\n", b"\r", b"\n"])
def test_line_ends_on_vector_warp_and_tile_boundaries(eol):
    """The line-end pass looks at 16-byte vectors, 32 of them per warp, 16 KB per block: put a line end (and the CR of a CR LF)
    on the last byte in front of each of these boundaries, and one byte to either side."""
    cfg = synth.synth_config(400, 100, seed=53)
    t1, _, _, _ = synth.generate(cfg, threads=1)
    raw = t1.tobytes()
    params = N.make_params(signature_len=8, skip_zone_len=0)
    with GpuBinner(params) as g:
        for boundary in (16, 512, 16384, 32768):
            for delta in (-1, 0, 1):
                txt = raw.replace(b"\n", eol)
                p = txt.find(eol, boundary - 1 - 150 if boundary > 200 else 0)      # a line end shortly in front of the boundary ...
                pad = boundary - 1 + delta - p                                     # ... moved so that its first byte sits at boundary - 1 + delta
                if pad < 0:
                    p = txt.find(eol, p + 1); pad = boundary - 1 + delta - p
                assert 0 <= pad < 200
                txt = txt[:1] + b"x" * pad + txt[1:]                               # a longer first title shifts everything behind it
                assert txt[boundary - 1 + delta: boundary - 1 + delta + len(eol)] == eol
                text = np.frombuffer(txt, dtype=np.uint8).copy()
                want = host_tables([text])[0]
                assert want.shape[0] == 400
                g.stage([N.make_chunk(text, None)])
                got = g.get_records(0)
                assert got.shape == want.shape, f"boundary {boundary}{delta:+d}: {got.shape[0]} records"
                for f in N.RECORD_DTYPE.names:
                    assert np.array_equal(got[f], want[f]), f"boundary {boundary}{delta:+d}: field {f}"
