"""Host-side edge cases around the device path (CPU tier): CR-only line ends in the record-table sizing, and PE chunk
cuts whose two files are not at the same read when the buffer fills up."""
import ctypes as C

import numpy as np
import pytest

import binfile_helpers as BF
from fastore_b200 import _native as N
from fastore_b200 import synth


def test_record_table_capacity_counts_lone_cr_line_ends():
    """SkipLine (FastqParser.cpp:46-68) takes a lone CR as a line end, so fsh_max_records has to count it: a CR-only
    FASTQ must parse completely instead of stopping with FSH_STOP_CAPACITY after two records."""
    cfg = synth.synth_config(500, 60, seed=31)
    t1, _, r1, _ = synth.generate(cfg, threads=1)
    cr = np.frombuffer(t1.tobytes().replace(b"\n", b"\r"), dtype=np.uint8).copy()
    lib = N.host_lib()
    assert lib.fsh_max_records(N.np_ptr(cr), cr.size) >= 500
    recs, st = synth.parse_chunk(cr, keep_headers=True, keep_comments=True, quality_method=0)
    assert len(recs) == 500 and st.stop_reason == 0
    assert np.array_equal(recs["seq_len"], r1["seq_len"]) and np.array_equal(recs["seq_off"], r1["seq_off"])
    # mixed: CRLF everywhere still counts one line end per line
    crlf = np.frombuffer(t1.tobytes().replace(b"\n", b"\r\n"), dtype=np.uint8).copy()
    assert lib.fsh_max_records(N.np_ptr(crlf), crlf.size) == 500 + 2
    recs, st = synth.parse_chunk(crlf, keep_headers=True, keep_comments=True, quality_method=0)
    assert len(recs) == 500 and st.stop_reason == 0


@pytest.mark.skipif(not BF.have_ref_tools(), reason="oracle/_ref tools not built (no /root/reference here)")
def test_pe_chunks_stay_paired_when_the_cut_points_differ(tmp_path):
    """Mate files whose records differ in size (longer titles in file 2) reach the cut window at different reads.  The
    reference's re-synchronisation (FastqStream.cpp:166-189) advances its read-id counter once per skipped *line*, so it
    skips 4x too far in one file and then 12x in the other and pairs the wrong mates for the rest of the chunk (its
    ASSERT(rid_1 == rid_2) is compiled out).  Our reader deliberately does the arithmetic per *record*: chunks always hold
    the same reads in both files.  Checked by decoding our bin files with the reference's own decoder: every pair comes
    back with its own mate."""
    files = BF.write_fastq(tmp_path, "in", 16000, 100, True, 240)
    recs2 = BF.fastq_records(files[1])
    with open(files[1], "wb") as f:                      # mate 2 gets a comment: every record of file 2 is 24 bytes longer
        for t, s, q in recs2:
            f.write(t + b" mate=2 lane=7 tile=1234\n" + s + b"\n+\n" + q + b"\n")
    flags = dict(paired=True, b=2)
    sizes = BF.host_chain(files, tmp_path / "ours", flags, BF.oracle_producer)
    assert len(sizes) >= 3
    outs = [tmp_path / "dec_1.fastq", tmp_path / "dec_2.fastq"]
    BF.decode_with_reference(tmp_path / "ours", outs, True)
    want = sorted((a[1], a[2], b[1], b[2]) for a, b in zip(BF.fastq_records(files[0]), recs2))
    got = sorted((a[1], a[2], b[1], b[2]) for a, b in zip(BF.fastq_records(outs[0]), BF.fastq_records(outs[1])))
    assert got == want
