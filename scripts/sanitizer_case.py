"""Small end-to-end case for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): PE and SE multi-chunk batches through
the pipelined call, a split resident run, the device-side parse and the rebin signature scan, each checked against the port."""
import os
import sys
sys.path.insert(0, '/root/repo'); sys.path.insert(0, '/root/repo/tests')
import numpy as np
import oracle_helpers as O
from fastore_b200 import _native as N
from fastore_b200 import synth
from fastore_b200.binner import GpuBinner


def blockdict(blk):
    return {"meta": blk.meta, "dna": blk.dna, "qua": blk.qua, "head": blk.head, "bins": blk.bins, "raw_dna_size": blk.raw_dna_size, "raw_head_size": blk.raw_head_size,
            "n_records": blk.n_records, "read_signature": blk.read_signature, "read_info": blk.read_info}


for paired in (True, False):
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    keep, chunks, texts = [], [], []
    sizes = [(3000, 150), (1, 100), (2500, 151), (700, 36)] if not os.environ.get('FSB_SAN_SMALL') else [(400, 150), (1, 100), (300, 151), (90, 36)]
    for ci, (n, L) in enumerate(sizes):
        cfg = synth.synth_config(n, L, paired=paired, seed=900 + ci, first_index=ci * 100000, nrich=0.05, lowcomplex=0.05, alln=0.01, tie=0.02)
        t = synth.generate(cfg, threads=2); keep.append(t); chunks.append(N.make_chunk(t[0], t[2], t[1], t[3])); texts.append(N.make_chunk(t[0], None, t[1]))
    want = [O.bin_chunk("orc", params, ch) for ch in chunks]
    with GpuBinner(params, per_read=True, sub_batch_records=1) as g:
        got = g.bin_chunks(chunks)                       # pipelined: one chunk per sub-batch, three buffer sets
        for ci in range(len(chunks)):
            O.assert_blocks_equal(blockdict(got[ci]), want[ci], f"pipelined chunk {ci}")
        g.set_run_split(2)                               # resident, two sub-batches on two streams
        g.stage([c for c in chunks if True]); g.run(); got = g.fetch()
        for ci in range(len(chunks)):
            O.assert_blocks_equal(blockdict(got[ci]), want[ci], f"split chunk {ci}")
        g.set_run_split(1)
        got = g.bin_chunks(texts)                        # device-side parse
        for ci in range(len(chunks)):
            O.assert_blocks_equal(blockdict(got[ci]), want[ci], f"device-parse chunk {ci}", per_read=False)
        if not paired:
            t = keep[0]
            sig, info = g.find_new_minimizers(t[0], t[2], int(want[0]["read_signature"][0]), 4)
            ws, wi = O.new_minimizers_port(params, t[0], t[2], int(want[0]["read_signature"][0]), 4)
            assert np.array_equal(sig, ws) and np.array_equal(info, wi)
print("sanitizer case OK")
