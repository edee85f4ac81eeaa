import sys
sys.path.insert(0,'/root/repo'); sys.path.insert(0,'/root/repo/tests')
import oracle_helpers as O
from fastore_b200 import _native as N
from fastore_b200 import synth
from fastore_b200.binner import GpuBinner
for paired in (True, False):
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=paired)
    keep, chunks = [], []
    for ci,(n,L) in enumerate([(3000,150),(1,100),(2500,151),(700,36)]):
        cfg = synth.synth_config(n, L, paired=paired, seed=900+ci, first_index=ci*100000, nrich=0.05, lowcomplex=0.05, alln=0.01, tie=0.02)
        t = synth.generate(cfg, threads=2); keep.append(t); chunks.append(N.make_chunk(t[0], t[2], t[1], t[3]))
    with GpuBinner(params, per_read=True) as g:
        got = g.bin_chunks(chunks)
    for ci,ch in enumerate(chunks):
        want = O.bin_chunk("orc", params, ch); blk = got[ci]
        d = {"meta": blk.meta, "dna": blk.dna, "qua": blk.qua, "head": blk.head, "bins": blk.bins, "raw_dna_size": blk.raw_dna_size, "raw_head_size": blk.raw_head_size, "n_records": blk.n_records, "read_signature": blk.read_signature, "read_info": blk.read_info}
        O.assert_blocks_equal(d, want, f"chunk {ci}")
print("sanitizer case OK")
