"""Mid-size GPU parity (-m gpu): a few hundred thousand records per case, so that the persistent kernels go round
their loops several times (K1: a warp walks over several warp batches with the next one prefetched; K4: a block
places several tiles with the next tile's slots in flight) for parameter sets other than the BASELINE one --
single-end, variable lengths, 250 bp with -p12 -s10, the reduced and max quality profiles, short reads.
Every case is compared with the C port bit for bit, as one chunk and as two chunks in one batch."""
import pytest

import oracle_helpers as O
from cases import CASES
from fastore_b200 import _native as N
from fastore_b200 import synth
from fastore_b200.binner import GpuBinner

pytestmark = pytest.mark.gpu

MID = {  # case of tests/cases.py -> records at this size
    "c1_se100_lossless": 300_000,
    "c2_pe150_fast_s10": 150_000,
    "c5_pe150_stress": 150_000,
    "se_varlen": 250_000,
    "pe_varlen": 150_000,
    "se_short_reads": 300_000,
    "pe_max_len_255_p4": 80_000,
}
MID.update({name: 120_000 for name, skw, pkw in CASES if name.startswith("c4_")})


def _block_dict(blk):
    return {"meta": blk.meta, "dna": blk.dna, "qua": blk.qua, "head": blk.head, "bins": blk.bins,
            "raw_dna_size": blk.raw_dna_size, "raw_head_size": blk.raw_head_size, "n_records": blk.n_records,
            "read_signature": blk.read_signature, "read_info": blk.read_info}


def _chunk(skw, params, n, first_index=0, seed_shift=0):
    kw = dict(skw)
    kw["n_records"] = n
    kw["seed"] = kw.get("seed", 1) + seed_shift
    kw["first_index"] = first_index
    cfg = synth.synth_config(**kw)
    t1, t2, r1, r2 = synth.generate(cfg, threads=8)
    if not params.reads_have_headers:
        r1 = r1.copy(); r1["head_len"] = 0
        if r2 is not None:
            r2 = r2.copy(); r2["head_len"] = 0
    return N.make_chunk(t1, r1, t2, r2), (t1, t2, r1, r2)


@pytest.mark.parametrize("name", sorted(MID))
def test_midsize_case_matches_oracle(name):
    skw, pkw = next((s, p) for n, s, p in CASES if n == name)
    params = N.make_params(**pkw)
    n = MID[name]
    a, keep_a = _chunk(skw, params, n)
    b, keep_b = _chunk(skw, params, n // 3 + 1, first_index=n, seed_shift=1000)
    with GpuBinner(params, per_read=True) as g:
        alone = g.bin_chunks([a])[0]
        both = g.bin_chunks([b, a])
    want_a = O.bin_chunk("orc", params, a)
    O.assert_blocks_equal(_block_dict(alone), want_a, f"{name}: one chunk")
    O.assert_blocks_equal(_block_dict(both[1]), want_a, f"{name}: second chunk of a batch")
    O.assert_blocks_equal(_block_dict(both[0]), O.bin_chunk("orc", params, b), f"{name}: first chunk of a batch")
