"""BASELINE.json configs[1] at full size on the GPU (-m gpu): 10 M synthetic 150 bp pairs, lossless binning
parameters, 12 chunks in one batch -- the workload bench.py times.

The whole batch is checked through size-independent properties (every record lands in exactly one bin,
descriptors add up to the streams, bins ascend with the N-bin last, raw sizes equal the input's), and then
EVERY chunk of the batch -- text offsets beyond 2^31, record indices in the millions -- is compared bit for
bit (streams, descriptors, per-read signature / position / flags) with the compiled reference's own
Categorize + PackToBins (oracle/_ref/libfastore_ref.so; the C port where the reference is not built), one
chunk per host thread.  That also shows that a chunk's block does not depend on the batch it was binned in."""
from concurrent.futures import ThreadPoolExecutor

import numpy as np
import pytest

import oracle_helpers as O
from fastore_b200 import _native as N
from fastore_b200.binner import GpuBinner

pytestmark = pytest.mark.gpu

N_PAIRS = 10_000_000


def test_baseline_config1_full_size():
    import bench
    w = bench.WORKLOADS["c2"]
    READ_LEN = w["L"]
    params = bench.make_params(w)
    chunks, keep = bench.workload_chunks(w, bench.rank_shards(w, 0, 1, N_PAIRS), pinned=False, threads=min(16, bench.host_threads()))
    assert len(chunks) >= 10 and sum(int(c.n_records) for c in chunks) == N_PAIRS
    with GpuBinner(params, per_read=True) as g:
        blocks = g.bin_chunks(chunks)
    nbin = 1 << 16
    for ci, (ch, blk) in enumerate(zip(chunks, blocks)):
        n = int(ch.n_records)
        bins = blk.bins
        assert blk.n_records == n
        assert int(bins["records_count"].sum()) == n, f"chunk {ci}: records in bins"
        sig = bins["signature"].astype(np.int64)
        assert np.all(np.diff(sig) > 0), f"chunk {ci}: bins not strictly ascending"
        assert sig[-1] <= nbin and (sig[:-1] < nbin).all()
        for field, stream in (("meta_size", blk.meta), ("dna_size", blk.dna), ("qua_size", blk.qua), ("head_size", blk.head)):
            assert int(bins[field].sum()) == stream.size, f"chunk {ci}: {field} does not add up to the stream"
        assert int(bins["raw_dna_size"].sum()) == blk.raw_dna_size == 2 * READ_LEN * n
        assert int(bins["raw_head_size"].sum()) == blk.raw_head_size
        # per-read results: every signature is a bin of the chunk, and the per-bin counts agree
        rs = blk.read_signature
        assert rs.shape[0] == n
        counts = np.bincount(rs, minlength=nbin + 1)
        assert np.array_equal(np.nonzero(counts)[0], sig), f"chunk {ci}: bins vs per-read signatures"
        assert np.array_equal(counts[sig], bins["records_count"].astype(np.int64)), f"chunk {ci}: per-bin record counts"
        # lossless 6-bit quality: 2 * 150 * 6 bits per pair, byte padding per bin only
        assert blk.qua.size >= (n * 2 * READ_LEN * 6) // 8 and blk.qua.size <= (n * 2 * READ_LEN * 6) // 8 + bins.shape[0]
    # every chunk against the compiled reference, bit for bit (ctypes releases the GIL: one chunk per host thread)
    kind = "ref" if O.have_reference() else "orc"

    def check(ci):
        blk = blocks[ci]
        got = {"meta": blk.meta, "dna": blk.dna, "qua": blk.qua, "head": blk.head, "bins": blk.bins,
               "raw_dna_size": blk.raw_dna_size, "raw_head_size": blk.raw_head_size, "n_records": blk.n_records,
               "read_signature": blk.read_signature, "read_info": blk.read_info}
        O.assert_blocks_equal(got, O.bin_chunk(kind, params, chunks[ci]), f"chunk {ci} of the 10 M-pair batch vs '{kind}'")
        return ci

    with ThreadPoolExecutor(max_workers=max(1, min(len(chunks), bench.host_threads()))) as ex:
        assert sorted(ex.map(check, range(len(chunks)))) == list(range(len(chunks)))
