"""world_size-2 check of the N>1 host logic on CPU (gloo): chunk -> rank assignment, ordered gather of
per-chunk results, max-over-ranks timing.  The per-chunk work is done by the oracle here (CPU tier);
on GPUs bench.py / the CLI put the CUDA path in its place -- the sharding code is the same."""
import hashlib
import os
import socket

import numpy as np
import pytest
import torch.distributed as dist
import torch.multiprocessing as mp

import oracle_helpers as O
from fastore_b200 import _native as N
from fastore_b200 import sharding, synth

N_CHUNKS = 5


def chunk_digest(ci: int) -> str:
    cfg = synth.synth_config(1500 + 100 * ci, 100, paired=True, seed=300, first_index=ci * 10000, nrich=0.05)
    t1, t2, r1, r2 = synth.generate(cfg, threads=1)
    params = N.make_params(signature_len=8, skip_zone_len=0, paired_end=True)
    d = O.bin_chunk("orc", params, N.make_chunk(t1, r1, t2, r2))
    h = hashlib.sha256()
    for s in ("meta", "dna", "qua", "head"):
        h.update(np.ascontiguousarray(d[s]).tobytes())
    h.update(d["bins"].tobytes())
    return h.hexdigest()


def worker(rank: int, world: int, port: int, q):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        mine = sharding.chunks_of_rank(N_CHUNKS, rank, world)
        local = {ci: chunk_digest(ci) for ci in mine}
        ordered = sharding.gather_in_chunk_order(local, N_CHUNKS, dst=0)
        slowest = sharding.max_over_ranks(1.0 + rank)
        total = sharding.sum_over_ranks(float(len(mine)))
        if rank == 0:
            q.put((ordered, slowest, total))
        with pytest.raises(ValueError):
            sharding.gather_in_chunk_order({(rank + 1) % world: "x"}, N_CHUNKS)      # a chunk of another rank
        dist.barrier()
    finally:
        dist.destroy_process_group()


def test_two_ranks_cover_all_chunks_in_order():
    world = 2
    assert sorted(sum((sharding.chunks_of_rank(N_CHUNKS, r, world) for r in range(world)), [])) == list(range(N_CHUNKS))
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs: p.start()
    ordered, slowest, total = q.get(timeout=300)
    for p in procs:
        p.join(timeout=300)
        assert p.exitcode == 0
    assert ordered == [chunk_digest(ci) for ci in range(N_CHUNKS)]       # same as one process, chunk order kept
    assert slowest == 2.0 and total == N_CHUNKS
