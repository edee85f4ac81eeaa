"""Multi-GPU plumbing: how chunks are spread over ranks and how results come back in chunk order.

The path shards by input chunk: chunks are independent units, each yields its own
BinaryBinBlock, and the bin-file format holds many blocks per bin.  There is no data-path collective;
torch.distributed is used for the barrier, for the max-over-ranks timing and -- when one rank writes
the bin file -- for collecting the per-chunk results in chunk order.  The same rule (chunk i ->
worker i mod G) is what the in-process dispatcher of fastore_bin_b200 uses (csrc/host/main.cpp)."""
from __future__ import annotations

import torch
import torch.distributed as dist


def chunks_of_rank(n_chunks: int, rank: int, world: int) -> list[int]:
    """Chunk indices binned by `rank`: round-robin, so every rank sees the whole file at the same pace."""
    return list(range(rank, n_chunks, world))


def max_over_ranks(x: float, device=None) -> float:
    """A multi-GPU time is the slowest rank's time."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(x: float, device=None) -> float:
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def gather_in_chunk_order(local: dict, n_chunks: int, dst: int = 0):
    """`local` maps chunk index -> picklable result of this rank.  Returns the list of all results in
    chunk order on rank `dst` (None elsewhere) -- the multi-process form of the ordered writer turn."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return [local[i] for i in range(n_chunks)]
    world, rank = dist.get_world_size(), dist.get_rank()
    for i in local:
        if i % world != rank:
            raise ValueError(f"chunk {i} does not belong to rank {rank}")
    out = [None] * world if rank == dst else None
    dist.gather_object(local, out, dst=dst)
    if rank != dst:
        return None
    merged = {}
    for part in out:
        merged.update(part)
    missing = [i for i in range(n_chunks) if i not in merged]
    if missing:
        raise RuntimeError(f"chunks {missing[:5]}... were binned by no rank")
    return [merged[i] for i in range(n_chunks)]
