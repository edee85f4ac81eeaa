"""Sub-batch overlap experiment (GPU box): one workload, many (run_split, K1 block batches, K4 block tiles) settings.

    python scripts/exp_overlap.py [pairs] > gpurun_out/overlap.txt
Prints ms per step of the resident path for each setting (10 steps after 3 warm-up steps, CUDA events)."""
import ctypes as C
import sys
sys.path.insert(0, '/root/repo')
import torch
import bench
from fastore_b200 import _native as N
from fastore_b200.binner import GpuBinner

n_pairs = int(sys.argv[1]) if len(sys.argv) > 1 else 10_000_000
w = bench.WORKLOADS["c2"]
lib = N.cuda_lib()
keep = []
chunks, ka = bench.workload_chunks(w, bench.rank_shards(w, 0, 1, n_pairs), pinned=True, lib=lib, keep=keep, threads=16)
params = bench.make_params(w)
stream = torch.cuda.Stream()
g = GpuBinner(params, device=0, stream=stream.cuda_stream)


def opt(o, v):
    g._check(lib.fsb_set_option(g._ctx, o, v))


def measure(split, r1, r4, always=0, steps=10):
    opt(5, split); opt(6, r1); opt(7, r4); opt(8, always)
    g.stage(chunks)
    for _ in range(3):
        g.run()
    g.sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            g.run()
        e1.record(stream)
    g.sync()
    return e0.elapsed_time(e1) / steps


print(f"{n_pairs} pairs, {len(chunks)} chunks")
print("split k1R k4R always  ms/step")
for split, r1, r4, always in [(1, 0, 0, 0), (1, 32, 64, 1), (1, 8, 16, 1), (1, 128, 256, 1),
                              (2, 32, 64, 0), (3, 32, 64, 0), (4, 32, 64, 0), (6, 32, 64, 0), (12, 32, 64, 0),
                              (4, 0, 0, 0), (4, 8, 16, 0), (4, 16, 32, 0), (4, 64, 128, 0), (4, 128, 256, 0), (4, 32, 16, 0), (4, 8, 64, 0),
                              (6, 16, 32, 0), (6, 64, 128, 0), (12, 16, 32, 0), (1, 0, 0, 0)]:
    ms = [measure(split, r1, r4, always) for _ in range(2)]
    print(f"{split:5d} {r1:4d} {r4:4d} {always:5d}   {ms[0]:.3f} {ms[1]:.3f}", flush=True)
g.close()
