"""Sub-batch overlap experiment (GPU box): one workload, many (run_split, K1 block batches, K4 block tiles, fused layout) settings.

    python scripts/exp_overlap.py [pairs] > gpurun_out/overlap.txt
Prints ms per step of the resident path for each setting (10 steps after 3 warm-up steps, CUDA events), and for the
unsplit settings the per-stage times."""
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
g = GpuBinner(params, device=0, stream=stream.cuda_stream, profile=True)


def opt(o, v):
    g._check(lib.fsb_set_option(g._ctx, o, v))


def measure(split, r1, r4, always=0, fused=1, steps=10):
    opt(5, split); opt(6, r1); opt(7, r4); opt(8, always); opt(9, fused)
    g.stage(chunks)
    for _ in range(3):
        g.run()
    g.sync()
    g.stage_times()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with torch.cuda.stream(stream):
        e0.record(stream)
        for _ in range(steps):
            g.run()
        e1.record(stream)
    g.sync()
    st, runs = g.stage_times()
    return e0.elapsed_time(e1) / steps, {k: round(v / max(runs, 1), 3) for k, v in st.items()} if runs else {}


print(f"{n_pairs} pairs, {len(chunks)} chunks")
print("split k1R k4R always fused  ms/step")
for split, r1, r4, always, fused in [(1, 0, 0, 0, 1), (1, 0, 0, 0, 0), (2, 0, 0, 0, 1), (2, 32, 64, 0, 1), (4, 0, 0, 0, 1), (4, 32, 64, 0, 1), (4, 16, 32, 0, 1),
                                     (6, 0, 0, 0, 1), (6, 32, 64, 0, 1), (12, 0, 0, 0, 1), (12, 32, 64, 0, 1), (4, 0, 64, 0, 1), (4, 32, 0, 0, 1), (1, 0, 0, 0, 1)]:
    a = measure(split, r1, r4, always, fused)
    b = measure(split, r1, r4, always, fused)
    print(f"{split:5d} {r1:4d} {r4:4d} {always:5d} {fused:5d}   {a[0]:.3f} {b[0]:.3f}  {b[1]}", flush=True)
g.close()
