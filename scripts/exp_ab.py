"""A/B of run-time options on one workload in one process (GPU box): python scripts/exp_ab.py "name=opt:val,opt:val" ...
Prints ms per step and the per-stage times of an unsplit pass for every setting (10 steps after 3 warm-up steps, CUDA events), twice."""
import sys
sys.path.insert(0, '/root/repo')
import torch
import bench
from fastore_b200 import _native as N
from fastore_b200.binner import GpuBinner

w = bench.WORKLOADS["c2"]
lib = N.cuda_lib()
keep = []
chunks, ka = bench.workload_chunks(w, bench.rank_shards(w, 0, 1, 10_000_000), pinned=True, lib=lib, keep=keep, threads=16)
params = bench.make_params(w)
stream = torch.cuda.Stream()
g = GpuBinner(params, device=0, stream=stream.cuda_stream, profile=True)
DEFAULTS = {5: 1, 6: 32, 7: 64, 8: 0, 9: 1, 12: 1}


def measure(opts, steps=10):
    for o, v in {**DEFAULTS, **opts}.items():
        g._check(lib.fsb_set_option(g._ctx, o, v))
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


settings = []
for a in sys.argv[1:]:
    name, _, rest = a.partition("=")
    settings.append((name, {int(x.split(":")[0]): int(x.split(":")[1]) for x in rest.split(",") if x}))
for rep in range(2):
    for name, opts in settings:
        ms, st = measure(opts)
        print(f"{name:24s} {ms:.3f}  {st}", flush=True)
g.close()
