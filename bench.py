#!/usr/bin/env python
"""bench.py -- reads/s of the fastore_bin categorise + scatter path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            (N>1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one pass of the hot path (K1 ingest -> stable sort -> layout scans -> K4 place) over
one batch of synthetic FASTQ chunks.  Workload at N=1: BASELINE.json configs[1] -- 10 M synthetic
150 bp paired-end reads (pairs), lossless binning parameters (-z -H -q0 -p8 -s0), cut into chunks of
the size the reference's `-b256` chunk cutter produces.  At N>1 every rank bins its own 10 M-pair
shard (chunks are independent units: weak scaling, no collective on the data path).

  value   whole-job reads/s with the chunks already resident in HBM (kernels only, CUDA events on
          the library's stream, max over ranks);
  e2e     the same metric through the C ABI call fsb_bin_chunks with HOST buffers: pinned chunk
          text + record tables host->device, kernels, packed streams + descriptors device->host,
          every step;
  roofline  dominant kernel's algorithmic bytes / its CUDA-event duration vs MEASURED_PEAKS.json;
  cpu_baseline  the compiled reference's Categorize + PackToBins (oracle/_ref) timed on the host
          cores on a bounded sample of the same workload (rank 0, N=1 only).

The oracle libraries are loaded only for the cpu_baseline leg and for --impl reference.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time
from pathlib import Path

import numpy as np

ROOT = Path(__file__).resolve().parent
sys.path.insert(0, str(ROOT))

METRIC = "reads_per_sec_binned"
UNIT = "reads/s"
READ_LEN = 150
PAIRS_PER_GPU = 10_000_000
SEED = 102
CHUNK_MIB = 256            # -b256
PE_CUT_WINDOW_MIB = 1      # FastqStream.h:142: the PE chunk cutter backs off 1 MiB


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def load_peaks():
    p = ROOT / "MEASURED_PEAKS.json"
    if p.exists():
        try:
            return float(json.loads(p.read_text())["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """nvidia-smi sampling of SM clocks and throttle reasons during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.samples = []
        self._stop = threading.Event()
        self._th = None

    def _run(self):
        while not self._stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i", str(self.idx)],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    f = [x.strip() for x in out.split(",")]
                    self.samples.append((float(f[1]), float(f[2]), f[4], f[5], f[6], f[7]))
            except Exception:
                pass
            self._stop.wait(0.02)

    def start(self):
        self._th = threading.Thread(target=self._run, daemon=True)
        self._th.start()

    def stop(self) -> dict:
        self._stop.set()
        if self._th:
            self._th.join(timeout=10)
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(s[0] for s in self.samples)
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[2:]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.samples[0][1], "reasons": sorted(reasons), "samples": len(sm)}


def pinned_alloc_factory(lib, keep):
    """numpy uint8 arrays over cudaHostAlloc'd memory (freed at exit)."""
    def alloc(nbytes):
        p = lib.fsb_host_alloc(max(int(nbytes), 1))
        if not p:
            raise MemoryError("fsb_host_alloc failed")
        keep.append(p)
        buf = (C.c_uint8 * int(nbytes)).from_address(p)
        return np.frombuffer(buf, dtype=np.uint8)
    return alloc


def pinned_records(lib, keep, recs):
    from fastore_b200 import _native as N
    raw = pinned_alloc_factory(lib, keep)(recs.nbytes)
    out = raw.view(N.RECORD_DTYPE)
    out[:] = recs
    return out


def workload_chunks(rank: int, n_pairs: int, pinned: bool, lib=None, keep=None, threads=None):
    """The rank's shard as a list of chunks cut like `-b256` cuts PE input (FastqStream.cpp:104-228):
    every chunk but the last holds the records that fit in (256 - 1) MiB of mate-1 text."""
    from fastore_b200 import _native as N
    from fastore_b200 import synth
    first = rank * n_pairs
    probe = synth.synth_config(1, READ_LEN, paired=True, seed=SEED, first_index=first + n_pairs - 1)
    b1, b2 = C.c_uint64(), C.c_uint64()
    N.host_lib().fsh_synth_size(C.byref(probe), C.byref(b1), C.byref(b2))
    per_chunk = ((CHUNK_MIB - PE_CUT_WINDOW_MIB) << 20) // int(b1.value)      # widest header of the shard
    chunks, keepalive, done = [], [], 0
    alloc = pinned_alloc_factory(lib, keep) if pinned else None
    while done < n_pairs:
        n = min(per_chunk, n_pairs - done)
        cfg = synth.synth_config(n, READ_LEN, paired=True, seed=SEED, first_index=first + done)
        t1, t2, r1, r2 = synth.generate(cfg, threads=threads, out=alloc)
        if pinned:
            r1, r2 = pinned_records(lib, keep, r1), pinned_records(lib, keep, r2)
        keepalive.append((t1, t2, r1, r2))
        chunks.append(N.make_chunk(t1, r1, t2, r2))
        done += n
    return chunks, keepalive


def lossless_pe_params():
    from fastore_b200 import _native as N
    return N.make_params(signature_len=8, skip_zone_len=0, paired_end=True, quality_method=N.FSB_QUA_NONE,
                         quality_offset=33, reads_have_headers=True)


def host_threads() -> int:
    try:
        return max(1, len(os.sched_getaffinity(0)))
    except Exception:
        return max(1, os.cpu_count() or 1)


def config_dict(n_gpus, n_pairs, n_chunks):
    return {"workload": "BASELINE configs[1]: 10M synthetic 150 bp paired-end reads (pairs), lossless binning params "
                        "(-z -H -q0 -p8 -s0), -b256 chunking, per GPU",
            "pairs_per_gpu": n_pairs, "reads_per_gpu": 2 * n_pairs, "read_len": READ_LEN, "chunks_per_gpu": n_chunks,
            "signature_len": 8, "skip_zone_len": 0, "quality_mode": "q0 (6 bit)", "headers": True,
            "sharding": f"{n_gpus} x independent chunk shards, no data-path collective",
            "l2_policy": "inputs (6.9 GB/GPU) and outputs (3.2 GB/GPU) far exceed the 126 MB L2; no flush needed"}


# -------------------------------------------------------------------------------------------------
def cpu_reference_leg(params, chunks, threads, target_s=12.0):
    """Time the compiled reference (oracle/_ref) -- or the C port when it is absent -- on a bounded
    sample: the first `n` pairs of chunk 0, split across `threads` workers like `-t N`."""
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_helpers as O
    from fastore_b200 import _native as N
    kind = "ref" if O.have_reference() else "orc"
    ch0 = chunks[0]
    # ~60 k pairs/s/core is the survey's figure for PE-150; size the sample for ~target_s
    n = int(min(ch0.n_records, max(20000, 50_000 * threads * target_s / 12.0 * 3)))
    sample = N.FsbChunk()
    C.memmove(C.byref(sample), C.byref(ch0), C.sizeof(N.FsbChunk))
    sample.n_records = n
    return kind, sample, n


def run_cpu(kind, params, sample, threads, reps=1):
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_helpers as O
    return O.time_bin_chunk(kind, params, sample, threads, reps)


def run_reference_arm(args):
    """--impl reference: the reference's own CPU implementation of the path on the host cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return 0
    from fastore_b200 import build
    build.build_host()
    sys.path.insert(0, str(ROOT / "tests"))
    import oracle_helpers as O
    if not O.PORT_LIB.exists() and not O.have_reference():
        build.build_oracle()
    params = lossless_pe_params()
    threads = host_threads()
    # bounded sample of the workload: ~1-3 s of work per step on all cores
    n_sample = int(min(PAIRS_PER_GPU, max(50_000, 150_000 * threads)))
    chunks, keep = workload_chunks(0, n_sample, pinned=False, threads=min(32, threads))
    kind, sample, n = cpu_reference_leg(params, chunks, threads)
    sample.n_records = min(n_sample, chunks[0].n_records)
    n = int(sample.n_records)
    # calibrate: one step = `reps` passes over the sample, about 2 s of wall time on all cores
    once = run_cpu(kind, params, sample, threads)
    reps = int(max(1, min(64, round(2.0 / max(once, 1e-3)))))
    for _ in range(args.warmup):
        run_cpu(kind, params, sample, threads, reps)
    t = 0.0
    for _ in range(args.steps):
        t += run_cpu(kind, params, sample, threads, reps)
    ms = 1e3 * t / max(args.steps, 1)
    value = 2 * n * reps / (ms / 1e3)
    n = n * reps
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic", "config": config_dict(args.gpus, PAIRS_PER_GPU, 0),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads,
                             "kind": "reference" if kind == "ref" else "port",
                             "sample": f"{n} pairs per step (first records of the rank-0 shard, repeated), Categorize+PackToBins on {threads} threads, "
                                       f"one slice per thread like fastore_bin -t{threads}"},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)
    return 0


# -------------------------------------------------------------------------------------------------
def run_b200_arm(args):
    import torch
    import torch.distributed as dist

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device -- this path has no CPU fallback (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("NCCL_DEBUG_FILE", "/dev/stderr")      # stdout carries exactly one JSON line
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from fastore_b200 import _native as N
    from fastore_b200 import build
    from fastore_b200.binner import GpuBinner
    build.build_host()
    if not build.CUDA_LIB.exists():
        build.build_cuda()
    lib = N.cuda_lib()
    params = lossless_pe_params()
    n_pairs = args.pairs
    keep_ptrs = []
    t0 = time.time()
    gen_threads = max(1, min(32, host_threads() // max(1, world)))
    chunks, keep = workload_chunks(rank, n_pairs, pinned=True, lib=lib, keep=keep_ptrs, threads=gen_threads)
    log(f"[rank {rank}] generated {n_pairs} pairs in {len(chunks)} chunks in {time.time() - t0:.1f}s")

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    from fastore_b200 import sharding

    def max_over_ranks(x: float) -> float:
        return sharding.max_over_ranks(x, device="cuda")

    stream = torch.cuda.Stream()
    g = GpuBinner(params, device=local, stream=stream.cuda_stream, profile=True)

    # ---- resident (kernel) timing ----------------------------------------------------------------
    g.stage(chunks)
    g.sync()
    for _ in range(args.warmup):
        g.run()
    g.sync()
    g.stage_times()                                   # reset the per-stage accumulators
    launches0 = g.stats()["kernel_launches"]
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    with torch.cuda.stream(stream):
        ev0.record(stream)
        for _ in range(args.steps):
            g.run()
        ev1.record(stream)
    g.sync()
    barrier()
    ms_total = max_over_ranks(ev0.elapsed_time(ev1))
    launches = g.stats()["kernel_launches"] - launches0
    stage_ms, runs = g.stage_times()
    blocks = g.fetch()
    alg_bytes = g.stats()["algorithmic_bytes"]        # input bytes consumed + output bytes produced, one pass
    out_bytes = sum(int(b.meta.size + b.dna.size + b.qua.size + b.head.size) for b in blocks)
    n_bins = sum(int(b.bins.shape[0]) for b in blocks)
    del blocks
    ms_per_step = ms_total / args.steps
    value = world * 2 * n_pairs / (ms_per_step / 1e3)

    # ---- end to end through the C ABI with host buffers -----------------------------------------------
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    blocks_c = (N.FsbBlock * len(chunks))()
    g._check(lib.fsb_bin_chunks(g._ctx, g._chunk_array(chunks), len(chunks), blocks_c))   # untimed pass: pinned result buffers get allocated
    st0 = g.stats()
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        arr = g._chunk_array(chunks)
        blocks_c = (N.FsbBlock * len(chunks))()
        g._check(lib.fsb_bin_chunks(g._ctx, arr, len(chunks), blocks_c))
    torch.cuda.synchronize()
    e2e_s = max_over_ranks(time.perf_counter() - t0)
    barrier()
    clocks = sampler.stop() if rank == 0 else None       # sampled through both timed regions (resident steps and end-to-end steps)
    st1 = g.stats()
    e2e_value = world * 2 * n_pairs * e2e_steps / e2e_s
    h2d = (st1["h2d_bytes"] - st0["h2d_bytes"]) // e2e_steps
    d2h = (st1["d2h_bytes"] - st0["d2h_bytes"]) // e2e_steps

    # ---- roofline of the dominant stage ---------------------------------------------------------------
    peak, peak_src = load_peaks()
    dom = max(stage_ms, key=lambda k: stage_ms[k])
    dom_ms = stage_ms[dom] / max(runs, 1)
    traffic = None
    try:                                               # dram__bytes_read + write per pair of the dominant kernel, from the committed ncu capture
        tj = json.loads((ROOT / "profiles" / "traffic.json").read_text())
        if dom in tj["bytes_per_pair"]:
            traffic = tj["bytes_per_pair"][dom] * n_pairs
    except Exception:
        pass
    roof = {"bound": "hbm", "kernel": dom, "achieved": alg_bytes / (dom_ms / 1e3) / 1e9, "peak": peak, "unit": "GB/s",
            "frac": alg_bytes / (dom_ms / 1e3) / 1e9 / peak, "traffic": traffic, "peak_source": peak_src,
            "algorithmic_bytes_per_launch": alg_bytes,
            "whole_path_achieved": alg_bytes / (ms_per_step / 1e3) / 1e9,
            "whole_path_frac": alg_bytes / (ms_per_step / 1e3) / 1e9 / peak,
            "stage_ms": {k: v / max(runs, 1) for k, v in stage_ms.items()}}

    line = None
    if rank == 0:
        line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
                "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8",
                "data": "synthetic", "config": config_dict(world, n_pairs, len(chunks)),
                "pairs_per_sec": value / 2, "bins_per_step_per_gpu": n_bins, "output_bytes_per_gpu": out_bytes,
                "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
                        "steps": e2e_steps, "ms_per_step": 1e3 * e2e_s / e2e_steps},
                "gpu_launches": int(launches), "roofline": roof, "clocks": clocks}
        if world == 1 and not args.no_cpu:
            threads = host_threads()
            kind, sample, n = cpu_reference_leg(params, chunks, threads)
            once = run_cpu(kind, params, sample, threads)                     # calibration pass
            reps = int(max(1, min(200, round(12.0 / max(once, 1e-3)))))       # about 12 s of work on all cores
            secs = run_cpu(kind, params, sample, threads, reps)
            line["cpu_baseline"] = {"value": 2 * n * reps / secs, "unit": UNIT, "cores": threads,
                                    "kind": "reference" if kind == "ref" else "port",
                                    "sample": f"{n} pairs of chunk 0 x {reps} passes, Categorize+PackToBins, {threads} threads "
                                              f"(one slice per thread like fastore_bin -t{threads}), {secs:.1f} s"}
    g.close()
    for p in keep_ptrs:
        lib.fsb_host_free(p)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line), flush=True)
    return 0


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs", type=int, default=PAIRS_PER_GPU, help="pairs per GPU (default: the BASELINE workload)")
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("note: the timing rules ask for >= 3 warm-up steps")
    if args.impl == "reference":
        return run_reference_arm(args)
    return run_b200_arm(args)


if __name__ == "__main__":
    sys.exit(main())
