"""Disk -> bin files wall clock (SURVEY.md 8d "third number"): the reference's fastore_bin e -t1 and
-t<cores> beside fastore_bin_b200, same FASTQ files on tmpfs, page cache warm, bin files of the GPU tool byte-compared
with the reference's -t1 output.  Run on the GPU box:

    python scripts/cli_wallclock.py [--pairs 2000000] > gpurun_out/cli_wallclock.json

configs[0] is run in full (1 M SE reads, -b16); configs[1] on its first --pairs pairs (-b256; the reference's -t1 needs
~90 s per 10 M pairs, so the default keeps the call short -- throughputs are per read and comparable)."""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
sys.path.insert(0, str(ROOT)); sys.path.insert(0, str(ROOT / "tests"))
import bench                      # noqa: E402
import binfile_helpers as BF      # noqa: E402
from fastore_b200 import synth    # noqa: E402


TRACE = False
SKIP_T1 = False


def timed(cmd):
    t0 = time.perf_counter()
    r = subprocess.run([str(c) for c in cmd], capture_output=True, text=True)
    dt = time.perf_counter() - t0
    if r.returncode != 0:
        raise RuntimeError(f"{cmd}: {r.stdout}{r.stderr}")
    return dt


def run_case(name, w, n, tmp, threads):
    cfg = synth.synth_config(n, w["L"], paired=w["paired"], seed=w["seed"], **w["synth"])
    t1, t2, _, _ = synth.generate(cfg, threads=min(32, threads), with_tables=False)
    files = [tmp / f"{name}_1.fastq"]
    files[0].write_bytes(t1.tobytes())
    if w["paired"]:
        files.append(tmp / f"{name}_2.fastq"); files[1].write_bytes(t2.tobytes())
    del t1, t2
    p = w["params"]
    flags = dict(paired=w["paired"], headers=bool(p.get("reads_have_headers", True)), comments=w.get("keep_comments", True), q=p.get("quality_method", 0),
                 k=p.get("signature_len", 8), s=p.get("skip_zone_len", 0), b=w["block_mib"])
    args = BF.flags_to_args(flags)
    inp = "-i" + " ".join(str(f) for f in files)
    mates = 2 if w["paired"] else 1
    out = {"config": name, "workload": w["desc"], "records": n, "reads": n * mates, "fastq_bytes": sum(f.stat().st_size for f in files), "flags": " ".join(args)}
    ref = BF.REF_DIR / "fastore_bin"
    if not SKIP_T1:
        t = timed([ref, "e", inp, f"-o{tmp / 'ref1'}", "-t1"] + args)
        out["reference_t1"] = {"seconds": t, "reads_per_s": n * mates / t, "threads": 1}
    tn = min(64, threads)
    t = timed([ref, "e", inp, f"-o{tmp / 'refN'}", f"-t{tn}"] + args)
    out["reference_tN"] = {"seconds": t, "reads_per_s": n * mates / t, "threads": tn}
    best = None
    for rep in range(2):                                        # second run: CUDA context creation and pinned allocation are what they are; take the better
        t = timed([BF.CLI, "e", inp, f"-o{tmp / 'gpu'}", "-P" + str(max(4, min(16, threads // 2)))] + args)
        best = t if best is None else min(best, t)
    if TRACE:
        w0 = time.time()
        r = subprocess.run([str(BF.CLI), "e", inp, f"-o{tmp / 'gpu_trace'}", "-v", "-P" + str(max(4, min(16, threads // 2)))] + args, capture_output=True, text=True)
        w1 = time.time()
        sys.stderr.write(f"---- {name}: fastore_bin_b200 -v ----\n[wall {w0:.3f}] spawn\n" + r.stderr.replace("\r", "\n") + f"\n[wall {w1:.3f}] reaped\n")
    out["fastore_bin_b200"] = {"seconds": best, "reads_per_s": n * mates / best, "gpus": "all", "parser_threads": max(4, min(16, threads // 2))}
    if not SKIP_T1:
        BF.assert_bin_files_equal(tmp / "gpu", tmp / "ref1", flags["headers"])
        out["byte_identical_to_reference_t1"] = True
        out["speedup_vs_t1"] = out["reference_t1"]["seconds"] / best
    out["speedup_vs_tN"] = out["reference_tN"]["seconds"] / best
    for f in tmp.iterdir():
        f.unlink()
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--pairs", type=int, default=2_000_000)
    ap.add_argument("--configs", default="c1,c2")
    ap.add_argument("--skip-t1", action="store_true", help="leave out the reference's -t1 run (and with it the byte comparison): for the full-size workload")
    ap.add_argument("--trace", action="store_true", help="one more run of the GPU tool with -v, its phase trace to stderr")
    a = ap.parse_args()
    global TRACE, SKIP_T1
    TRACE = a.trace
    SKIP_T1 = a.skip_t1
    threads = bench.host_threads()
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    tmp = Path(tempfile.mkdtemp(prefix="fsb_wall_", dir=base))
    res = {"host_threads": threads, "tmp": str(base), "cases": []}
    try:
        for name in a.configs.split(","):
            w = bench.WORKLOADS[name]
            n = w["n"] if name == "c1" else min(w["n"], a.pairs)
            res["cases"].append(run_case(name, w, n, tmp, threads))
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(res, indent=1))


if __name__ == "__main__":
    main()
