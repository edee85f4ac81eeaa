"""North-star acceptance on one shard of BASELINE configs[2] (GPU box): the FASTQ files of a 12.5 M-pair shard (seed 1030 + shard)
on tmpfs -> reference `fastore_bin e -t1 -z -H -q0 -p8 -s0 -b256` and `fastore_bin_b200` with the same flags -> the four bin files
byte-compared (modulo the never-initialised padding of the parameter dump).

    python scripts/acceptance_shard.py [--shard 0] [--pairs 12500000] > gpurun_out/acceptance_shard0.json"""
import argparse
import hashlib
import json
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


def sha(path):
    h = hashlib.sha256()
    with open(path, "rb") as f:
        for blk in iter(lambda: f.read(1 << 24), b""):
            h.update(blk)
    return h.hexdigest()[:16]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--shard", type=int, default=0)
    ap.add_argument("--pairs", type=int, default=12_500_000)
    ap.add_argument("--gpus", type=int, default=0, help="0 = all")
    a = ap.parse_args()
    w = bench.WORKLOADS["c3"]
    base = Path("/dev/shm") if Path("/dev/shm").is_dir() else Path(tempfile.gettempdir())
    tmp = Path(tempfile.mkdtemp(prefix="fsb_acc_", dir=base))
    out = {"workload": w["desc"], "shard": a.shard, "seed": w["seed"] + a.shard, "pairs": a.pairs, "flags": "-z -H -q0 -p8 -s0 -b256"}
    try:
        t0 = time.time()
        files = [tmp / "s_1.fastq", tmp / "s_2.fastq"]
        with open(files[0], "wb") as f1, open(files[1], "wb") as f2:      # in pieces: the shard's text is ~8 GB
            step = 2_500_000
            for first in range(0, a.pairs, step):
                cfg = synth.synth_config(min(step, a.pairs - first), w["L"], paired=True, seed=w["seed"] + a.shard, first_index=first)
                t1, t2, _, _ = synth.generate(cfg, threads=min(32, bench.host_threads()), with_tables=False)
                f1.write(t1.tobytes()); f2.write(t2.tobytes())
        out["fastq_bytes"] = sum(f.stat().st_size for f in files)
        out["generate_s"] = round(time.time() - t0, 1)
        flags = dict(paired=True, b=256)
        args = BF.flags_to_args(flags)
        inp = "-i" + " ".join(str(f) for f in files)
        t0 = time.perf_counter()
        subprocess.run([str(BF.REF_DIR / "fastore_bin"), "e", inp, f"-o{tmp / 'ref'}", "-t1"] + args, check=True, capture_output=True)
        out["reference_t1_s"] = round(time.perf_counter() - t0, 2)
        t0 = time.perf_counter()
        r = subprocess.run([str(BF.CLI), "e", inp, f"-o{tmp / 'gpu'}", "-P16"] + ([f"-G{a.gpus}"] if a.gpus else []) + args, capture_output=True, text=True)
        out["fastore_bin_b200_s"] = round(time.perf_counter() - t0, 2)
        if r.returncode != 0:
            raise RuntimeError(r.stderr)
        BF.assert_bin_files_equal(tmp / "gpu", tmp / "ref", True)
        out["byte_identical"] = True
        out["files"] = {ext: {"bytes": Path(str(tmp / "gpu") + ext).stat().st_size, "sha256_16": sha(str(tmp / "gpu") + ext)} for ext in (".bdna", ".bqua", ".bhead", ".bmeta")}
        out["reads_per_s"] = {"reference_t1": 2 * a.pairs / out["reference_t1_s"], "fastore_bin_b200": 2 * a.pairs / out["fastore_bin_b200_s"]}
    finally:
        shutil.rmtree(tmp, ignore_errors=True)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
