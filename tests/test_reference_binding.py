"""The drop-in proof (GPU tier): oracle/_ref/fastore_bin_gpu is the reference's own fastore_bin -- same sources, built by
oracle/Makefile target ref_gpu -- whose Categorize + PackToBins pair (BinModule.cpp:130-133 / :379-382) goes through
integration/GpuBinEncoder.h, i.e. one fsb_bin_chunks call per chunk.  Reader, parser, statistics and bin-file writer are
the reference's.  Its files must equal the stock binary's byte for byte (modulo the never-initialised padding of the
parameter dump), and the reference's own decoder must read them back."""
import subprocess

import pytest

import binfile_helpers as BF

GPU_BIN = BF.REF_DIR / "fastore_bin_gpu"
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not GPU_BIN.exists(), reason="oracle/_ref/fastore_bin_gpu not built (make -C oracle ref_gpu; needs /root/reference)")]

CASES = [
    ("se100", dict(n=25000, L=100, paired=False, seed=401, nrich=0.03), dict(b=2)),
    ("pe150", dict(n=18000, L=150, paired=True, seed=402, nrich=0.03, lowcomplex=0.03), dict(paired=True, b=2)),
    ("pe100_reduced", dict(n=9000, L=100, paired=True, seed=403, header_comments=True), dict(paired=True, s=10, q=2, comments=False, b=2)),
    ("se250_max_noheads", dict(n=6000, L=250, paired=False, seed=404), dict(k=12, s=10, q=1, headers=False, b=2)),
]


@pytest.mark.parametrize("name,gen,flags", CASES, ids=[c[0] for c in CASES])
def test_reference_with_gpu_binding_writes_the_same_files(tmp_path, name, gen, flags):
    files = BF.write_fastq(tmp_path, name, gen["n"], gen["L"], gen["paired"], gen["seed"], **{k: v for k, v in gen.items() if k not in ("n", "L", "paired", "seed")})
    BF.run_reference_bin(files, tmp_path / "ref", flags)
    cmd = [str(GPU_BIN), "e", "-i" + " ".join(str(f) for f in files), f"-o{tmp_path / 'gpu'}", "-t1"] + BF.flags_to_args(flags)
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout + r.stderr
    BF.assert_bin_files_equal(tmp_path / "gpu", tmp_path / "ref", flags.get("headers", True))
    paired = bool(flags.get("paired"))
    outs = [tmp_path / "dec_1.fastq"] + ([tmp_path / "dec_2.fastq"] if paired else [])
    BF.decode_with_reference(tmp_path / "gpu", outs, paired)
    if flags.get("headers", True) and flags.get("comments", True) and flags.get("q", 0) == 0:
        for src, dec in zip(files, outs):
            assert sorted(BF.fastq_records(src)) == sorted(BF.fastq_records(dec))
