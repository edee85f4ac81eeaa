"""fastore_rebin's signature scan: DnaRebalancer::FindNewMinimizer.  The C port is pinned to the reference's own
member function (oracle/_ref/libfastore_ref_rebin.so), the kernel's per-thread routine runs on the host against the port (CPU tier),
and the kernel itself through the C ABI (GPU tier)."""
import ctypes as C

import numpy as np
import pytest

import oracle_helpers as O
from fastore_b200 import _native as N
from fastore_b200 import synth

SETS = [dict(k=8, s=0, parity=2), dict(k=8, s=0, parity=8), dict(k=8, s=10, parity=4), dict(k=10, s=4, parity=16), dict(k=6, s=2, parity=2), dict(k=12, s=10, parity=4)]


def reads(n, L, seed, **kw):
    cfg = synth.synth_config(n, L, seed=seed, nrich=0.1, lowcomplex=0.1, alln=0.01, tie=0.05, **kw)
    t1, _, r1, _ = synth.generate(cfg, threads=2)
    return t1, r1


def a_signature_of(params, text, recs, i):
    """a signature that really occurs: the plain minimizer of read i"""
    raw = text.tobytes()
    return O.find_minimizer("orc", params, raw[int(recs["seq_off"][i]): int(recs["seq_off"][i]) + int(recs["seq_len"][i])])[0]


@pytest.mark.skipif(not O.REBIN_REF_LIB.exists(), reason="oracle/_ref not built (no /root/reference here)")
@pytest.mark.parametrize("ps", SETS, ids=[f"k{p['k']}s{p['s']}p{p['parity']}" for p in SETS])
def test_port_equals_the_reference_member_function(ps):
    params = N.make_params(signature_len=ps["k"], skip_zone_len=ps["s"])
    text, recs = reads(600, 100, 70 + ps["k"])
    raw = text.tobytes()
    curs = [a_signature_of(params, text, recs, i) for i in (0, 7, 99)] + [0, 5]
    for cur in curs:
        for r in recs[::3]:
            seq = raw[int(r["seq_off"]): int(r["seq_off"]) + int(r["seq_len"])]
            assert O.find_new_minimizer("orc", params, seq, cur, ps["parity"]) == O.find_new_minimizer("ref", params, seq, cur, ps["parity"]), (cur, seq)


@pytest.mark.parametrize("ps", SETS, ids=[f"k{p['k']}s{p['s']}p{p['parity']}" for p in SETS])
def test_kernel_core_on_the_host_equals_the_port(ps):
    from test_kernel_core_on_cpu import EMUL_LIB, EMUL_SRC, ROOT
    import subprocess
    deps = [EMUL_SRC] + list((ROOT / "fastore_b200" / "csrc").glob("*.cuh")) + [ROOT / "include" / "fastore_b200.h"]
    if not EMUL_LIB.exists() or any(d.stat().st_mtime > EMUL_LIB.stat().st_mtime for d in deps):
        EMUL_LIB.parent.mkdir(exist_ok=True)
        subprocess.run(["g++", "-std=c++17", "-O2", "-fPIC", "-shared", "-Wall", "-x", "c++", str(EMUL_SRC), "-o", str(EMUL_LIB)], check=True)
    lib = C.CDLL(str(EMUL_LIB))
    lib.emul_new_minimizers.restype = C.c_int
    lib.emul_new_minimizers.argtypes = [C.POINTER(N.FsbParams), C.c_void_p, C.c_uint64, C.c_void_p, C.c_uint64, C.c_uint32, C.c_uint32, C.c_void_p, C.c_void_p]
    params = N.make_params(signature_len=ps["k"], skip_zone_len=ps["s"])
    for L, seed in ((100, 81), (151, 82), (36, 83), (250, 84)):
        text, recs = reads(800, L, seed + ps["k"], min_len=max(1, L // 3))
        for cur in (a_signature_of(params, text, recs, 3), 0):
            want = O.new_minimizers_port(params, text, recs, cur, ps["parity"])
            sig = np.zeros(len(recs), dtype=np.uint32); info = np.zeros(len(recs), dtype=np.uint32)
            assert lib.emul_new_minimizers(C.byref(params), N.np_ptr(text), text.size, N.np_ptr(recs), len(recs), cur, ps["parity"], N.np_ptr(sig), N.np_ptr(info)) == 0
            assert np.array_equal(sig, want[0]), f"L {L} cur {cur}: signature of read {int(np.nonzero(sig != want[0])[0][0])}"
            assert np.array_equal(info, want[1]), f"L {L} cur {cur}: info of read {int(np.nonzero(info != want[1])[0][0])}"


@pytest.mark.gpu
@pytest.mark.parametrize("ps", SETS, ids=[f"k{p['k']}s{p['s']}p{p['parity']}" for p in SETS])
def test_kernel_equals_the_port(ps):
    from fastore_b200.binner import GpuBinner, FastoreError
    params = N.make_params(signature_len=ps["k"], skip_zone_len=ps["s"])
    with GpuBinner(params) as g:
        for L, seed in ((100, 91), (151, 92), (36, 93), (255, 94)):
            text, recs = reads(3000, L, seed + ps["k"], min_len=max(1, L // 3))
            for cur in (a_signature_of(params, text, recs, 3), 0):
                want = O.new_minimizers_port(params, text, recs, cur, ps["parity"])
                sig, info = g.find_new_minimizers(text, recs, cur, ps["parity"])
                assert np.array_equal(sig, want[0]), f"L {L} cur {cur}: signature of read {int(np.nonzero(sig != want[0])[0][0])}"
                assert np.array_equal(info, want[1]), f"L {L} cur {cur}: info of read {int(np.nonzero(info != want[1])[0][0])}"
        with pytest.raises(FastoreError):
            g.find_new_minimizers(text, recs, 0, 3)          # the divisor is a power of two
