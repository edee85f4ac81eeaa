"""Pins the C port (oracle/fastore_oracle.c) against the compiled reference (oracle/_ref).

The reference has no golden vectors of its own, so the port is trusted only
because it reproduces the reference's streams, descriptors and per-read tuples byte for byte here.
On a box without oracle/_ref the same check runs against tests/golden/ (test_golden.py)."""
import numpy as np
import pytest

import oracle_helpers as O
from cases import CASES, make_case
from fastore_b200 import _native as N

pytestmark = pytest.mark.skipif(not O.have_reference(), reason="oracle/_ref not built (no /root/reference here)")


@pytest.mark.parametrize("name", [c[0] for c in CASES])
def test_port_matches_compiled_reference(name):
    params, chunk, keep = make_case(name)
    a = O.bin_chunk("orc", params, chunk)
    b = O.bin_chunk("ref", params, chunk)
    O.assert_blocks_equal(a, b, name)
    assert a["n_records"] == chunk.n_records
    assert int(a["bins"]["records_count"].sum()) == chunk.n_records


def test_find_minimizer_random_strings():
    rng = np.random.default_rng(5)
    for k, s in ((8, 0), (8, 10), (12, 10), (4, 0), (13, 3)):
        params = N.make_params(signature_len=k, skip_zone_len=s)
        for _ in range(300):
            L = int(rng.integers(1, 256))
            alphabet = np.frombuffer(b"ACGTN" if rng.random() < 0.5 else b"ACGT", dtype=np.uint8)
            if rng.random() < 0.3:
                alphabet = np.frombuffer(b"AAAC", dtype=np.uint8)
            seq = alphabet[rng.integers(0, alphabet.size, L)].tobytes()
            assert O.find_minimizer("orc", params, seq) == O.find_minimizer("ref", params, seq), (k, s, seq)
