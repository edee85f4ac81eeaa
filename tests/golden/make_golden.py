#!/usr/bin/env python
"""Generates tests/golden/golden.json (+ small .npz with full outputs) from the COMPILED REFERENCE
(oracle/_ref/libfastore_ref.so: the reference's own Categorize + PackToBins objects).  Run in the
build container, where /root/reference exists:

    python tests/golden/make_golden.py

The reference ships no golden vectors of its own; these pin the oracle port and
the CUDA path on boxes where the reference cannot be compiled.  Inputs are regenerated from seeds by
the C++ generator; their digest is stored too, so a generator change is detected, not mis-attributed."""
import hashlib
import json
import sys
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
sys.path.insert(0, str(HERE.parent.parent))
sys.path.insert(0, str(HERE.parent))

import oracle_helpers as O  # noqa: E402
from cases import CASES, make_case  # noqa: E402

FULL = ("se_single_record", "pe_three_records")          # tiny cases stored byte for byte


def digest(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def block_digests(d: dict) -> dict:
    out = {s: {"size": int(d[s].size), "sha256": digest(d[s])} for s in ("meta", "dna", "qua", "head")}
    out["bins"] = {"count": int(d["bins"].shape[0]), "sha256": digest(d["bins"])}
    out["read_signature"] = digest(d["read_signature"])
    out["read_info"] = digest(d["read_info"])
    out["raw_dna_size"], out["raw_head_size"], out["n_records"] = int(d["raw_dna_size"]), int(d["raw_head_size"]), int(d["n_records"])
    return out


def main():
    assert O.have_reference(), "oracle/_ref is not built"
    golden = {}
    for name, _, _ in CASES:
        params, chunk, keep = make_case(name)
        d = O.bin_chunk("ref", params, chunk)
        t1, t2 = keep[0], keep[1]
        golden[name] = {"input_sha256": [digest(t1)] + ([digest(t2)] if t2 is not None else []), "block": block_digests(d)}
        if name in FULL:
            np.savez_compressed(HERE / f"{name}.npz", **{k: d[k] for k in ("meta", "dna", "qua", "head", "bins", "read_signature", "read_info")})
    (HERE / "golden.json").write_text(json.dumps(golden, indent=1, sort_keys=True) + "\n")
    print(f"wrote {len(golden)} cases")


if __name__ == "__main__":
    main()
