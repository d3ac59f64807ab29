#!/usr/bin/env python3
"""Extract the Poseidon parameters of the reference's Fiat-Shamir sponges (poseidon/src/pasta/{fp,fq}_kimchi.rs: MDS matrix and
55 x 3 round constants, decimal strings) and its own hash test vectors (poseidon/tests/test_vectors/kimchi.json, Fp) into
tests/golden/poseidon_kimchi.json.  Run ONCE in the build container; nothing is computed here."""
import json
import os
import re

REF = "/root/reference/poseidon"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "poseidon_kimchi.json")


def params(path):
    text = open(path).read()
    nums = [int(x) for x in re.findall(r'"(\d{20,})"', text)]
    assert len(nums) == 9 + 55 * 3, len(nums)
    mds_pos, rc_pos = text.index("mds:"), text.index("round_constants:")
    assert mds_pos < rc_pos
    return {"mds": [nums[3 * i: 3 * i + 3] for i in range(3)], "round_constants": [nums[9 + 3 * r: 12 + 3 * r] for r in range(55)]}


out = {"source": "poseidon/src/pasta/fp_kimchi.rs, fq_kimchi.rs; poseidon/tests/test_vectors/kimchi.json",
       "fp": params(f"{REF}/src/pasta/fp_kimchi.rs"), "fq": params(f"{REF}/src/pasta/fq_kimchi.rs"),
       "fp_hash_vectors": json.load(open(f"{REF}/tests/test_vectors/kimchi.json"))["test_vectors"]}
json.dump(out, open(OUT, "w"))
print(len(out["fp_hash_vectors"]), "vectors;", os.path.getsize(OUT), "bytes")
