#!/usr/bin/env python3
"""Extract the hard-coded expected byte buffers of the reference's serialization regressions
(poly-commitment/tests/commitment.rs:288-443: ser_regression_canonical_{srs (Vesta, Pallas), polycomm, opening_proof}) into
tests/golden/ser_regression.json.  Run ONCE in the build container; nothing is computed here."""
import json
import os
import re

SRC = "/root/reference/poly-commitment/tests/commitment.rs"
OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "ser_regression.json")

text = open(SRC).read()
bufs = [[int(x) for x in re.findall(r"\d+", m)] for m in re.findall(r"let buf_expected: Vec<u8> = vec!\[(.*?)\];", text, flags=re.S)]
assert len(bufs) == 4, len(bufs)
names = ["srs_vesta_trusted_setup_depth8", "srs_pallas_trusted_setup_depth8", "polycomm_vesta_srs128_deg300_chunks6", "opening_proof_vesta_srs128"]
json.dump({"source": "poly-commitment/tests/commitment.rs:288-443 (seed [0u8; 32], commit 1494cf97 per the file's comments)",
           **{n: b for n, b in zip(names, bufs)}}, open(OUT, "w"))
print({n: len(b) for n, b in zip(names, bufs)})
