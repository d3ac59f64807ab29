#!/usr/bin/env python3
"""Throughput of a batch of independent 2^16-point MSMs on one basis (zk_msm_batch: the 15 witness columns / 7 t-chunks of a
kimchi proof) as a function of the number of concurrent lanes; scalars in pinned host memory (read over PCIe by the recode
kernel), results on the host.  Wall clock around the public call, median of 5."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proof_systems_b200 as zk
from bench import splitmix64_limbs
from oracle import oracle as orc
ctx = zk.Context(0)
z = np.load(os.path.join(ROOT, "tests", "golden", "pallas_srs.npz"))
g = ctx.decompress_points(zk.PALLAS, z["g_cmp"])
n, k = 1 << 16, 16
bases = ctx.upload_bases(zk.PALLAS, g)
pin = torch.empty((k, n, 4), dtype=torch.int64).pin_memory()
sc = pin.numpy().view(np.uint64)
for kind in ("uniform", "witness"):
    for j in range(k):
        sc[j] = splitmix64_limbs(100 + j, n)
        if kind == "witness":
            sc[j, : n - 3] = 0; sc[j, : n - 10, 0] = 1
    want = orc.msm(orc.PALLAS, g, sc[3], threads=orc.host_threads())
    rows = {}
    for lanes in (1, 2, 3, 4):
        ctx.set_option("msm_lanes", lanes)
        ts = []
        for _ in range(6):
            t = time.perf_counter(); out = ctx.msm_batch(bases, sc); ts.append(time.perf_counter() - t)
        ok = bool(np.array_equal(zk.jacobian_to_affine(zk.PALLAS, out[3]), want))
        rows[lanes] = {"ms_per_msm": round(1e3 * float(np.median(ts[1:])) / k, 4), "bit_exact": ok}
    print(kind, json.dumps(rows), flush=True)
