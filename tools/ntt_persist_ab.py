#!/usr/bin/env python3
"""A/B of the NTT launch geometry: one CTA per tile (1024 tiles of a 2^20 transform on 888 resident CTA slots = 1.15 waves) against
the balanced persistent grid (every CTA transforms the same number of columns).  Device-resident Fp data, L2 flushed, median of 15."""
import json, os, sys
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proof_systems_b200 as zk
from bench import splitmix64_limbs
ctx = zk.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
def timed(fn, reps=15):
    ts = []
    for _ in range(reps + 3):
        flush.fill_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[3:]))
rows = []
for log_n, batch in ((20, 1), (16, 1), (18, 1), (19, 16), (16, 16), (21, 1)):
    d = torch.from_numpy(splitmix64_limbs(2, batch << log_n).view(np.int64)).cuda()
    ref = None
    for mode in (0, 1, 0, 1):
        ctx.set_option("ntt_persistent", mode)
        t = timed(lambda: ctx.ntt_dev(zk.FP, d.data_ptr(), log_n, batch=batch))
        rows.append({"log_n": log_n, "batch": batch, "persistent": mode, "ms": round(t, 4)})
        print(rows[-1], flush=True)
ctx.set_option("ntt_persistent", 1)
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "ntt_persist_ab.json"), "w"), indent=1)
