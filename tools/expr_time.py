#!/usr/bin/env python3
"""zk_expr_eval_dev at the prover's size: the poseidon gate (15 constraints, 60 constants, 15 cached S-box powers) over d8 of a 2^16-row
circuit (2^19 points, 31 resident columns) and the generic gate over d4 (2^18 points); device time per evaluation, and the oracle's
(all host threads) on a 2^16-point slice for scale."""
import json, os, sys, time
import numpy as np, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import proof_systems_b200 as zk
import gate_programs as gp
from bench import splitmix64_limbs
from oracle import oracle as orc
LOG_N = int(os.environ.get("LOG_N", "16")); n = 1 << LOG_N; m = 8 * n
ctx = zk.Context(0)
stream = torch.cuda.Stream(); ctx.set_stream(stream.cuda_stream)
def timed(fn, reps=7):
    ts = []
    for _ in range(reps + 2):
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream); fn(); e1.record(stream); e1.synchronize(); ts.append(e0.elapsed_time(e1))
    return float(np.median(ts[2:]))
fid = zk.FP
mont = lambda a: orc.to_mont(fid, a)
cols_h = [mont(splitmix64_limbs(10 + k, m) & np.uint64((1 << 62) - 1)) for k in range(31)]   # any residues do
dev = [torch.from_numpy(a.view(np.int64)).cuda() for a in cols_h]
alphas, mds = mont(splitmix64_limbs(3, 15) >> np.uint64(2)), mont(splitmix64_limbs(4, 9) >> np.uint64(2)).reshape(3, 3, 4)
out = torch.empty((m, 4), dtype=torch.int64, device="cuda")
rows = {}
pos = gp.poseidon_gate(zk.ExprProgram(), alphas, mds)
cols8 = [(d.data_ptr(), m, 8) for d in dev]
rows["poseidon gate over d8, 2^%d points, %d tokens" % (LOG_N + 3, len(pos.tokens))] = {"gpu_ms": round(timed(lambda: pos.evaluations(ctx, fid, cols8, m, 8, out.data_ptr())), 3)}
gen = gp.generic_gate(zk.ExprProgram(), alphas[:2])
cols4 = cols8[:30] + [(dev[30].data_ptr(), 4 * n, 4)]
rows["generic gate over d4, 2^%d points, %d tokens" % (LOG_N + 2, len(gen.tokens))] = {"gpu_ms": round(timed(lambda: gen.evaluations(ctx, fid, cols4, 4 * n, 4, out.data_ptr())), 3)}
# the oracle on a slice (same program, 2^16 points of a d8 over 2^13 rows), scaled
rec = gp.poseidon_gate(gp.Recorder(), alphas, mds)
sl = 1 << 16
t0 = time.perf_counter(); orc.expr_eval(fid, rec.ops, rec.args, rec.literals, [(c[:sl], 8) for c in cols_h], sl); dt = time.perf_counter() - t0
rows["poseidon gate, CPU oracle (%d threads), scaled from 2^16 points" % orc.host_threads()] = {"cpu_ms": round(1e3 * dt * m / sl, 1)}
print(json.dumps(rows, indent=1))
json.dump(rows, open(os.path.join(ROOT, "gpurun_out", "expr_time.json"), "w"), indent=1)
