#!/usr/bin/env python3
"""BASELINE config 5 proxy: replay the MSM / NTT schedule that kimchi's ProverProof::create_recursive issues for a
2^16-gate circuit (kimchi/src/prover.rs:187-1515, operands as in kimchi/src/bench.rs:59-125), once on the
GPU library (host pointers in, host results out: the calls the Rust shim of INTEGRATION.md would make) and once on the
CPU oracle (all host threads), and report seconds for the replayed portion.  No Rust toolchain exists in this image, so
the protocol logic between the calls (sponges, expression evaluation, base folding of `open`) is NOT replayed.

Schedule per proof (n = 2^16, curve Vesta, scalar field Fp, no lookups, one chunk):
  15 x commit_evaluations_non_hiding  MSM on the Lagrange basis, scalars {1 x (n-10), 0 x 7, random x 3}   prover.rs:329-351
  15 x iFFT(n)                         witness interpolation                                                prover.rs:370-381
   1 x iFFT(n) + 1 x MSM(n) dense      permutation aggregation z, commit                                    prover.rs:679-682
  16 x FFT(8n) from n coefficients     evaluate over d8                                                     constraints.rs:488-507
   1 x iFFT(4n) + 1 x iFFT(8n)         quotient                                                             prover.rs:907
   7 x MSM(n) dense (shared bases)     t commitment                                                         prover.rs:923
   2 x iFFT(n)                         ft polynomial, combine_polys                                         prover.rs:1163, utils.rs:195-198
  16 x 2 MSMs of n/2^(r+1) + 2 points  IPA rounds L and R on the (folded, non-resident) bases               ipa.rs:943-961
"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import proof_systems_b200 as zk
from oracle import oracle as orc

LOG_N = int(os.environ.get("LOG_N", "16"))
N = 1 << LOG_N
CID, FS = orc.VESTA, orc.FP          # kimchi over Vesta: scalars in Fp


def inputs():
    z = np.load(os.path.join(ROOT, "tests", "golden", "vesta_srs.npz"))
    g0 = orc.decompress(CID, z["g_cmp"].tobytes())
    g = orc.extend_bases(CID, g0, N)                      # stand-in SRS (the fixture keeps 2048 Vesta generators)
    lag = orc.extend_bases(CID, g0[::-1].copy(), N)       # stand-in Lagrange basis: any n on-curve points
    wit = np.zeros((15, N, 4), dtype=np.uint64)
    wit[:, : N - 10, 0] = 1                               # Montgomery form of 1 is not 1: convert below
    wit_m = orc.to_mont(FS, wit.reshape(-1, 4)).reshape(15, N, 4)
    wit_m[:, N - 3:] = orc.to_mont(FS, orc.random_scalars(FS, 45, seed=5)).reshape(15, 3, 4)
    dense = orc.to_mont(FS, orc.random_scalars(FS, 8 * N, seed=6)).reshape(8, N, 4)     # z + 7 chunks of t
    big = orc.to_mont(FS, orc.random_scalars(FS, 8 * N, seed=7))
    return g, lag, wit_m, dense, big


def pinned(a):
    """copy of `a` in page-locked host memory (what a production caller would hand to the library)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).pin_memory()
    return t.numpy().view(np.uint64)


def run_gpu(g, lag, wit_m, dense, big):
    ctx = zk.Context(0)
    t_setup = time.perf_counter()
    srs = zk.SRS(ctx, CID, g, g[0])
    srs.add_lagrange_basis(N, lag)
    g_table = ctx.upload_bases(CID, g)                    # the IPA rounds read the SRS table (the bases are never folded)
    setup = time.perf_counter() - t_setup
    out = {}

    def stage(name, fn):
        t0 = time.perf_counter()
        r = fn()
        out[name] = time.perf_counter() - t0
        return r

    # working buffers in pinned host memory; transforms run in place on them (a Rust caller transforms its own Vec)
    p_wit, p_dense, p_big = pinned(wit_m), pinned(dense), pinned(big)
    p_pad = pinned(np.zeros((16, 8 * N, 4), dtype=np.uint64))
    p_wit2, p_big4 = pinned(wit_m), pinned(big[: 4 * N])

    def once():
        stage("15 witness commitments", lambda: srs.commit_evaluations_non_hiding_batch(N, p_wit))
        stage("15 iFFT(n)", lambda: ctx.ntt_inplace(FS, p_wit2, inverse=True))
        stage("z: iFFT(n) + MSM", lambda: (ctx.ntt_inplace(FS, p_dense[0], inverse=True), srs.commit_non_hiding(p_dense[0], 1)))
        p_pad[:15, :N] = p_wit
        p_pad[15, :N] = p_dense[0]
        stage("16 FFT(8n)", lambda: ctx.ntt_inplace(FS, p_pad, in_len=N))
        stage("iFFT(4n) + iFFT(8n)", lambda: (ctx.ntt_inplace(FS, p_big4, inverse=True), ctx.ntt_inplace(FS, p_big, inverse=True)))
        stage("t: 7 MSMs", lambda: srs.commit_non_hiding(p_dense[1:].reshape(-1, 4), 7))
        stage("2 iFFT(n)", lambda: ctx.ntt_inplace(FS, p_dense[:2], inverse=True))

        def open_rounds():
            # the folding loop of SRS::open with a, b resident and the bases taken from the SRS table (csrc/ipa.cu); the
            # challenges are stand-ins (the sponge is the caller's)
            rounds = zk.IpaRounds(ctx, g_table, p_dense[0], p_dense[1])
            res = []
            for r in range(LOG_N):
                res.append(rounds.lr())
                rounds.fold(orc_scalars[2 * r], orc_scalars[2 * r + 1])
            res.append(rounds.sg())
            rounds.close()
            return res
        stage("open: 2 x 16 MSMs", open_rounds)

    global orc_scalars
    orc_scalars = orc.random_scalars(FS, N, seed=8)
    once()                    # warm-up (tables, allocations)
    out.clear()
    t0 = time.perf_counter()
    once()
    total = time.perf_counter() - t0
    return {"setup_s": setup, "total_s": total, "stages_s": out, "kernel_launches": ctx.launch_count}


def run_cpu(g, lag, wit_m, dense, big):
    th = orc.host_threads()
    out, used = {}, {}
    cands = sorted({t for t in (1, 8, 32, th) if t <= th})

    def stage(name, fn):
        """every stage gets its best thread count (the oracle's OpenMP loops do not scale to 128 threads on small inputs)"""
        best = None
        for t in cands:
            t0 = time.perf_counter()
            fn(t)
            el = time.perf_counter() - t0
            if best is None or el < best:
                best, used[name] = el, t
        out[name] = best

    sc = orc.random_scalars(FS, N, seed=8)
    stage("15 witness commitments", lambda t: [orc.msm_mont(CID, lag, wit_m[k], threads=t) for k in range(15)])
    stage("15 iFFT(n)", lambda t: [orc.ntt(FS, wit_m[k], inverse=True, threads=t) for k in range(15)])
    stage("z: iFFT(n) + MSM", lambda t: (orc.ntt(FS, dense[0], inverse=True, threads=t), orc.msm_split2(CID, g, orc.from_mont(FS, dense[0]), threads=t)))

    def fft8(t):
        for k in range(16):
            pad = np.zeros((8 * N, 4), dtype=np.uint64)
            pad[:N] = wit_m[k] if k < 15 else dense[0]
            orc.ntt(FS, pad, threads=t)
    stage("16 FFT(8n)", fft8)
    stage("iFFT(4n) + iFFT(8n)", lambda t: (orc.ntt(FS, big[: 4 * N], inverse=True, threads=t), orc.ntt(FS, big, inverse=True, threads=t)))
    stage("t: 7 MSMs", lambda t: [orc.msm_mont(CID, g, dense[1 + k], threads=t) for k in range(7)])
    stage("2 iFFT(n)", lambda t: [orc.ntt(FS, dense[k], inverse=True, threads=t) for k in range(2)])

    def open_rounds(t):
        for r in range(LOG_N):
            m = (N >> (r + 1)) + 2
            orc.msm(CID, g[:m], sc[:m], threads=t)
            orc.msm(CID, g[N - m:], sc[:m], threads=t)
    stage("open: 2 x 16 MSMs (no base folding counted)", open_rounds)
    return {"total_s": sum(out.values()), "stages_s": out, "threads_per_stage": used, "host_threads": th}



if __name__ == "__main__":
    data = inputs()
    gpu = run_gpu(*data)
    cpu = run_cpu(*data)
    rep = {"log_n": LOG_N, "gpu": gpu, "cpu_oracle": cpu, "speedup_replayed_portion": cpu["total_s"] / gpu["total_s"],
           "note": "MSM/NTT schedule of one kimchi proof (SURVEY.md 3.1); protocol logic between the calls is not replayed; "
                   "reference's published whole-prover time for 2^16 gates: 6.3 s (README.md:41, unspecified hardware)"}
    print(json.dumps(rep, indent=1))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "replay_kimchi.json"), "w"), indent=1)
