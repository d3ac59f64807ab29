#!/usr/bin/env python3
"""BASELINE config 5 proxy: replay the MSM / NTT schedule that kimchi's ProverProof::create_recursive issues for a 2^16-gate
circuit (kimchi/src/prover.rs:187-1515, operands as in kimchi/src/bench.rs:59-125) on the GPU library — host
pointers in, host results out: the calls the Rust shim (crates/zkb200) makes — and on the CPU oracle, and CHECK every stage's output
bit for bit against the oracle.  No Rust toolchain exists in this image, so the protocol logic between the calls (sponges,
expression evaluation) is not replayed; the opening proof runs through zk_srs_open with a stand-in transcript.

The SRS is the reference's own: all 2^16 Vesta generators of srs/vesta.srs (tests/golden/vesta_srs.npz) and the Lagrange basis of the
2^16 domain, computed on the device from them and required to equal the basis stored in srs/test_vesta.srs (sha256 of all 65536
entries, pinned in the fixture).

Schedule per proof (n = 2^16, curve Vesta, scalar field Fp, no lookups, one chunk):
  15 x commit_evaluations_non_hiding  MSM on the Lagrange basis, scalars {1 x (n-10), 0 x 7, random x 3}   prover.rs:329-351
  15 x iFFT(n)                         witness interpolation                                                prover.rs:370-381
   1 x iFFT(n) + 1 x MSM(n) dense      permutation aggregation z, commit                                    prover.rs:679-682
  16 x FFT(8n) from n coefficients     evaluate over d8                                                     constraints.rs:488-507
   1 x iFFT(4n) + 1 x iFFT(8n)         quotient                                                             prover.rs:907
   7 x MSM(n) dense (shared bases)     t commitment                                                         prover.rs:923
   2 x iFFT(n)                         ft polynomial, combine_polys                                         prover.rs:1163, utils.rs:195-198
   1 x SRS::open                       45 polynomials, 2 evaluation points, 16 rounds                       prover.rs:1279-1345, ipa.rs:823-1061

    python tools/replay_kimchi.py            # timing + checks, writes gpurun_out/replay_kimchi.json
    tests/test_gpu_replay.py                 # the same function under pytest (checks only)
"""
import hashlib
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LOG_N = int(os.environ.get("LOG_N", "16"))


def pinned(a):
    """copy of `a` in page-locked host memory (what a production caller would hand to the library)"""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(a).view(np.int64)).pin_memory()
    return t.numpy().view(np.uint64)


def replay(zk, orc, ctx, log_n=LOG_N, check=True, n_open_polys=45):
    """Runs the schedule once (after one warm-up pass) and returns {"stages_s": ..., "checks": ...}; with check=True every stage output
    is compared with the oracle and an AssertionError names the first stage that differs."""
    N = 1 << log_n
    CID, FS = orc.VESTA, orc.FP
    z = np.load(os.path.join(ROOT, "tests", "golden", "vesta_srs.npz"))
    assert z["g_cmp"].shape[0] >= N, "the fixture must hold all generators of the domain"
    g = ctx.decompress_points(zk.VESTA, z["g_cmp"][:N])
    h = orc.to_mont(orc.FQ, np.ascontiguousarray(z["h_xy"]).view("<u8").reshape(2, 4)).reshape(8)
    t0 = time.perf_counter()
    srs = zk.SRS(ctx, CID, g, h)
    lag = srs.get_lagrange_basis_from_domain_size(N)                    # device: group iFFT of the generators (ipa.rs:1065-1172)
    setup_s = time.perf_counter() - t0
    checks = {}
    if log_n == 16:
        canon = orc.from_mont(orc.FQ, lag.reshape(-1, 4)).astype("<u8").tobytes()
        checks["lagrange_basis_equals_srs_test_vesta"] = hashlib.sha256(canon).digest() == bytes(z["lag_65536_sha256"])
        assert checks["lagrange_basis_equals_srs_test_vesta"], "the device's 2^16 Lagrange basis differs from srs/test_vesta.srs"

    wit = np.zeros((15, N, 4), dtype=np.uint64)
    wit[:, : N - 10, 0] = 1
    wit_m = orc.to_mont(FS, wit.reshape(-1, 4)).reshape(15, N, 4)
    wit_m[:, N - 3:] = orc.to_mont(FS, orc.random_scalars(FS, 45, seed=5)).reshape(15, 3, 4)
    dense = orc.to_mont(FS, orc.random_scalars(FS, 8 * N, seed=6)).reshape(8, N, 4)     # z + 7 chunks of t
    big = orc.to_mont(FS, orc.random_scalars(FS, 8 * N, seed=7))
    rnd = lambda k, seed: orc.to_mont(FS, orc.random_scalars(FS, k, seed=seed))
    # opening: 45 entries like prover.rs:1279-1338 — here 8 dense polynomials referenced repeatedly plus the 15 witness columns in
    # coefficient form (after their iFFT), one blinder each; two evaluation points (zeta, zeta * omega)
    open_bl = rnd(n_open_polys, 31)
    elm, polyscale, evalscale = rnd(2, 32), rnd(1, 33)[0], rnd(1, 34)[0]
    draws = rnd(2 * log_n + 2, 35)
    chals = []

    def transcript():
        chals.clear()
        st = [11]

        def nxt():
            st[0] = (st[0] * 6364136223846793005 + 1442695040888963407) % (1 << 64)
            return st[0]
        u_base = lambda cip: g[7]

        def rc(i, l, r):
            u = orc.to_mont(FS, orc.ints_to_limbs([nxt() * (1 << 64) + nxt()]))[0]
            chals.append(u)
            return u
        fc = lambda delta: orc.to_mont(FS, orc.ints_to_limbs([nxt() + 1]))[0]
        return u_base, rc, fc

    p_wit, p_dense, p_big = pinned(wit_m), pinned(dense), pinned(big)
    p_pad = pinned(np.zeros((16, 8 * N, 4), dtype=np.uint64))
    p_wit2, p_big4 = pinned(wit_m), pinned(big[: 4 * N])
    out, res = {}, {}

    def stage(name, fn):
        t0 = time.perf_counter()
        r = fn()
        out[name] = time.perf_counter() - t0
        res[name] = r
        return r

    def once():
        p_wit2[:] = wit_m
        p_dense[:] = dense
        p_big[:] = big
        p_big4[:] = big[: 4 * N]
        stage("15 witness commitments", lambda: srs.commit_evaluations_non_hiding_batch(N, p_wit))
        stage("15 iFFT(n)", lambda: ctx.ntt_inplace(FS, p_wit2, inverse=True))
        stage("z: iFFT(n) + MSM", lambda: (ctx.ntt_inplace(FS, p_dense[0], inverse=True), srs.commit_non_hiding(p_dense[0], 1))[1])
        p_pad[:15, :N] = p_wit2
        p_pad[15, :N] = p_dense[0]
        stage("16 FFT(8n)", lambda: ctx.ntt_inplace(FS, p_pad, in_len=N))
        stage("iFFT(4n) + iFFT(8n)", lambda: (ctx.ntt_inplace(FS, p_big4, inverse=True), ctx.ntt_inplace(FS, p_big, inverse=True)))
        stage("t: 7 MSMs", lambda: srs.commit_non_hiding(p_dense[1:].reshape(-1, 4), 7))
        stage("2 iFFT(n)", lambda: ctx.ntt_inplace(FS, p_dense[1:3], inverse=True))
        polys = [(p_wit2[k % 15] if k < 30 else p_dense[k % 8], 0, open_bl[k:k + 1]) for k in range(n_open_polys)]
        stage("open (45 polynomials, 16 rounds)", lambda: zk.srs_open(srs, polys, elm, polyscale, evalscale, draws, *transcript()))

    once()                    # warm-up (tables, allocations, lanes)
    out.clear()
    t0 = time.perf_counter()
    once()
    total = time.perf_counter() - t0
    report = {"log_n": log_n, "setup_s": setup_s, "total_s": total, "stages_s": dict(out), "kernel_launches": ctx.launch_count}

    # ---- the same NTT stages with the columns RESIDENT: one upload of the 16 evaluation columns, then
    #      iFFT(n) -> FFT(8n) (out of place into the d8 buffers the quotient reads) -> iFFT(4n) + iFFT(8n), nothing leaves the device.
    #      Results are compared with what the host-path stages above produced (themselves checked against the oracle below).
    cols16 = np.concatenate([wit_m, dense[0:1]])
    d_cols, d_ev8 = ctx.dev_alloc(cols16.nbytes), ctx.dev_alloc(16 * 8 * N * 32)
    d_big, d_big4 = ctx.dev_alloc(8 * N * 32), ctx.dev_alloc(4 * N * 32)
    res_r = {}

    def rstage(name, fn):
        sync = lambda: ctx.dev_download(d_cols, (1, 4))       # a 32-byte read-back: waits for everything queued on the context's stream
        sync()
        t0 = time.perf_counter()
        fn()
        sync()
        res_r[name] = time.perf_counter() - t0
    try:
        for rep in range(2):                                  # the second pass is the measurement (tables and buffers warm)
            rstage("upload of the 16 evaluation columns (32 B x 16 n)", lambda: ctx.dev_upload(d_cols, cols16))
            ctx.dev_upload(d_big, big); ctx.dev_upload(d_big4, big[: 4 * N])
            rstage("16 iFFT(n), resident", lambda: ctx.ntt_dev(FS, d_cols, log_n, batch=16, inverse=True))
            rstage("16 FFT(8n), resident", lambda: ctx.ntt_dev_oop(FS, d_cols, N, N, d_ev8, log_n + 3, batch=16))
            rstage("iFFT(4n) + iFFT(8n), resident", lambda: (ctx.ntt_dev(FS, d_big4, log_n + 2, inverse=True), ctx.ntt_dev(FS, d_big, log_n + 3, inverse=True)))
        got = ctx.dev_download(d_cols, (16, N, 4))
        assert np.array_equal(got[:15], p_wit2) and np.array_equal(got[15], p_dense[0]), "resident iFFT(n) differs from the host path"
        assert np.array_equal(ctx.dev_download(d_ev8, (16, 8 * N, 4)), p_pad), "resident FFT(8n) differs from the host path"
        assert np.array_equal(ctx.dev_download(d_big, (8 * N, 4)), p_big) and np.array_equal(ctx.dev_download(d_big4, (4 * N, 4)), p_big4)
    finally:
        for ptr in (d_cols, d_ev8, d_big, d_big4):
            ctx.dev_free(ptr)
    report["resident_stages_s"] = res_r
    if not check:
        srs.close()
        return report

    # ---------------------------------------------------------------------------------------------- the oracle, stage by stage
    th = orc.host_threads()
    cpu = {}

    def cstage(name, fn):
        t0 = time.perf_counter()
        r = fn()
        cpu[name] = time.perf_counter() - t0
        return r

    want = cstage("15 witness commitments", lambda: [orc.msm_mont(CID, lag, wit_m[k], threads=th) for k in range(15)])
    for k in range(15):
        assert np.array_equal(res["15 witness commitments"][k].chunks[0], want[k]), ("witness commitment", k)
    w_coeffs = cstage("15 iFFT(n)", lambda: np.stack([orc.ntt(FS, wit_m[k], inverse=True, threads=th) for k in range(15)]))
    assert np.array_equal(p_wit2, w_coeffs), "15 iFFT(n)"
    z_coeffs = cstage("z: iFFT(n) + MSM", lambda: orc.ntt(FS, dense[0], inverse=True, threads=th))
    assert np.array_equal(p_dense[0], z_coeffs), "z iFFT"
    assert np.array_equal(res["z: iFFT(n) + MSM"].chunks[0], orc.msm_mont(CID, g, z_coeffs, threads=th)), "z commitment"

    def fft8():
        for k in range(16):
            pad = np.zeros((8 * N, 4), dtype=np.uint64)
            pad[:N] = w_coeffs[k] if k < 15 else z_coeffs
            assert np.array_equal(p_pad[k], orc.ntt(FS, pad, threads=th)), ("FFT(8n)", k)
    cstage("16 FFT(8n)", fft8)
    cstage("iFFT(4n) + iFFT(8n)", lambda: (np.testing.assert_array_equal(p_big4, orc.ntt(FS, big[: 4 * N], inverse=True, threads=th)),
                                            np.testing.assert_array_equal(p_big, orc.ntt(FS, big, inverse=True, threads=th))))
    t_want = cstage("t: 7 MSMs", lambda: [orc.msm_mont(CID, g, dense[1 + k], threads=th) for k in range(7)])
    assert np.array_equal(res["t: 7 MSMs"].chunks, np.stack(t_want)), "t commitment"
    d12 = cstage("2 iFFT(n)", lambda: np.stack([orc.ntt(FS, dense[1 + k], inverse=True, threads=th) for k in range(2)]))
    assert np.array_equal(p_dense[1:3], d12), "2 iFFT(n)"

    # ---- the opening proof: sg = <b_poly_coefficients(chals), g> (commitment.rs:565-581), round 0's L and R as the reference's two
    #      (n/2 + 2)-point MSMs (ipa.rs:938-960), z1 / z2 from a0, r_prime (ipa.rs:1046-1052): Python integers + oracle MSMs
    proof = res["open (45 polynomials, 16 rounds)"]
    m = orc.MODULUS[FS]
    ints = lambda a: orc.limbs_to_ints(orc.from_mont(FS, np.ascontiguousarray(a).reshape(-1, 4)))
    p_dense_host = [p_wit2[k % 15] if k < 30 else (z_coeffs if k % 8 == 0 else d12[k % 8 - 1] if k % 8 in (1, 2) else dense[k % 8]) for k in range(n_open_polys)]
    ps, es = ints(polyscale)[0], ints(evalscale)[0]

    def cpu_open():
        a = np.zeros(N, dtype=object)
        scale, comb = 1, 0
        bl = ints(open_bl)
        for k in range(n_open_polys):
            a = (a + scale * np.array(ints(p_dense_host[k]), dtype=object)) % m
            comb = (comb + bl[k] * scale) % m
            scale = scale * ps % m
        b = np.zeros(N, dtype=object)
        sc = 1
        for e in ints(elm):
            pw = np.empty(N, dtype=object)
            cur = 1
            for i in range(N):
                pw[i] = cur
                cur = cur * e % m
            b = (b + sc * pw) % m
            sc = sc * es % m
        hh = N // 2
        dr = ints(draws)
        ip_l = int(np.dot(a[hh:], b[:hh]) % m)
        ip_r = int(np.dot(a[:hh], b[hh:]) % m)
        l0 = orc.msm(CID, np.concatenate([g[:hh], h[None], g[7][None]]), orc.ints_to_limbs([int(x) for x in a[hh:]] + [dr[0], ip_l]), threads=th)
        r0 = orc.msm(CID, np.concatenate([g[hh:], h[None], g[7][None]]), orc.ints_to_limbs([int(x) for x in a[:hh]] + [dr[1], ip_r]), threads=th)
        us = ints(np.stack(chals))
        # fold a and b with the recorded challenges to a0, b0 (ipa.rs:980-1003)
        for u in us:
            ui = pow(u, -1, m)
            half = len(a) // 2
            a = (a[:half] + ui * a[half:]) % m
            b = (b[:half] + u * b[half:]) % m
        s = [1]
        for u in us:
            s = [v for t in s for v in (t, t * u % m)]
        sg = orc.msm(CID, g, orc.ints_to_limbs(s), threads=th)
        r_prime = comb
        for r, u in enumerate(us):
            r_prime = (r_prime + dr[2 * r] * pow(u, -1, m) + dr[2 * r + 1] * u) % m
        return l0, r0, sg, int(a[0]), int(b[0]), r_prime, dr
    l0, r0, sg, a0, b0, r_prime, dr = cstage("open (45 polynomials, 16 rounds)", cpu_open)
    assert np.array_equal(proof.lr[0, 0], l0) and np.array_equal(proof.lr[0, 1], r0), "open: round 0"
    assert np.array_equal(proof.sg, sg), "open: sg"
    # the stand-in transcript's final challenge c = nxt() + 1 is the last draw of its generator: recompute z1, z2 from the proof's own c
    # through the verifier's identity  z1 = a0 c + d,  z2 = r_prime c + r_delta
    c_num = (ints(proof.z1)[0] - dr[-2]) * pow(a0, -1, m) % m
    assert ints(proof.z2)[0] == (r_prime * c_num + dr[-1]) % m, "open: z1 / z2"
    # delta = d (g0 + b0 U) + r_delta h
    delta = orc.msm(CID, np.stack([sg, g[7], h]), orc.ints_to_limbs([dr[-2], dr[-2] * b0 % m, dr[-1]]))
    assert np.array_equal(proof.delta, delta), "open: delta"
    checks.update({k: True for k in out})
    report["cpu_oracle"] = {"stages_s": cpu, "host_threads": th,
                            "note": "the oracle's time per stage INCLUDES the comparison; the open row is the checker's subset (2 MSMs of n/2+2, sg, folds in Python), not a CPU open"}
    report["checks"] = checks
    srs.close()
    return report


if __name__ == "__main__":
    import proof_systems_b200 as zk
    from oracle import oracle as orc
    ctx = zk.Context(0)
    rep = replay(zk, orc, ctx, LOG_N, check=os.environ.get("REPLAY_CHECK", "1") != "0")
    rep["note"] = ("MSM/NTT schedule of one kimchi proof (SURVEY.md 3.1), every stage compared bit for bit with the CPU oracle; protocol logic "
                   "between the calls is not replayed; reference's published whole-prover time for 2^16 gates: 6.3 s (README.md:41, unspecified hardware)")
    print(json.dumps(rep, indent=1, default=str))
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(rep, open(os.path.join(ROOT, "gpurun_out", "replay_kimchi.json"), "w"), indent=1, default=str)
