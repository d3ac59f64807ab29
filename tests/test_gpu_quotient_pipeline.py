"""SURVEY.md §8f row 3 end to end: one quotient polynomial computed without leaving the device — the prover's sequence
kimchi/src/prover.rs:370-381 (iFFT(n) of the columns) -> circuits/constraints.rs:488-507 (FFT(8n)) -> prover.rs:794-892 (gate constraints
through the expression evaluator into t4 / t8, the permutation part through its kernel) -> prover.rs:905-918 (iFFT(4n) + iFFT(8n),
+ public, division by the vanishing polynomial, + bnd) -> prover.rs:921 (the 7 chunk commitments of t).  ONE upload of the d1
evaluations, ONE download of 7 points; every intermediate the test reads back is compared bit for bit with the oracle's chain
(ntt, expr_eval, perm_quot, divide_by_vanishing, msm)."""
import numpy as np
import pytest

import gate_programs as gp
import proof_systems_b200 as zk

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ctx():
    c = zk.Context(0)
    yield c
    c.close()


@pytest.mark.parametrize("length,log_n", [(8 * 7 + 3, 3), (64, 3), (8, 3), (5, 3), (4096 * 8, 12), (4096 * 7 + 1, 12), (9, 0)])
def test_divide_by_vanishing_vs_oracle(ctx, orc, length, log_n):
    fid, n = zk.FQ, 1 << log_n
    f = orc.to_mont(fid, orc.random_scalars(fid, length, seed=length))
    want_q, want_r = orc.divide_by_vanishing(fid, f, log_n)
    d_f, d_q = ctx.dev_alloc(f.nbytes), ctx.dev_alloc(max(length - n, 1) * 32)
    try:
        ctx.dev_upload(d_f, f)
        zero_rem = ctx.poly_divide_by_vanishing_dev(fid, d_f, length, log_n, d_q)
        assert zero_rem == (not want_r.any())
        if length > n:
            assert np.array_equal(ctx.dev_download(d_q, (length - n, 4)), want_q)
    finally:
        ctx.dev_free(d_f); ctx.dev_free(d_q)


def test_exact_multiple_of_the_vanishing_polynomial_has_zero_remainder(ctx, orc):
    """f = q (x^n - 1) built on the host: the call returns q and reports a zero remainder — the prover's check at prover.rs:910-914"""
    fid, log_n = zk.FP, 10
    n, P = 1 << log_n, orc.FP_MODULUS
    q = orc.limbs_to_ints(orc.random_scalars(fid, 7 * n, seed=5))
    f = [0] * (8 * n)
    for j, c in enumerate(q):
        f[j + n] = (f[j + n] + c) % P
        f[j] = (f[j] - c) % P
    fm, qm = orc.to_mont(fid, orc.ints_to_limbs(f)), orc.to_mont(fid, orc.ints_to_limbs(q))
    d_f, d_q = ctx.dev_alloc(fm.nbytes), ctx.dev_alloc(qm.nbytes)
    try:
        ctx.dev_upload(d_f, fm)
        assert ctx.poly_divide_by_vanishing_dev(fid, d_f, 8 * n, log_n, d_q) is True
        assert np.array_equal(ctx.dev_download(d_q, (7 * n, 4)), qm)
        fm[3] = orc.to_mont(fid, orc.ints_to_limbs([(f[3] + 1) % P]))[0]                # one coefficient off: remainder != 0
        ctx.dev_upload(d_f, fm)
        assert ctx.poly_divide_by_vanishing_dev(fid, d_f, 8 * n, log_n, d_q) is False
    finally:
        ctx.dev_free(d_f); ctx.dev_free(d_q)


def test_one_quotient_polynomial_without_leaving_the_device(ctx, orc, vesta_srs):
    fid, log_n = zk.FP, 9                                   # Vesta's scalar field
    n, m4, m8 = 1 << log_n, 4 << log_n, 8 << log_n
    rnd = lambda k, seed: orc.to_mont(fid, orc.random_scalars(fid, k, seed=seed))
    cols = rnd(16 * n, 11).reshape(16, n, 4)                # w_0..w_14 and z as evaluations over d1
    coeff8 = rnd(15 * m8, 12).reshape(15, m8, 4)            # per-index arrays, resident in a real prover (zk_index_cache_section)
    sigma8, zkpm = rnd(7 * m8, 13).reshape(7, m8, 4), rnd(m8, 14)
    gen_sel4, pos_sel8 = rnd(m4, 15), rnd(m8, 16)
    beta, gamma, alpha0, shifts = rnd(1, 17)[0], rnd(1, 18)[0], rnd(1, 19)[0], rnd(7, 20)
    alphas, mds = rnd(17, 21), rnd(9, 22).reshape(3, 3, 4)
    public, bnd = rnd(n, 23), rnd(7 * n, 24)
    G = vesta_srs
    # ---------------------------------------------------------------- oracle
    coeffs = np.stack([orc.ntt(fid, cols[j], inverse=True) for j in range(16)])
    ev8 = []
    for j in range(16):
        pad = np.zeros((m8, 4), dtype=np.uint64); pad[:n] = coeffs[j]
        ev8.append(orc.ntt(fid, pad))
    gen = gp.generic_gate(gp.Recorder(), alphas[:2])
    pos = gp.poseidon_gate(gp.Recorder(), alphas[2:], mds)
    cols_gen = [(ev8[k], 8) for k in range(15)] + [(coeff8[k], 8) for k in range(15)] + [(gen_sel4, 4)]
    cols_pos = cols_gen[:30] + [(pos_sel8, 8)]
    t4 = orc.expr_eval(fid, gen.ops, gen.args, gen.literals, cols_gen, m4)
    t8 = orc.perm_quot(fid, np.stack(ev8[:7]), ev8[15], sigma8, zkpm, beta, gamma, alpha0, shifts)
    t8 = orc.expr_eval(fid, pos.ops, pos.args, pos.literals, cols_pos, m8, acc=t8)
    t4c, f = orc.ntt(fid, t4, inverse=True), orc.ntt(fid, t8, inverse=True)
    add = lambda a, b: orc.to_mont(fid, orc.ints_to_limbs([(x + y) % orc.FP_MODULUS for x, y in zip(orc.limbs_to_ints(orc.from_mont(fid, a)), orc.limbs_to_ints(orc.from_mont(fid, b)))]))
    f[:m4] = add(f[:m4], t4c)
    f[:n] = add(f[:n], public)
    quot, rem = orc.divide_by_vanishing(fid, f, log_n)
    assert rem.any()                                        # random columns satisfy no circuit: the prover would stop here
    quot = add(quot, bnd)
    want_comm = [orc.msm(G.cid, G.g[:n], orc.from_mont(fid, quot[c * n:(c + 1) * n])) for c in range(7)]
    # ---------------------------------------------------------------- device
    bufs = []
    def put(a):
        p = ctx.dev_alloc(a.nbytes); bufs.append(p); ctx.dev_upload(p, a); return p
    def alloc(nb):
        p = ctx.dev_alloc(nb); bufs.append(p); return p
    bases = ctx.upload_bases(G.cid, G.g[:n])
    try:
        d_cols = put(cols)                                                             # the ONE upload of per-proof data
        d_coeff8, d_sigma, d_zkpm, d_gsel, d_psel = put(coeff8), put(sigma8), put(zkpm), put(gen_sel4), put(pos_sel8)   # per-index, resident
        d_public, d_bnd = put(public), put(bnd)
        d_ev8, d_t4, d_t8, d_q = alloc(16 * m8 * 32), alloc(m4 * 32), alloc(m8 * 32), alloc(7 * n * 32)
        ctx.ntt_dev(fid, d_cols, log_n, batch=16, inverse=True)                         # prover.rs:370-381
        ctx.ntt_dev_oop(fid, d_cols, n, n, d_ev8, log_n + 3, batch=16)                  # constraints.rs:488-507
        w8 = [(d_ev8 + k * m8 * 32, m8, 8) for k in range(15)]
        c8 = [(d_coeff8 + k * m8 * 32, m8, 8) for k in range(15)]
        gp.generic_gate(zk.ExprProgram(), alphas[:2]).evaluations(ctx, fid, w8 + c8 + [(d_gsel, m4, 4)], m4, 4, d_t4)               # prover.rs:794-812
        ctx.perm_quotient_dev(fid, log_n + 3, [w[0] for w in w8[:7]], d_ev8 + 15 * m8 * 32, [d_sigma + k * m8 * 32 for k in range(7)], d_zkpm,
                              beta, gamma, alpha0, shifts, d_t8)                                                                  # prover.rs:815-824
        gp.poseidon_gate(zk.ExprProgram(), alphas[2:], mds).evaluations(ctx, fid, w8 + c8 + [(d_psel, m8, 8)], m8, 8, d_t8, accumulate=True)   # :826-882
        assert np.array_equal(ctx.dev_download(d_t4, (m4, 4)), t4) and np.array_equal(ctx.dev_download(d_t8, (m8, 4)), t8)
        ctx.ntt_dev(fid, d_t4, log_n + 2, inverse=True)                                 # prover.rs:906: t4.interpolate() + t8.interpolate()
        ctx.ntt_dev(fid, d_t8, log_n + 3, inverse=True)
        ctx.poly_add_dev(fid, d_t8, d_t4, m4)
        ctx.poly_add_dev(fid, d_t8, d_public, n)                                        # f += &public_poly
        assert ctx.poly_divide_by_vanishing_dev(fid, d_t8, m8, log_n, d_q) is False     # prover.rs:909-914
        ctx.poly_add_dev(fid, d_q, d_bnd, 7 * n)                                        # quotient += &bnd
        assert np.array_equal(ctx.dev_download(d_q, (7 * n, 4)), quot)
        got = [zk.jacobian_to_affine(G.cid, ctx.msm_dev(bases, d_q + c * n * 32, n, mont=True)) for c in range(7)]   # prover.rs:921, commit_non_hiding
        for c in range(7):
            assert np.array_equal(got[c], want_comm[c]), c
    finally:
        bases.free()
        for p in bufs:
            ctx.dev_free(p)
