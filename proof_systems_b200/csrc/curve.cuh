// curve.cuh — Pallas / Vesta group law for the MSM kernels: y^2 = x^3 + 5 (a = 0), prime order
// (curves/src/pasta/curves/pallas.rs:19-46, vesta.rs:18-45).
//
// The reference accumulates buckets in Jacobian coordinates (ark-ec short_weierstrass::Projective, reached from
// poly-commitment/src/ipa.rs:672 etc.).  Here buckets live in XYZZ coordinates (x = X/ZZ, y = Y/ZZZ, ZZ^3 = ZZZ^2):
// the mixed addition costs 8M + 2S instead of 7M + 4S-with-extra-adds and never needs Z itself; only the final
// result is compared, after affine normalisation, so the coordinate system is free (SURVEY.md "Hard parts").
// Pasta curves have cofactor 1, so no twisted-Edwards model exists.  Formulas: EFD madd-2008-s / add-2008-s /
// dbl-2008-s-1 / mdbl-2008-s-1 with a = 0.
//
// Encodings: affine identity = (0, 0) (not on the curve); XYZZ identity = ZZ == 0 — an all-zero struct, so a
// cudaMemset-cleared bucket array is an array of identities.
#pragma once
#include "field.cuh"

namespace zkb {

struct alignas(16) affine_t {
    fe x, y;
};
struct alignas(16) xyzz_t {
    fe X, Y, ZZ, ZZZ;
};

ZK_HD bool affine_is_inf(const affine_t& p) { return fe_is_zero(p.x) && fe_is_zero(p.y); }
ZK_HD bool xyzz_is_inf(const xyzz_t& p) { return fe_is_zero(p.ZZ); }
ZK_HD xyzz_t xyzz_identity() {
    xyzz_t r;
    r.X = fe_zero(); r.Y = fe_zero(); r.ZZ = fe_zero(); r.ZZZ = fe_zero();
    return r;
}
template <class F> ZK_HD xyzz_t xyzz_from_affine(const affine_t& p) {
    xyzz_t r;
    if (affine_is_inf(p)) return xyzz_identity();
    r.X = p.x; r.Y = p.y; r.ZZ = fe_one<F>(); r.ZZZ = fe_one<F>();
    return r;
}
template <class F> ZK_HD affine_t affine_neg(const affine_t& p) {
    affine_t r;
    r.x = p.x;
    r.y = fe_neg<F>(p.y);
    return r;
}
template <class F> ZK_HD xyzz_t xyzz_neg(const xyzz_t& p) {
    xyzz_t r = p;
    r.Y = fe_neg<F>(p.Y);
    return r;
}

// 2*q for an affine q != identity (mdbl-2008-s-1, a = 0)
template <class F, bool COMPACT = false> ZK_HD xyzz_t xyzz_mdbl(const affine_t& q) {
    xyzz_t r;
    fe U = fe_dbl<F>(q.y);
    fe V = fe_mulx<F, COMPACT>(U, U);
    fe W = fe_mulx<F, COMPACT>(U, V);
    fe S = fe_mulx<F, COMPACT>(q.x, V);
    fe xx = fe_mulx<F, COMPACT>(q.x, q.x);
    fe M = fe_add<F>(fe_dbl<F>(xx), xx);
    fe X3 = fe_sub<F>(fe_mulx<F, COMPACT>(M, M), fe_dbl<F>(S));
    fe Y3 = fe_sub<F>(fe_mulx<F, COMPACT>(M, fe_sub<F>(S, X3)), fe_mulx<F, COMPACT>(W, q.y));
    r.X = X3; r.Y = Y3; r.ZZ = V; r.ZZZ = W;
    return r;
}

// 2*p (dbl-2008-s-1, a = 0).  y = 0 cannot occur on a prime-order curve, but the formula then yields ZZ = 0 anyway.
template <class F, bool COMPACT = false> ZK_HD xyzz_t xyzz_dbl(const xyzz_t& p) {
    if (xyzz_is_inf(p)) return p;
    xyzz_t r;
    fe U = fe_dbl<F>(p.Y);
    fe V = fe_mulx<F, COMPACT>(U, U);
    fe W = fe_mulx<F, COMPACT>(U, V);
    fe S = fe_mulx<F, COMPACT>(p.X, V);
    fe xx = fe_mulx<F, COMPACT>(p.X, p.X);
    fe M = fe_add<F>(fe_dbl<F>(xx), xx);
    fe X3 = fe_sub<F>(fe_mulx<F, COMPACT>(M, M), fe_dbl<F>(S));
    fe Y3 = fe_sub<F>(fe_mulx<F, COMPACT>(M, fe_sub<F>(S, X3)), fe_mulx<F, COMPACT>(W, p.Y));
    r.X = X3; r.Y = Y3;
    r.ZZ = fe_mulx<F, COMPACT>(V, p.ZZ);
    r.ZZZ = fe_mulx<F, COMPACT>(W, p.ZZZ);
    return r;
}

// p + q, q affine (madd-2008-s), all special cases handled: q or p identity, p == q (doubling), p == -q.
template <class F> ZK_HD xyzz_t xyzz_madd(const xyzz_t& p, const affine_t& q) {
    if (affine_is_inf(q)) return p;
    if (xyzz_is_inf(p)) return xyzz_from_affine<F>(q);
    fe U2 = fe_mul<F>(q.x, p.ZZ);
    fe S2 = fe_mul<F>(q.y, p.ZZZ);
    fe Pd = fe_sub<F>(U2, p.X);
    fe Rd = fe_sub<F>(S2, p.Y);
    if (fe_is_zero(Pd)) {
        if (fe_is_zero(Rd)) return xyzz_mdbl<F>(q);
        return xyzz_identity();
    }
    xyzz_t r;
    fe PP = fe_sqr<F>(Pd);
    fe PPP = fe_mul<F>(Pd, PP);
    fe Q = fe_mul<F>(p.X, PP);
    fe X3 = fe_sub<F>(fe_sub<F>(fe_sqr<F>(Rd), PPP), fe_dbl<F>(Q));
    fe Y3 = fe_sub<F>(fe_mul<F>(Rd, fe_sub<F>(Q, X3)), fe_mul<F>(p.Y, PPP));
    r.X = X3; r.Y = Y3;
    r.ZZ = fe_mul<F>(p.ZZ, PP);
    r.ZZZ = fe_mul<F>(p.ZZZ, PPP);
    return r;
}

// p + q, both XYZZ (add-2008-s), all special cases handled.
template <class F, bool COMPACT = false> ZK_HD xyzz_t xyzz_add(const xyzz_t& p, const xyzz_t& q) {
    if (xyzz_is_inf(q)) return p;
    if (xyzz_is_inf(p)) return q;
    fe U1 = fe_mulx<F, COMPACT>(p.X, q.ZZ);
    fe U2 = fe_mulx<F, COMPACT>(q.X, p.ZZ);
    fe S1 = fe_mulx<F, COMPACT>(p.Y, q.ZZZ);
    fe S2 = fe_mulx<F, COMPACT>(q.Y, p.ZZZ);
    fe Pd = fe_sub<F>(U2, U1);
    fe Rd = fe_sub<F>(S2, S1);
    if (fe_is_zero(Pd)) {
        if (fe_is_zero(Rd)) return xyzz_dbl<F, COMPACT>(p);
        return xyzz_identity();
    }
    xyzz_t r;
    fe PP = fe_mulx<F, COMPACT>(Pd, Pd);
    fe PPP = fe_mulx<F, COMPACT>(Pd, PP);
    fe Q = fe_mulx<F, COMPACT>(U1, PP);
    fe X3 = fe_sub<F>(fe_sub<F>(fe_mulx<F, COMPACT>(Rd, Rd), PPP), fe_dbl<F>(Q));
    fe Y3 = fe_sub<F>(fe_mulx<F, COMPACT>(Rd, fe_sub<F>(Q, X3)), fe_mulx<F, COMPACT>(S1, PPP));
    r.X = X3; r.Y = Y3;
    r.ZZ = fe_mulx<F, COMPACT>(fe_mulx<F, COMPACT>(p.ZZ, q.ZZ), PP);
    r.ZZZ = fe_mulx<F, COMPACT>(fe_mulx<F, COMPACT>(p.ZZZ, q.ZZZ), PPP);
    return r;
}

// [k] P, k canonical (8 x u32), MSB first.  Leading zero bits are skipped; k = 0 gives the identity.
template <class F> ZK_HD xyzz_t xyzz_scalar_mul(const xyzz_t& p, const fe& k) {
    xyzz_t acc = xyzz_identity();
    bool started = false;
    for (int i = 255; i >= 0; i--) {
        if (started) acc = xyzz_dbl<F>(acc);
        if ((k.v[i >> 5] >> (i & 31)) & 1u) {
            acc = started ? xyzz_add<F>(acc, p) : p;
            started = true;
        }
    }
    return acc;
}

// XYZZ -> affine with one field inversion: x = X / ZZ, y = Y / ZZZ.   1/ZZ = (ZZ*ZZZ)^-1 * ZZZ, 1/ZZZ = (..)^-1 * ZZ
template <class F> ZK_HD affine_t xyzz_to_affine(const xyzz_t& p) {
    affine_t r;
    if (xyzz_is_inf(p)) { r.x = fe_zero(); r.y = fe_zero(); return r; }
    fe t = fe_inv<F>(fe_mul<F>(p.ZZ, p.ZZZ));
    r.x = fe_mul<F>(p.X, fe_mul<F>(t, p.ZZZ));
    r.y = fe_mul<F>(p.Y, fe_mul<F>(t, p.ZZ));
    return r;
}

}  // namespace zkb
