/*
 * oracle/pasta_oracle.c — CPU ORACLE: TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C restatement of what the reference (o1-labs/proof-systems @ 136c1df9) computes on the CPU for
 * the MSM + NTT hot path: Pasta field arithmetic, Pallas/Vesta group law, ark-ec style signed-digit
 * Pippenger (VariableBaseMSM::msm_bigint) and ark-poly style radix-2 FFT/iFFT.  The arithmetic itself
 * lives in crates.io dependencies that are NOT in /root/reference (ark-ff / ark-ec / ark-poly 0.5.0,
 * Cargo.lock:171-279), so the algorithms are restated from their published behaviour and PINNED against the
 * reference's own vectors (tests/test_oracle_*.py):
 *   - curves/tests/pasta_curves.rs:11-74 (canonical<->Montgomery, affine-add KAT)
 *   - kimchi/src/proof.rs:1163-1204 (16-point Vesta MSM KAT)
 *   - srs/test_pallas.srs, srs/test_vesta.srs Lagrange bases (every entry is an n-point MSM answer and
 *     pins the iFFT root / ordering / 1/n scaling), srs/pallas.srs, srs/vesta.srs (compressed points)
 *
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load this.
 * The product (proof_systems_b200/) never links or calls it.
 *
 * Exported C API: field_id 0 = Fp (Pallas base / Vesta scalar), 1 = Fq (Vesta base / Pallas scalar);
 * curve_id 0 = Pallas (coordinates Fp, scalars Fq), 1 = Vesta (coordinates Fq, scalars Fp).
 * Field elements: 4 x u64 LE, Montgomery form unless a name says "canonical".  Affine points: x || y
 * (8 x u64), identity = all zero.  Jacobian: X || Y || Z (12 x u64), identity Z = 0.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------ Fp */
#define FN(x) fp_##x
#define F_MOD {0x992d30ed00000001ULL, 0x224698fc094cf91bULL, 0x0ULL, 0x4000000000000000ULL}
#define F_R {0x34786d38fffffffdULL, 0x992c350be41914adULL, 0xffffffffffffffffULL, 0x3fffffffffffffffULL}
#define F_R2 {0x8c78ecb30000000fULL, 0xd7d30dbd8b0de0e7ULL, 0x7797a99bc3c95d18ULL, 0x096d41af7b9cb714ULL}
#define F_INV 11037532056220336127ULL
#define F_ROOT {0xa28db849bad6dbf0ULL, 0x9083cd03d3b539dfULL, 0xfba6b9ca9dc8448eULL, 0x3ec928747b89c6daULL}
#include "field_impl.h"
#include "ntt_impl.h"
#undef FN
#undef F_MOD
#undef F_R
#undef F_R2
#undef F_INV
#undef F_ROOT

/* ------------------------------------------------------------------ Fq */
#define FN(x) fq_##x
#define F_MOD {0x8c46eb2100000001ULL, 0x224698fc0994a8ddULL, 0x0ULL, 0x4000000000000000ULL}
#define F_R {0x5b2b3e9cfffffffdULL, 0x992c350be3420567ULL, 0xffffffffffffffffULL, 0x3fffffffffffffffULL}
#define F_R2 {0xfc9678ff0000000fULL, 0x67bb433d891a16e3ULL, 0x7fae231004ccf590ULL, 0x096d41af7ccfdaa9ULL}
#define F_INV 10108024940646105087ULL
#define F_ROOT {0x218077428c9942deULL, 0xcc49578921b60494ULL, 0xac2e5d27b2efbee2ULL, 0x0b79fa897f2db056ULL}
#include "field_impl.h"
#include "ntt_impl.h"
#undef FN
#undef F_MOD
#undef F_R
#undef F_R2
#undef F_INV
#undef F_ROOT

/* ------------------------------------------------------------------ curves */
#define CN(x) pallas_##x
#define BF(x) fp_##x
#include "curve_impl.h"
#undef CN
#undef BF
#define CN(x) vesta_##x
#define BF(x) fq_##x
#include "curve_impl.h"
#undef CN
#undef BF

#define EXPORT __attribute__((visibility("default")))

static int default_threads(int t) {
    if (t > 0) return t;
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

EXPORT int orc_max_threads(void) { return default_threads(0); }

/* ---- field API ---- */
EXPORT void orc_fe_modulus(int fid, uint64_t out[4]) { memcpy(out, fid == 0 ? fp_MOD : fq_MOD, 32); }
EXPORT void orc_fe_mul(int fid, const uint64_t *a, const uint64_t *b, uint64_t *r) {
    if (fid == 0) fp_mul((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
    else fq_mul((fq_t *)r, (const fq_t *)a, (const fq_t *)b);
}
EXPORT void orc_fe_add(int fid, const uint64_t *a, const uint64_t *b, uint64_t *r) {
    if (fid == 0) fp_add((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
    else fq_add((fq_t *)r, (const fq_t *)a, (const fq_t *)b);
}
EXPORT void orc_fe_sub(int fid, const uint64_t *a, const uint64_t *b, uint64_t *r) {
    if (fid == 0) fp_sub((fp_t *)r, (const fp_t *)a, (const fp_t *)b);
    else fq_sub((fq_t *)r, (const fq_t *)a, (const fq_t *)b);
}
EXPORT void orc_fe_inv(int fid, const uint64_t *a, uint64_t *r) {
    if (fid == 0) fp_inv((fp_t *)r, (const fp_t *)a);
    else fq_inv((fq_t *)r, (const fq_t *)a);
}
EXPORT int orc_fe_sqrt(int fid, const uint64_t *a, uint64_t *r) {
    if (fid == 0) return fp_sqrt((fp_t *)r, (const fp_t *)a);
    return fq_sqrt((fq_t *)r, (const fq_t *)a);
}
EXPORT void orc_fe_pow(int fid, const uint64_t *a, const uint64_t *e, uint64_t *r) {
    if (fid == 0) fp_pow((fp_t *)r, (const fp_t *)a, e);
    else fq_pow((fq_t *)r, (const fq_t *)a, e);
}
/* vectorised canonical <-> Montgomery (n elements) */
EXPORT void orc_fe_to_mont(int fid, const uint64_t *a, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (fid == 0) fp_to_mont((fp_t *)(r + 4 * i), a + 4 * i);
        else fq_to_mont((fq_t *)(r + 4 * i), a + 4 * i);
    }
}
EXPORT void orc_fe_from_mont(int fid, const uint64_t *a, uint64_t *r, size_t n) {
    for (size_t i = 0; i < n; i++) {
        if (fid == 0) fp_from_mont(r + 4 * i, (const fp_t *)(a + 4 * i));
        else fq_from_mont(r + 4 * i, (const fq_t *)(a + 4 * i));
    }
}
EXPORT void orc_fe_root_of_unity(int fid, unsigned log_n, uint64_t *r) {
    if (fid == 0) fp_root_of_unity((fp_t *)r, log_n);
    else fq_root_of_unity((fq_t *)r, log_n);
}
EXPORT void orc_fe_two_adic_root(int fid, uint64_t *r) { memcpy(r, fid == 0 ? fp_ROOT : fq_ROOT, 32); }
EXPORT void orc_fe_r2(int fid, uint64_t *r) { memcpy(r, fid == 0 ? fp_R2 : fq_R2, 32); }
EXPORT void orc_fe_one(int fid, uint64_t *r) { memcpy(r, fid == 0 ? fp_R1 : fq_R1, 32); }

/* ---- NTT API ---- */
EXPORT void orc_ntt(int fid, uint64_t *data, unsigned log_n, int inverse, int coset, int threads) {
    threads = default_threads(threads);
    if (fid == 0) fp_ntt((fp_t *)data, log_n, inverse, coset, threads);
    else fq_ntt((fq_t *)data, log_n, inverse, coset, threads);
}
EXPORT void orc_dft_naive(int fid, const uint64_t *in, uint64_t *out, unsigned log_n, int inverse) {
    if (fid == 0) fp_dft_naive((fp_t *)out, (const fp_t *)in, log_n, inverse);
    else fq_dft_naive((fq_t *)out, (const fq_t *)in, log_n, inverse);
}

/* permutation part of the quotient over d8 (kimchi/src/circuits/polynomials/permutation.rs:223-357); see ntt_impl.h */
EXPORT void orc_perm_quot(int fid, const uint64_t *w, size_t w_stride, const uint64_t *z, const uint64_t *sigma, size_t sigma_stride,
                          const uint64_t *zkpm, const uint64_t *beta, const uint64_t *gamma, const uint64_t *alpha0, const uint64_t *shift,
                          unsigned next_shift, unsigned log_m, uint64_t *out, int threads) {
    threads = default_threads(threads);
    if (fid == 0)
        fp_perm_quot((const fp_t *)w, w_stride, (const fp_t *)z, (const fp_t *)sigma, sigma_stride, (const fp_t *)zkpm, (const fp_t *)beta,
                     (const fp_t *)gamma, (const fp_t *)alpha0, (const fp_t *)shift, next_shift, log_m, (fp_t *)out, threads);
    else
        fq_perm_quot((const fq_t *)w, w_stride, (const fq_t *)z, (const fq_t *)sigma, sigma_stride, (const fq_t *)zkpm, (const fq_t *)beta,
                     (const fq_t *)gamma, (const fq_t *)alpha0, (const fq_t *)shift, next_shift, log_m, (fq_t *)out, threads);
}

/* RPN constraint evaluation over a whole domain (kimchi/src/circuits/expr.rs:856-940, :1938-1990); see ntt_impl.h */
EXPORT int orc_expr_eval(int fid, const uint32_t *ops, const uint32_t *args, size_t n_tok, const uint64_t *literals, size_t n_lit,
                         const uint64_t *const *col_evals, const uint64_t *col_len, const uint32_t *col_mult, size_t n_cols, uint64_t out_len,
                         int accumulate, uint64_t *out, int threads) {
    threads = default_threads(threads);
    if (fid == 0)
        return fp_expr_eval(ops, args, n_tok, (const fp_t *)literals, n_lit, (const fp_t *const *)col_evals, col_len, col_mult, n_cols, out_len,
                            accumulate, (fp_t *)out, threads);
    return fq_expr_eval(ops, args, n_tok, (const fq_t *)literals, n_lit, (const fq_t *const *)col_evals, col_len, col_mult, n_cols, out_len,
                        accumulate, (fq_t *)out, threads);
}

/* DensePolynomial::divide_by_vanishing_poly over d1 (kimchi/src/prover.rs:909); see ntt_impl.h */
EXPORT void orc_divide_by_vanishing(int fid, const uint64_t *f, size_t len, unsigned log_n, uint64_t *quot, uint64_t *rem) {
    if (fid == 0) fp_divide_by_vanishing((const fp_t *)f, len, (size_t)1 << log_n, (fp_t *)quot, (fp_t *)rem);
    else fq_divide_by_vanishing((const fq_t *)f, len, (size_t)1 << log_n, (fq_t *)quot, (fq_t *)rem);
}

/* ---- curve API ---- */
EXPORT int orc_on_curve(int cid, const uint64_t *xy) {
    if (cid == 0) return pallas_aff_on_curve((const pallas_aff *)xy);
    return vesta_aff_on_curve((const vesta_aff *)xy);
}
EXPORT void orc_affine_add(int cid, const uint64_t *p, const uint64_t *q, uint64_t *r) {
    if (cid == 0) {
        pallas_jac j; pallas_jac_from_aff(&j, (const pallas_aff *)p);
        pallas_jac_add_mixed(&j, &j, (const pallas_aff *)q);
        pallas_jac_to_aff((pallas_aff *)r, &j);
    } else {
        vesta_jac j; vesta_jac_from_aff(&j, (const vesta_aff *)p);
        vesta_jac_add_mixed(&j, &j, (const vesta_aff *)q);
        vesta_jac_to_aff((vesta_aff *)r, &j);
    }
}
EXPORT void orc_jac_add(int cid, const uint64_t *p, const uint64_t *q, uint64_t *r) {
    if (cid == 0) pallas_jac_add((pallas_jac *)r, (const pallas_jac *)p, (const pallas_jac *)q);
    else vesta_jac_add((vesta_jac *)r, (const vesta_jac *)p, (const vesta_jac *)q);
}
EXPORT void orc_jac_to_affine(int cid, const uint64_t *p, uint64_t *r) {
    if (cid == 0) pallas_jac_to_aff((pallas_aff *)r, (const pallas_jac *)p);
    else vesta_jac_to_aff((vesta_aff *)r, (const vesta_jac *)p);
}
/* k canonical; result affine */
EXPORT void orc_scalar_mul(int cid, const uint64_t *p, const uint64_t *k, uint64_t *r) {
    if (cid == 0) {
        pallas_jac j; pallas_scalar_mul(&j, (const pallas_aff *)p, k);
        pallas_jac_to_aff((pallas_aff *)r, &j);
    } else {
        vesta_jac j; vesta_scalar_mul(&j, (const vesta_aff *)p, k);
        vesta_jac_to_aff((vesta_aff *)r, &j);
    }
}
/* n compressed points (33 B each) -> n affine points; returns number decoded OK (== n on success) */
EXPORT size_t orc_decompress(int cid, const uint8_t *in, uint64_t *out, size_t n, int threads) {
    threads = default_threads(threads);
    size_t ok = 0;
#pragma omp parallel for num_threads(threads) reduction(+ : ok)
    for (size_t i = 0; i < n; i++) {
        int r = cid == 0 ? pallas_decompress((pallas_aff *)(out + 8 * i), in + 33 * i)
                         : vesta_decompress((vesta_aff *)(out + 8 * i), in + 33 * i);
        ok += (size_t)r;
    }
    return ok;
}

EXPORT void orc_extend_bases(int cid, const uint64_t *in, size_t m, uint64_t *out, size_t n) {
    if (cid == 0) pallas_extend_bases((pallas_aff *)out, (const pallas_aff *)in, m, n);
    else vesta_extend_bases((vesta_aff *)out, (const vesta_aff *)in, m, n);
}

/* MSM == G::Group::msm_bigint(bases, scalars): scalars canonical.  out_jac: 12 u64 (may be NULL),
 * out_aff: 8 u64 (may be NULL).  algo 0 = Pippenger (ark-like), 1 = naive definition. */
EXPORT void orc_msm(int cid, const uint64_t *bases, const uint64_t *scalars, size_t n, int algo, int threads,
                    uint64_t *out_jac, uint64_t *out_aff) {
    threads = default_threads(threads);
    if (cid == 0) {
        pallas_jac r;
        if (algo == 0) pallas_msm_pippenger(&r, (const pallas_aff *)bases, scalars, n, threads);
        else pallas_msm_naive(&r, (const pallas_aff *)bases, scalars, n);
        if (out_jac) memcpy(out_jac, &r, sizeof r);
        if (out_aff) pallas_jac_to_aff((pallas_aff *)out_aff, &r);
    } else {
        vesta_jac r;
        if (algo == 0) vesta_msm_pippenger(&r, (const vesta_aff *)bases, scalars, n, threads);
        else vesta_msm_naive(&r, (const vesta_aff *)bases, scalars, n);
        if (out_jac) memcpy(out_jac, &r, sizeof r);
        if (out_aff) vesta_jac_to_aff((vesta_aff *)out_aff, &r);
    }
}

/* The reference's commit_non_hiding for a polynomial of exactly |g| coefficients (poly-commitment/src/ipa.rs:652-662):
 * rayon::join of two half-size MSMs, then one addition — the "vertical" split benchmarked in benches/msm.rs:71-88.
 * The 2 x windows tasks run as one flat list (rayon's work stealing), see msm_pippenger_parts. */
EXPORT void orc_msm_split2(int cid, const uint64_t *bases, const uint64_t *scalars, size_t n, int threads, uint64_t *out_aff) {
    threads = default_threads(threads);
    if (cid == 0) {
        pallas_jac r; pallas_msm_pippenger_parts(&r, (const pallas_aff *)bases, scalars, n, 2, threads);
        pallas_jac_to_aff((pallas_aff *)out_aff, &r);
    } else {
        vesta_jac r; vesta_msm_pippenger_parts(&r, (const vesta_aff *)bases, scalars, n, 2, threads);
        vesta_jac_to_aff((vesta_aff *)out_aff, &r);
    }
}

/* == G::Group::msm(bases, scalars): scalars in Montgomery form (into_bigint first, like
 * VariableBaseMSM::msm at poly-commitment/src/ipa.rs:649,658). */
EXPORT void orc_msm_mont(int cid, const uint64_t *bases, const uint64_t *scalars_mont, size_t n, int threads,
                         uint64_t *out_aff) {
    uint64_t *can = (uint64_t *)malloc(n * 32 + 32);
    /* scalar field of Pallas is Fq (fid 1), of Vesta is Fp (fid 0) */
    orc_fe_from_mont(cid == 0 ? 1 : 0, scalars_mont, can, n);
    orc_msm(cid, bases, can, n, 0, threads, NULL, out_aff);
    free(can);
}

/*
 * Group iFFT of n = 2^log_n points, then batch-normalise: the body of SRS::lagrange_basis for one chunk
 * (poly-commitment/src/ipa.rs:1152-1164).  in: n affine points (g[0..n]); out: n affine points.
 * Scalars live in the curve's scalar field.  Used to pin the oracle's NTT conventions against srs/test_*.srs.
 */
#define GROUP_INTT(CNAME, SF)                                                                              \
    static void CNAME##_group_intt(CNAME##_aff *out, const CNAME##_aff *in, unsigned log_n, int threads) { \
        size_t n = (size_t)1 << log_n;                                                                     \
        CNAME##_jac *a = (CNAME##_jac *)malloc(n * sizeof(CNAME##_jac));                                   \
        for (size_t i = 0; i < n; i++) CNAME##_jac_from_aff(&a[i], &in[i]);                                \
        SF##_t w; SF##_root_of_unity(&w, log_n);                                                           \
        SF##_t wi; SF##_inv(&wi, &w);                                                                      \
        /* DIF with w^-1: natural in, bit-reversed out */                                                  \
        for (unsigned s = 0; s < log_n; s++) {                                                             \
            size_t half = n >> (s + 1);                                                                    \
            SF##_t wstep; SF##_pow_u64(&wstep, &wi, (uint64_t)1 << s);                                     \
            _Pragma("omp parallel for num_threads(threads) schedule(static)")                              \
            for (size_t k = 0; k < n / 2; k++) {                                                           \
                size_t grp = k / half, j = k % half;                                                       \
                size_t i0 = grp * 2 * half + j, i1 = i0 + half;                                            \
                CNAME##_jac u = a[i0], v = a[i1], nv, d;                                                   \
                CNAME##_jac_add(&a[i0], &u, &v);                                                           \
                CNAME##_jac_neg(&nv, &v);                                                                  \
                CNAME##_jac_add(&d, &u, &nv);                                                              \
                SF##_t tw; SF##_pow_u64(&tw, &wstep, j);                                                   \
                uint64_t twc[4]; SF##_from_mont(twc, &tw);                                                 \
                CNAME##_jac_scalar_mul(&a[i1], &d, twc);                                                   \
            }                                                                                              \
        }                                                                                                  \
        /* scale by 1/n */                                                                                 \
        SF##_t nn, ninv; uint64_t nc[4] = {(uint64_t)n, 0, 0, 0};                                          \
        SF##_to_mont(&nn, nc); SF##_inv(&ninv, &nn);                                                       \
        uint64_t ninvc[4]; SF##_from_mont(ninvc, &ninv);                                                   \
        _Pragma("omp parallel for num_threads(threads) schedule(static)")                                  \
        for (size_t i = 0; i < n; i++) CNAME##_jac_scalar_mul(&a[i], &a[i], ninvc);                        \
        /* bit-reverse */                                                                                  \
        for (size_t i = 0; i < n; i++) {                                                                   \
            size_t r = 0;                                                                                  \
            for (unsigned b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);                   \
            if (i < r) { CNAME##_jac t = a[i]; a[i] = a[r]; a[r] = t; }                                    \
        }                                                                                                  \
        CNAME##_jac_batch_to_aff(out, a, n);                                                               \
        free(a);                                                                                           \
    }
GROUP_INTT(pallas, fq)
GROUP_INTT(vesta, fp)

EXPORT void orc_group_intt(int cid, const uint64_t *in, uint64_t *out, unsigned log_n, int threads) {
    threads = default_threads(threads);
    if (cid == 0) pallas_group_intt((pallas_aff *)out, (const pallas_aff *)in, log_n, threads);
    else vesta_group_intt((vesta_aff *)out, (const vesta_aff *)in, log_n, threads);
}
