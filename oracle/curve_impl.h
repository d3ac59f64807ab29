/*
 * oracle/curve_impl.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Short-Weierstrass curve y^2 = x^3 + 5 (a = 0, b = 5, cofactor 1) over a Pasta base field.
 * Included twice by pasta_oracle.c:
 *   CN(name) -> pallas_name / vesta_name
 *   BF(name) -> base-field functions (fp_ for Pallas, fq_ for Vesta)
 * Reference constants: curves/src/pasta/curves/pallas.rs:8-46, vesta.rs:8-45.
 *
 * Representation restated from arkworks (ark-ec 0.5.0 short_weierstrass, pinned in Cargo.lock:171-279,
 * source not on disk): Affine{x, y, infinity}; Projective = Jacobian (X:Y:Z), x = X/Z^2, y = Y/Z^3,
 * Z = 0 is the identity.  Here an affine identity is encoded as (x, y) = (0, 0), which is not on the
 * curve (0 != 5), so the encoding is unambiguous.
 * Formulas are the public EFD ones (dbl-2009-l, madd-2007-bl, add-2007-bl) for a = 0.
 */

typedef struct { BF(t) x, y; } CN(aff);
typedef struct { BF(t) X, Y, Z; } CN(jac);

static inline int CN(aff_is_inf)(const CN(aff) *p) { return BF(is_zero)(&p->x) && BF(is_zero)(&p->y); }
static inline void CN(aff_set_inf)(CN(aff) *p) { BF(set_zero)(&p->x); BF(set_zero)(&p->y); }
static inline int CN(jac_is_inf)(const CN(jac) *p) { return BF(is_zero)(&p->Z); }
static inline void CN(jac_set_inf)(CN(jac) *p) { BF(set_one)(&p->X); BF(set_one)(&p->Y); BF(set_zero)(&p->Z); }
static inline void CN(jac_from_aff)(CN(jac) *r, const CN(aff) *p) {
    if (CN(aff_is_inf)(p)) { CN(jac_set_inf)(r); return; }
    r->X = p->x; r->Y = p->y; BF(set_one)(&r->Z);
}

static int CN(aff_on_curve)(const CN(aff) *p) {
    if (CN(aff_is_inf)(p)) return 1;
    BF(t) y2, x3, five;
    uint64_t c5[4] = {5, 0, 0, 0};
    BF(to_mont)(&five, c5);
    BF(sqr)(&y2, &p->y);
    BF(sqr)(&x3, &p->x); BF(mul)(&x3, &x3, &p->x);
    BF(add)(&x3, &x3, &five);
    return BF(eq)(&y2, &x3);
}

static void CN(jac_double)(CN(jac) *r, const CN(jac) *p) {
    if (CN(jac_is_inf)(p)) { *r = *p; return; }
    BF(t) A, B, C, D, E, F, t, X3, Y3, Z3;
    BF(sqr)(&A, &p->X);
    BF(sqr)(&B, &p->Y);
    BF(sqr)(&C, &B);
    BF(add)(&t, &p->X, &B); BF(sqr)(&t, &t); BF(sub)(&t, &t, &A); BF(sub)(&t, &t, &C);
    BF(dbl)(&D, &t);
    BF(dbl)(&E, &A); BF(add)(&E, &E, &A);
    BF(sqr)(&F, &E);
    BF(dbl)(&t, &D); BF(sub)(&X3, &F, &t);
    BF(sub)(&t, &D, &X3); BF(mul)(&Y3, &E, &t);
    BF(dbl)(&t, &C); BF(dbl)(&t, &t); BF(dbl)(&t, &t);
    BF(sub)(&Y3, &Y3, &t);
    BF(mul)(&Z3, &p->Y, &p->Z); BF(dbl)(&Z3, &Z3);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}

/* r = p + q, q affine (mixed) */
static void CN(jac_add_mixed)(CN(jac) *r, const CN(jac) *p, const CN(aff) *q) {
    if (CN(aff_is_inf)(q)) { *r = *p; return; }
    if (CN(jac_is_inf)(p)) { CN(jac_from_aff)(r, q); return; }
    BF(t) Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;
    BF(sqr)(&Z1Z1, &p->Z);
    BF(mul)(&U2, &q->x, &Z1Z1);
    BF(mul)(&S2, &q->y, &p->Z); BF(mul)(&S2, &S2, &Z1Z1);
    BF(sub)(&H, &U2, &p->X);
    BF(sub)(&rr, &S2, &p->Y);
    if (BF(is_zero)(&H)) {
        if (BF(is_zero)(&rr)) { CN(jac_double)(r, p); return; }
        CN(jac_set_inf)(r); return;
    }
    BF(dbl)(&rr, &rr);
    BF(sqr)(&HH, &H);
    BF(dbl)(&I, &HH); BF(dbl)(&I, &I);
    BF(mul)(&J, &H, &I);
    BF(mul)(&V, &p->X, &I);
    BF(sqr)(&X3, &rr); BF(sub)(&X3, &X3, &J); BF(dbl)(&t, &V); BF(sub)(&X3, &X3, &t);
    BF(sub)(&t, &V, &X3); BF(mul)(&Y3, &rr, &t);
    BF(mul)(&t, &p->Y, &J); BF(dbl)(&t, &t); BF(sub)(&Y3, &Y3, &t);
    BF(add)(&Z3, &p->Z, &H); BF(sqr)(&Z3, &Z3); BF(sub)(&Z3, &Z3, &Z1Z1); BF(sub)(&Z3, &Z3, &HH);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}

static void CN(jac_add)(CN(jac) *r, const CN(jac) *p, const CN(jac) *q) {
    if (CN(jac_is_inf)(q)) { *r = *p; return; }
    if (CN(jac_is_inf)(p)) { *r = *q; return; }
    BF(t) Z1Z1, Z2Z2, U1, U2, S1, S2, H, I, J, rr, V, t, X3, Y3, Z3;
    BF(sqr)(&Z1Z1, &p->Z);
    BF(sqr)(&Z2Z2, &q->Z);
    BF(mul)(&U1, &p->X, &Z2Z2);
    BF(mul)(&U2, &q->X, &Z1Z1);
    BF(mul)(&S1, &p->Y, &q->Z); BF(mul)(&S1, &S1, &Z2Z2);
    BF(mul)(&S2, &q->Y, &p->Z); BF(mul)(&S2, &S2, &Z1Z1);
    BF(sub)(&H, &U2, &U1);
    BF(sub)(&rr, &S2, &S1);
    if (BF(is_zero)(&H)) {
        if (BF(is_zero)(&rr)) { CN(jac_double)(r, p); return; }
        CN(jac_set_inf)(r); return;
    }
    BF(dbl)(&rr, &rr);
    BF(dbl)(&I, &H); BF(sqr)(&I, &I);
    BF(mul)(&J, &H, &I);
    BF(mul)(&V, &U1, &I);
    BF(sqr)(&X3, &rr); BF(sub)(&X3, &X3, &J); BF(dbl)(&t, &V); BF(sub)(&X3, &X3, &t);
    BF(sub)(&t, &V, &X3); BF(mul)(&Y3, &rr, &t);
    BF(mul)(&t, &S1, &J); BF(dbl)(&t, &t); BF(sub)(&Y3, &Y3, &t);
    BF(add)(&Z3, &p->Z, &q->Z); BF(sqr)(&Z3, &Z3); BF(sub)(&Z3, &Z3, &Z1Z1); BF(sub)(&Z3, &Z3, &Z2Z2);
    BF(mul)(&Z3, &Z3, &H);
    r->X = X3; r->Y = Y3; r->Z = Z3;
}

static void CN(jac_neg)(CN(jac) *r, const CN(jac) *p) { r->X = p->X; r->Z = p->Z; BF(neg)(&r->Y, &p->Y); }
static void CN(aff_neg)(CN(aff) *r, const CN(aff) *p) { r->x = p->x; BF(neg)(&r->y, &p->y); }

/* ark: into_affine — x = X/Z^2, y = Y/Z^3 */
static void CN(jac_to_aff)(CN(aff) *r, const CN(jac) *p) {
    if (CN(jac_is_inf)(p)) { CN(aff_set_inf)(r); return; }
    BF(t) zi, zi2, zi3;
    BF(inv)(&zi, &p->Z);
    BF(sqr)(&zi2, &zi);
    BF(mul)(&zi3, &zi2, &zi);
    BF(mul)(&r->x, &p->X, &zi2);
    BF(mul)(&r->y, &p->Y, &zi3);
}

/* ark: Group::normalize_batch */
static void CN(jac_batch_to_aff)(CN(aff) *out, const CN(jac) *in, size_t n) {
    BF(t) *z = (BF(t) *)malloc(n * sizeof(BF(t)));
    for (size_t i = 0; i < n; i++) z[i] = in[i].Z;
    BF(batch_inv)(z, n);
    for (size_t i = 0; i < n; i++) {
        if (CN(jac_is_inf)(&in[i])) { CN(aff_set_inf)(&out[i]); continue; }
        BF(t) zi2, zi3;
        BF(sqr)(&zi2, &z[i]);
        BF(mul)(&zi3, &zi2, &z[i]);
        BF(mul)(&out[i].x, &in[i].X, &zi2);
        BF(mul)(&out[i].y, &in[i].Y, &zi3);
    }
    free(z);
}

/* [k]P by MSB-first double-and-add; k canonical 4 x u64 LE */
static void CN(scalar_mul)(CN(jac) *r, const CN(aff) *p, const uint64_t k[4]) {
    CN(jac) acc; CN(jac_set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        CN(jac_double)(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) CN(jac_add_mixed)(&acc, &acc, p);
    }
    *r = acc;
}
static void CN(jac_scalar_mul)(CN(jac) *r, const CN(jac) *p, const uint64_t k[4]) {
    CN(jac) acc; CN(jac_set_inf)(&acc);
    for (int i = 255; i >= 0; i--) {
        CN(jac_double)(&acc, &acc);
        if ((k[i / 64] >> (i % 64)) & 1) CN(jac_add)(&acc, &acc, p);
    }
    *r = acc;
}

/* Definition-level MSM: sum_i [s_i] P_i one scalar multiplication at a time. */
static void CN(msm_naive)(CN(jac) *r, const CN(aff) *bases, const uint64_t *scalars, size_t n) {
    CN(jac) acc; CN(jac_set_inf)(&acc);
    for (size_t i = 0; i < n; i++) {
        CN(jac) t;
        CN(scalar_mul)(&t, &bases[i], scalars + 4 * i);
        CN(jac_add)(&acc, &acc, &t);
    }
    *r = acc;
}

static inline size_t CN(ceil_log2)(size_t x) {
    size_t l = 0;
    while (((size_t)1 << l) < x) l++;
    return l;
}

/*
 * Restatement of ark-ec 0.5.0 VariableBaseMSM::msm_bigint for groups with cheap negation (the call the
 * reference makes at poly-commitment/src/ipa.rs:672,943,953 and commitment.rs:382,387):
 *   window c = 3 if n < 32 else ceil_log2(n)*69/100 + 2; ceil(255/c) signed digits per scalar in
 *   [-2^(c-1), 2^(c-1)); one bucket array per window, bucket[|d|-1] +/-= base (mixed addition), zero digits
 *   skipped; running-sum bucket reduction; windows in parallel; Horner combine with c doublings per window.
 * (Algorithm recalled from the published crate — SURVEY.md Appendix C; only the resulting group element
 * is pinned by the reference's vectors.)  Lengths: uses min(len(bases), len(scalars)) like msm_bigint.
 */
static void CN(msm_pippenger_parts)(CN(jac) *r, const CN(aff) *bases, const uint64_t *scalars, size_t n, unsigned parts, int threads) {
    /* parts = 1: one msm_bigint call.  parts = 2: the reference's rayon::join of two half-size calls followed by one
     * addition (poly-commitment/src/ipa.rs:652-662).  All (part, window) pairs form ONE flat task list — what rayon's
     * work stealing makes of the nested parallelism — so no nested OpenMP teams are created. */
    if (n == 0) { CN(jac_set_inf)(r); return; }
    if (parts < 1) parts = 1;
    if (parts > n) parts = (unsigned)n;
    size_t pn_max = (n + parts - 1) / parts;
    unsigned c = pn_max < 32 ? 3 : (unsigned)(CN(ceil_log2)(pn_max) * 69 / 100 + 2);   /* ark: ln_without_floats(size) + 2 */
    const unsigned num_bits = 255;
    unsigned nwin = (num_bits + c - 1) / c;
    int32_t *digits = (int32_t *)malloc((size_t)n * nwin * sizeof(int32_t));
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t i = 0; i < n; i++) {
        const uint64_t *s = scalars + 4 * i;
        uint64_t carry = 0;
        for (unsigned w = 0; w < nwin; w++) {
            unsigned off = w * c;
            unsigned limb = off / 64, sh = off % 64;
            uint64_t bits = s[limb] >> sh;
            if (sh + c > 64 && limb + 1 < 4) bits |= s[limb + 1] << (64 - sh);
            uint64_t coef = carry + (bits & (((uint64_t)1 << c) - 1));
            carry = (coef + ((uint64_t)1 << (c - 1))) >> c;
            int64_t d = (int64_t)coef - (int64_t)(carry << c);
            if (w == nwin - 1) d += (int64_t)(carry << c);
            digits[i * nwin + w] = (int32_t)d;
        }
    }
    CN(jac) *wsum = (CN(jac) *)malloc((size_t)parts * nwin * sizeof(CN(jac)));
    size_t nb = (size_t)1 << (c - 1);
    /* the last window can hold an un-recoded digit up to 2^c */
    size_t nb_last = (size_t)1 << c;
    int tasks = (int)(parts * nwin);
    int thr = threads < tasks ? threads : tasks;
#pragma omp parallel for num_threads(thr) schedule(dynamic, 1)
    for (int t = 0; t < tasks; t++) {
        unsigned p = (unsigned)t / nwin, w = (unsigned)t % nwin;
        size_t lo = n * p / parts, hi = n * (p + 1) / parts;
        size_t nbw = (w == nwin - 1) ? nb_last : nb;
        CN(jac) *bk = (CN(jac) *)malloc(nbw * sizeof(CN(jac)));
        for (size_t b = 0; b < nbw; b++) CN(jac_set_inf)(&bk[b]);
        for (size_t i = lo; i < hi; i++) {
            int32_t d = digits[i * nwin + w];
            if (d > 0) CN(jac_add_mixed)(&bk[d - 1], &bk[d - 1], &bases[i]);
            else if (d < 0) { CN(aff) nq; CN(aff_neg)(&nq, &bases[i]); CN(jac_add_mixed)(&bk[-d - 1], &bk[-d - 1], &nq); }
        }
        CN(jac) run, tot; CN(jac_set_inf)(&run); CN(jac_set_inf)(&tot);
        for (size_t b = nbw; b-- > 0;) {
            CN(jac_add)(&run, &run, &bk[b]);
            CN(jac_add)(&tot, &tot, &run);
        }
        wsum[t] = tot;
        free(bk);
    }
    CN(jac) total; CN(jac_set_inf)(&total);
    for (unsigned p = 0; p < parts; p++) {
        CN(jac) acc = wsum[p * nwin + nwin - 1];
        for (int w = (int)nwin - 2; w >= 0; w--) {
            for (unsigned k = 0; k < c; k++) CN(jac_double)(&acc, &acc);
            CN(jac_add)(&acc, &acc, &wsum[p * nwin + w]);
        }
        CN(jac_add)(&total, &total, &acc);
    }
    *r = total;
    free(wsum);
    free(digits);
}

static void CN(msm_pippenger)(CN(jac) *r, const CN(aff) *bases, const uint64_t *scalars, size_t n, int threads) {
    CN(msm_pippenger_parts)(r, bases, scalars, n, 1, threads);
}

/*
 * Point decompression, ark-serialize compressed SW form as used by SerdeAs
 * (utils/src/serialization.rs:65-106): 32-byte LE canonical x, then one flag byte:
 * bit 7 = "y is the larger of {y, p-y}", bit 6 = infinity.  Returns 0 on failure.
 */
static int CN(decompress)(CN(aff) *r, const uint8_t in[33]) {
    uint8_t flags = in[32];
    uint64_t xc[4];
    memcpy(xc, in, 32);
    /* ark-serialize: SWFlags::from_u8 knows only 0x00 / 0x80 / 0x40; the field element must be canonical (x < modulus) */
    if ((flags & 0x3f) != 0 || (flags & 0xc0) == 0xc0) return 0;
    if (BF(geq_mod)(xc)) return 0;
    if (flags & 0x40) { CN(aff_set_inf)(r); return 1; }
    BF(t) x, rhs, y, five, ny;
    uint64_t c5[4] = {5, 0, 0, 0};
    BF(to_mont)(&x, xc);
    BF(to_mont)(&five, c5);
    BF(sqr)(&rhs, &x); BF(mul)(&rhs, &rhs, &x); BF(add)(&rhs, &rhs, &five);
    if (!BF(sqrt)(&y, &rhs)) return 0;
    BF(neg)(&ny, &y);
    uint64_t yc[4], nyc[4];
    BF(from_mont)(yc, &y);
    BF(from_mont)(nyc, &ny);
    int y_is_larger = 0;
    for (int i = 3; i >= 0; i--) {
        if (yc[i] > nyc[i]) { y_is_larger = 1; break; }
        if (yc[i] < nyc[i]) { y_is_larger = 0; break; }
    }
    int want_larger = (flags & 0x80) != 0;
    r->x = x;
    r->y = (want_larger == y_is_larger) ? y : ny;
    return 1;
}

/*
 * Deterministic extension of a base set for the synthetic large-MSM configs ("2^20 bases =
 * srs/test_vesta.srs g[..] extended deterministically by batch-affine additions of earlier entries"):
 *   out[0..m) = in;   out[k*m + i] = out[(k-1)*m + i] + in[(i + k) % m]        (one shared inversion per block of m)
 * Any on-curve points are valid MSM bases; the oracle and the GPU run on the identical array.
 */
static void CN(extend_bases)(CN(aff) *out, const CN(aff) *in, size_t m, size_t n) {
    memcpy(out, in, (m < n ? m : n) * sizeof(CN(aff)));
    BF(t) *den = (BF(t) *)malloc(m * sizeof(BF(t)));
    for (size_t k = 1; k * m < n; k++) {
        size_t cnt = (k + 1) * m <= n ? m : n - k * m;
        const CN(aff) *prev = out + (k - 1) * m;
        CN(aff) *cur = out + k * m;
        for (size_t i = 0; i < cnt; i++) BF(sub)(&den[i], &in[(i + k) % m].x, &prev[i].x);   /* x2 - x1 */
        BF(batch_inv)(den, cnt);
        for (size_t i = 0; i < cnt; i++) {
            const CN(aff) *p = &prev[i], *q = &in[(i + k) % m];
            if (BF(is_zero)(&den[i]) || CN(aff_is_inf)(p) || CN(aff_is_inf)(q)) {   /* same x or identity: generic law */
                CN(jac) j; CN(jac_from_aff)(&j, p); CN(jac_add_mixed)(&j, &j, q); CN(jac_to_aff)(&cur[i], &j);
                continue;
            }
            BF(t) lam, t, x3, y3;
            BF(sub)(&t, &q->y, &p->y); BF(mul)(&lam, &t, &den[i]);
            BF(sqr)(&x3, &lam); BF(sub)(&x3, &x3, &p->x); BF(sub)(&x3, &x3, &q->x);
            BF(sub)(&t, &p->x, &x3); BF(mul)(&y3, &lam, &t); BF(sub)(&y3, &y3, &p->y);
            cur[i].x = x3; cur[i].y = y3;
        }
    }
    free(den);
}
