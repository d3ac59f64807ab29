/*
 * oracle/field_impl.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * 255-bit prime-field arithmetic, 4 x u64 little-endian limbs, Montgomery form with
 * R = 2^256.  Included twice by pasta_oracle.c, once per Pasta field, with
 *   FN(name)  -> name mangling (fp_name / fq_name)
 *   F_MOD     -> modulus limbs            (reference: curves/src/pasta/fields/fp.rs:31-36, fq.rs:29-34)
 *   F_R, F_R2 -> R mod m, R^2 mod m       (fp.rs:38-50, fq.rs:36-48)
 *   F_INV     -> -m^{-1} mod 2^64         (fp.rs:80, fq.rs:79)
 *   F_ROOT    -> 2^32-th root of unity, Montgomery form (fp.rs:24-27, fq.rs:21-24)
 *
 * The in-memory element convention (4 x u64 LE Montgomery limbs) is the one arkworks'
 * Fp256<MontBackend<_,4>> uses and the one the reference itself relies on when it casts raw
 * limbs (kimchi/src/cached_prover_index.rs:486-530).
 */

typedef struct { uint64_t l[4]; } FN(t);

static const uint64_t FN(MOD)[4] = F_MOD;
static const uint64_t FN(R1)[4] = F_R;
static const uint64_t FN(R2)[4] = F_R2;
static const uint64_t FN(ROOT)[4] = F_ROOT;

static inline int FN(is_zero)(const FN(t) *a) { return (a->l[0] | a->l[1] | a->l[2] | a->l[3]) == 0; }
static inline int FN(eq)(const FN(t) *a, const FN(t) *b) {
    return ((a->l[0] ^ b->l[0]) | (a->l[1] ^ b->l[1]) | (a->l[2] ^ b->l[2]) | (a->l[3] ^ b->l[3])) == 0;
}
static inline void FN(set_zero)(FN(t) *a) { a->l[0] = a->l[1] = a->l[2] = a->l[3] = 0; }
static inline void FN(set_one)(FN(t) *a) { memcpy(a->l, FN(R1), 32); }

/* a >= m ? */
static inline int FN(geq_mod)(const uint64_t a[4]) {
    for (int i = 3; i >= 0; i--) {
        if (a[i] > FN(MOD)[i]) return 1;
        if (a[i] < FN(MOD)[i]) return 0;
    }
    return 1;
}
static inline void FN(sub_mod_raw)(uint64_t a[4]) {
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)a[i] - FN(MOD)[i] - (uint64_t)br;
        a[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
}

static inline void FN(add)(FN(t) *r, const FN(t) *a, const FN(t) *b) {
    unsigned __int128 c = 0;
    uint64_t t[4];
    for (int i = 0; i < 4; i++) {
        c += (unsigned __int128)a->l[i] + b->l[i];
        t[i] = (uint64_t)c;
        c >>= 64;
    }
    /* m < 2^255 so a+b < 2^256: no carry out */
    if (FN(geq_mod)(t)) FN(sub_mod_raw)(t);
    memcpy(r->l, t, 32);
}
static inline void FN(sub)(FN(t) *r, const FN(t) *a, const FN(t) *b) {
    uint64_t t[4];
    unsigned __int128 br = 0;
    for (int i = 0; i < 4; i++) {
        unsigned __int128 d = (unsigned __int128)a->l[i] - b->l[i] - (uint64_t)br;
        t[i] = (uint64_t)d;
        br = (d >> 64) & 1;
    }
    if (br) {
        unsigned __int128 c = 0;
        for (int i = 0; i < 4; i++) {
            c += (unsigned __int128)t[i] + FN(MOD)[i];
            t[i] = (uint64_t)c;
            c >>= 64;
        }
    }
    memcpy(r->l, t, 32);
}
static inline void FN(neg)(FN(t) *r, const FN(t) *a) {
    FN(t) z; FN(set_zero)(&z);
    FN(sub)(r, &z, a);
}
static inline void FN(dbl)(FN(t) *r, const FN(t) *a) { FN(add)(r, a, a); }

/* Montgomery product a*b*R^{-1} mod m (CIOS, 4 limbs). */
static inline void FN(mul)(FN(t) *r, const FN(t) *a, const FN(t) *b) {
    uint64_t t[6] = {0, 0, 0, 0, 0, 0};
    for (int i = 0; i < 4; i++) {
        unsigned __int128 c = 0;
        for (int j = 0; j < 4; j++) {
            c += (unsigned __int128)a->l[j] * b->l[i] + t[j];
            t[j] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[4] = (uint64_t)c;
        t[5] = (uint64_t)(c >> 64);
        uint64_t q = t[0] * F_INV;
        c = (unsigned __int128)q * FN(MOD)[0] + t[0];
        c >>= 64;
        for (int j = 1; j < 4; j++) {
            c += (unsigned __int128)q * FN(MOD)[j] + t[j];
            t[j - 1] = (uint64_t)c;
            c >>= 64;
        }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = t[5] + (uint64_t)(c >> 64);
    }
    if (t[4] || FN(geq_mod)(t)) FN(sub_mod_raw)(t);
    memcpy(r->l, t, 32);
}
static inline void FN(sqr)(FN(t) *r, const FN(t) *a) { FN(mul)(r, a, a); }

/* canonical integer -> Montgomery, and back (ark: from_bigint / into_bigint) */
static inline void FN(to_mont)(FN(t) *r, const uint64_t a[4]) {
    FN(t) x, r2;
    memcpy(x.l, a, 32);
    memcpy(r2.l, FN(R2), 32);
    FN(mul)(r, &x, &r2);
}
static inline void FN(from_mont)(uint64_t r[4], const FN(t) *a) {
    FN(t) one = {{1, 0, 0, 0}}, o;
    FN(mul)(&o, a, &one);
    memcpy(r, o.l, 32);
}

/* r = a^e, e a canonical 256-bit exponent (4 limbs LE) */
static void FN(pow)(FN(t) *r, const FN(t) *a, const uint64_t e[4]) {
    FN(t) acc; FN(set_one)(&acc);
    int started = 0;
    for (int i = 255; i >= 0; i--) {
        if (started) FN(sqr)(&acc, &acc);
        if ((e[i / 64] >> (i % 64)) & 1) { FN(mul)(&acc, &acc, a); started = 1; }
    }
    *r = acc;
}
static void FN(pow_u64)(FN(t) *r, const FN(t) *a, uint64_t e) {
    uint64_t ee[4] = {e, 0, 0, 0};
    FN(pow)(r, a, ee);
}
/* Fermat inverse; 0 -> 0 */
static void FN(inv)(FN(t) *r, const FN(t) *a) {
    uint64_t e[4];
    memcpy(e, FN(MOD), 32);
    e[0] -= 2; /* low limb ends in ...0001, no borrow */
    FN(pow)(r, a, e);
}
/* Montgomery's trick: in-place batch inversion of n elements (zeros stay zero). */
static void FN(batch_inv)(FN(t) *v, size_t n) {
    if (n == 0) return;
    FN(t) *pre = (FN(t) *)malloc(n * sizeof(FN(t)));
    FN(t) acc; FN(set_one)(&acc);
    for (size_t i = 0; i < n; i++) {
        pre[i] = acc;
        if (!FN(is_zero)(&v[i])) FN(mul)(&acc, &acc, &v[i]);
    }
    FN(t) ia; FN(inv)(&ia, &acc);
    for (size_t i = n; i-- > 0;) {
        if (FN(is_zero)(&v[i])) continue;
        FN(t) t; FN(mul)(&t, &ia, &pre[i]);
        FN(mul)(&ia, &ia, &v[i]);
        v[i] = t;
    }
    free(pre);
}

/* Square root by Tonelli-Shanks (two-adicity 32, fp.rs:21 / fq.rs:19).  Returns 1 and sets r if a is a
 * square, else 0.  Used only for point decompression of the compressed SRS fixtures
 * (utils/src/serialization.rs:65-84 -> ark compressed form). */
static int FN(sqrt)(FN(t) *r, const FN(t) *a) {
    if (FN(is_zero)(a)) { FN(set_zero)(r); return 1; }
    /* m - 1 = 2^32 * T */
    uint64_t T[4], Tm1h[4];
    /* T = (m-1) >> 32 */
    uint64_t m1[4]; memcpy(m1, FN(MOD), 32); m1[0] -= 1;
    for (int i = 0; i < 4; i++) T[i] = (m1[i] >> 32) | (i < 3 ? (m1[i + 1] << 32) : 0);
    /* (T-1)/2 */
    uint64_t t1[4]; memcpy(t1, T, 32); t1[0] -= 1;
    for (int i = 0; i < 4; i++) Tm1h[i] = (t1[i] >> 1) | (i < 3 ? (t1[i + 1] << 63) : 0);
    FN(t) w, x, b, z;
    FN(pow)(&w, a, Tm1h);           /* w = a^((T-1)/2) */
    FN(mul)(&x, a, &w);             /* x = a^((T+1)/2) */
    FN(mul)(&b, &x, &w);            /* b = a^T */
    memcpy(z.l, FN(ROOT), 32);      /* z = g^T, order 2^32 */
    int v = 32;
    FN(t) one; FN(set_one)(&one);
    while (!FN(eq)(&b, &one)) {
        int k = 0;
        FN(t) b2 = b;
        while (!FN(eq)(&b2, &one)) { FN(sqr)(&b2, &b2); k++; if (k >= v) return 0; }
        FN(t) ww = z;
        for (int j = 0; j < v - k - 1; j++) FN(sqr)(&ww, &ww);
        FN(sqr)(&z, &ww);
        FN(mul)(&b, &b, &z);
        FN(mul)(&x, &x, &ww);
        v = k;
    }
    *r = x;
    return 1;
}

/* n-th root of unity for n = 2^log_n: ROOT^(2^(32-log_n))  (ark: F::get_root_of_unity; generator 5,
 * fp.rs:10 / fq.rs:10; domain chain kimchi/src/circuits/domains.rs:40-69). */
static void FN(root_of_unity)(FN(t) *r, unsigned log_n) {
    FN(t) w; memcpy(w.l, FN(ROOT), 32);
    for (unsigned i = log_n; i < 32; i++) FN(sqr)(&w, &w);
    *r = w;
}
