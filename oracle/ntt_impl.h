/*
 * oracle/ntt_impl.h — TEST INFRASTRUCTURE (CPU oracle), not product code.
 *
 * Radix-2 NTT over a Pasta field, restating ark-poly 0.5.0 Radix2EvaluationDomain::{fft,ifft}_in_place
 * (crate pinned in Cargo.lock:171-279, source not on disk) as reached from the reference at
 * kimchi/src/prover.rs:289,377,907,1163, kimchi/src/circuits/constraints.rs:494,
 * poly-commitment/src/utils.rs:195-198 (list: SURVEY.md §8 row a8):
 *   forward  out[i] = sum_j a[j] * (g*w^i)^j      (g = 1 for a plain domain, g = 5 for the default coset)
 *   inverse  a[j]   = g^{-j} * n^{-1} * sum_i out[i] * w^{-ij}
 * natural order in and out, w = ROOT^(2^(32-log_n)), elements in Montgomery form.
 * The w / index order / 1/n conventions are pinned by the Lagrange-basis golden vectors of
 * srs/test_*.srs (produced by the same generic routine, poly-commitment/src/ipa.rs:1161).
 *
 * Shape follows the CPU cost model of the reference: per-call root table of n/2
 * entries, log n butterfly layers parallelised per layer, one bit-reversal permutation.
 */

static void FN(bitrev_permute)(FN(t) *a, unsigned log_n, int threads) {
    size_t n = (size_t)1 << log_n;
#pragma omp parallel for num_threads(threads) schedule(static) if (n >= 65536 && threads > 1)
    for (size_t i = 0; i < n; i++) {
        size_t r = 0;
        for (unsigned b = 0; b < log_n; b++) r |= ((i >> b) & 1) << (log_n - 1 - b);
        if (i < r) { FN(t) t = a[i]; a[i] = a[r]; a[r] = t; }
    }
}

/* in-place, natural in -> natural out, using root w (order n) */
static void FN(ntt_core)(FN(t) *a, unsigned log_n, const FN(t) *w, int threads) {
    size_t n = (size_t)1 << log_n;
    if (n == 1) return;
    /* roots[i] = w^i, i < n/2 */
    FN(t) *roots = (FN(t) *)malloc((n / 2) * sizeof(FN(t)));
    {
        /* blocked so it can be parallelised like ark's roots_of_unity */
        size_t half = n / 2;
        size_t blk = 1024;
        size_t nblk = (half + blk - 1) / blk;
        FN(t) wblk; FN(pow_u64)(&wblk, w, blk);
        FN(t) *starts = (FN(t) *)malloc(nblk * sizeof(FN(t)));
        FN(set_one)(&starts[0]);
        for (size_t b = 1; b < nblk; b++) FN(mul)(&starts[b], &starts[b - 1], &wblk);
#pragma omp parallel for num_threads(threads > 32 ? 32 : threads) schedule(static)
        for (size_t b = 0; b < nblk; b++) {
            FN(t) cur = starts[b];
            size_t end = (b + 1) * blk < half ? (b + 1) * blk : half;
            for (size_t i = b * blk; i < end; i++) { roots[i] = cur; FN(mul)(&cur, &cur, w); }
        }
        free(starts);
    }
    /* decimation in frequency: natural in, bit-reversed out.  One parallel region per layer: the n/2 butterflies of a
     * layer, indexed k = grp * half + j, are cut into contiguous chunks (what ark-poly's per-layer chunking amounts to).
     * Threads are capped: at least 4096 butterflies per thread and layer, at most 32 threads (the transform is memory bound
     * beyond that; oversubscribing a 128-thread host only adds barrier cost). */
    int thr = threads;
    if (thr > 32) thr = 32;
    if ((size_t)thr > n / 8192 + 1) thr = (int)(n / 8192 + 1);
    const size_t total = n / 2;
    const size_t nchunks = (size_t)thr * 4 < total ? (size_t)thr * 4 : 1;
    for (unsigned s = 0; s < log_n; s++) {
        const size_t half = n >> (s + 1);        /* butterfly span */
        const size_t step = (size_t)1 << s;      /* root stride; also the number of groups */
#pragma omp parallel for num_threads(thr) schedule(static) if (thr > 1)
        for (size_t ch = 0; ch < nchunks; ch++) {
            size_t k0 = total * ch / nchunks, k1 = total * (ch + 1) / nchunks;
            size_t grp = k0 / half, j = k0 % half;
            for (size_t k = k0; k < k1; k++) {
                FN(t) *lo = a + grp * 2 * half + j, *hi = lo + half;
                FN(t) u = *lo, v = *hi, d;
                FN(add)(lo, &u, &v);
                FN(sub)(&d, &u, &v);
                FN(mul)(hi, &d, &roots[j * step]);
                if (++j == half) { j = 0; grp++; }
            }
        }
    }
    free(roots);
    FN(bitrev_permute)(a, log_n, thr);
}

static void FN(ntt)(FN(t) *a, unsigned log_n, int inverse, int coset, int threads) {
    size_t n = (size_t)1 << log_n;
    FN(t) w; FN(root_of_unity)(&w, log_n);
    FN(t) g;
    { uint64_t c5[4] = {5, 0, 0, 0}; FN(to_mont)(&g, c5); }
    if (!inverse) {
        if (coset) {
            FN(t) p; FN(set_one)(&p);
            for (size_t i = 0; i < n; i++) { FN(mul)(&a[i], &a[i], &p); FN(mul)(&p, &p, &g); }
        }
        FN(ntt_core)(a, log_n, &w, threads);
    } else {
        FN(t) wi; FN(inv)(&wi, &w);
        FN(ntt_core)(a, log_n, &wi, threads);
        FN(t) ninv, nn;
        uint64_t nc[4] = {(uint64_t)n, 0, 0, 0};
        FN(to_mont)(&nn, nc);
        FN(inv)(&ninv, &nn);
        FN(t) gi; FN(inv)(&gi, &g);
        FN(t) p = ninv;
        for (size_t i = 0; i < n; i++) {
            FN(mul)(&a[i], &a[i], &p);
            if (coset) FN(mul)(&p, &p, &gi);
        }
    }
}

/* definition-level O(n^2) transform for tiny cross-checks */
static void FN(dft_naive)(FN(t) *out, const FN(t) *in, unsigned log_n, int inverse) {
    size_t n = (size_t)1 << log_n;
    FN(t) w; FN(root_of_unity)(&w, log_n);
    if (inverse) FN(inv)(&w, &w);
    for (size_t i = 0; i < n; i++) {
        FN(t) wi; FN(pow_u64)(&wi, &w, i);
        FN(t) acc; FN(set_zero)(&acc);
        FN(t) p; FN(set_one)(&p);
        for (size_t j = 0; j < n; j++) {
            FN(t) t; FN(mul)(&t, &in[j], &p);
            FN(add)(&acc, &acc, &t);
            FN(mul)(&p, &p, &wi);
        }
        out[i] = acc;
    }
    if (inverse) {
        FN(t) ninv, nn;
        uint64_t nc[4] = {(uint64_t)n, 0, 0, 0};
        FN(to_mont)(&nn, nc); FN(inv)(&ninv, &nn);
        for (size_t i = 0; i < n; i++) FN(mul)(&out[i], &out[i], &ninv);
    }
}

/*
 * Permutation part of kimchi's quotient polynomial, evaluation form over d8 — restates
 * kimchi/src/circuits/polynomials/permutation.rs:223-357 (`perm`) point by point:
 *   shifts[i] = z[i]              * prod_k (w_k[i] + gamma + x_i * beta * shift_k)          (:283-302)
 *   sigmas[i] = z[(i + 8) mod m]  * prod_k (w_k[i] + gamma + sigma_k[i] * beta)             (:304-328; z_next = lagrange.d8.next.z)
 *   perm[i]   = (shifts[i] - sigmas[i]) * alpha0 * zkpm[i]                                  (:330-331)
 * with x_i = omega_m^i the points of d8 (precomputations().poly_x_d1 evaluated over d8), zkpm =
 * permutation_vanishing_polynomial_l, sigma_k = permutation_coefficients8[k].  Checker of zk_perm_quotient_dev.
 */
static void FN(perm_quot)(const FN(t) *w, size_t w_stride, const FN(t) *z, const FN(t) *sigma, size_t sigma_stride, const FN(t) *zkpm,
                          const FN(t) *beta, const FN(t) *gamma, const FN(t) *alpha0, const FN(t) *shift, unsigned next_shift,
                          unsigned log_m, FN(t) *out, int threads) {
    const size_t m = (size_t)1 << log_m;
    FN(t) omega;
    FN(root_of_unity)(&omega, log_m);
#pragma omp parallel for num_threads(threads) schedule(static)
    for (size_t blk = 0; blk < (m + 1023) / 1024; blk++) {
        FN(t) x;
        FN(pow_u64)(&x, &omega, blk * 1024);
        const size_t end = (blk + 1) * 1024 < m ? (blk + 1) * 1024 : m;
        for (size_t i = blk * 1024; i < end; i++) {
            FN(t) bx, shifts = z[i], sigmas = z[(i + next_shift) % m];
            FN(mul)(&bx, beta, &x);
            for (int k = 0; k < 7; k++) {
                FN(t) wg, t1, t2;
                FN(add)(&wg, &w[k * w_stride + i], gamma);
                FN(mul)(&t1, &bx, &shift[k]);
                FN(add)(&t1, &t1, &wg);
                FN(mul)(&shifts, &shifts, &t1);
                FN(mul)(&t2, &sigma[k * sigma_stride + i], beta);
                FN(add)(&t2, &t2, &wg);
                FN(mul)(&sigmas, &sigmas, &t2);
            }
            FN(t) d;
            FN(sub)(&d, &shifts, &sigmas);
            FN(mul)(&d, &d, alpha0);
            FN(mul)(&out[i], &d, &zkpm[i]);
            FN(mul)(&x, &x, &omega);
        }
    }
}

/* ---------------------------------------------------------------------------------------------------------------
 * RPN constraint evaluation: PolishToken::evaluate (kimchi/src/circuits/expr.rs:856-940) run at every index i of an evaluation
 * domain of out_len points, cells read with the SubEvals indexing of Expr::evaluations (expr.rs:1976-1982):
 *     cell(col, row) at i = col.evals[(len / out_len * i + domain_mult * shift) % len],  shift = 0 (Curr) | 1 (Next)
 * Opcodes (the oracle's own numbering; Challenge / Constant terms arrive as literals, feature flags resolved by the caller):
 *     0 literal k | 1 cell (col | next << 31) | 2 dup | 3 pow n | 4 add | 5 mul | 6 sub | 7 store | 8 load k
 * Returns 0, or -1 on the reference's failure modes (EmptyStack, final stack length != 1, index out of range).
 * Checker of zk_expr_eval_dev.
 */
static int FN(expr_eval)(const uint32_t *ops, const uint32_t *args, size_t n_tok, const FN(t) *literals, size_t n_lit,
                         const FN(t) *const *col_evals, const uint64_t *col_len, const uint32_t *col_mult, size_t n_cols,
                         uint64_t out_len, int accumulate, FN(t) *out, int threads) {
    int bad = 0;
#pragma omp parallel for num_threads(threads) schedule(static)
    for (uint64_t i = 0; i < out_len; i++) {
        FN(t) stack[64], cache[256];
        size_t sp = 0, nc = 0;
        int fail = 0;
        for (size_t t = 0; t < n_tok && !fail; t++) {
            const uint32_t a = args[t];
            switch (ops[t]) {
            case 0: if (a >= n_lit || sp >= 64) { fail = 1; break; } stack[sp++] = literals[a]; break;
            case 1: {
                const uint32_t c = a & 0x7fffffffu;
                if (c >= n_cols || sp >= 64) { fail = 1; break; }
                const uint64_t len = col_len[c], scale = len / out_len;
                stack[sp++] = col_evals[c][(scale * i + (uint64_t)col_mult[c] * (a >> 31)) % len];
                break;
            }
            case 2: if (sp < 1 || sp >= 64) { fail = 1; break; } stack[sp] = stack[sp - 1]; sp++; break;
            case 3: if (sp < 1) { fail = 1; break; } { FN(t) r; FN(pow_u64)(&r, &stack[sp - 1], a); stack[sp - 1] = r; } break;
            case 4: if (sp < 2) { fail = 1; break; } FN(add)(&stack[sp - 2], &stack[sp - 2], &stack[sp - 1]); sp--; break;
            case 5: if (sp < 2) { fail = 1; break; } FN(mul)(&stack[sp - 2], &stack[sp - 2], &stack[sp - 1]); sp--; break;
            case 6: if (sp < 2) { fail = 1; break; } FN(sub)(&stack[sp - 2], &stack[sp - 2], &stack[sp - 1]); sp--; break;
            case 7: if (sp < 1 || nc >= 256) { fail = 1; break; } cache[nc++] = stack[sp - 1]; break;
            case 8: if (a >= nc || sp >= 64) { fail = 1; break; } stack[sp++] = cache[a]; break;
            default: fail = 1;
            }
        }
        if (fail || sp != 1) {
#pragma omp atomic write
            bad = 1;
            continue;
        }
        if (accumulate) FN(add)(&out[i], &out[i], &stack[0]);
        else out[i] = stack[0];
    }
    return bad ? -1 : 0;
}

/* ---------------------------------------------------------------------------------------------------------------
 * DensePolynomial::divide_by_vanishing_poly(domain) for a plain (non-coset) radix-2 domain of n points, as ark-poly 0.5.0 computes it
 * (crate not vendored, Cargo.lock:171-279; call site kimchi/src/prover.rs:909): with fewer than n coefficients the quotient is zero
 * and the polynomial is the remainder; otherwise
 *     quotient  = coeffs[n..]  and, for every further chunk i = 1 .. len / n - 1, quotient[j] += coeffs[n (i + 1) + j]
 *     remainder = coeffs[..n]  and remainder[j] += quotient[j]
 * quot receives len - n coefficients, rem receives n (zero-padded).  Checker of zk_poly_divide_by_vanishing_dev.
 */
static void FN(divide_by_vanishing)(const FN(t) *f, size_t len, size_t n, FN(t) *quot, FN(t) *rem) {
    FN(t) zero;
    memset(&zero, 0, sizeof zero);
    if (len < n) {
        for (size_t j = 0; j < n; j++) rem[j] = j < len ? f[j] : zero;
        return;
    }
    const size_t qlen = len - n;
    for (size_t j = 0; j < qlen; j++) quot[j] = f[n + j];
    for (size_t i = 1; i < len / n + (len % n ? 1 : 0); i++)
        for (size_t j = 0; n * (i + 1) + j < len; j++) FN(add)(&quot[j], &quot[j], &f[n * (i + 1) + j]);
    for (size_t j = 0; j < n; j++) {
        rem[j] = f[j];
        if (j < qlen) FN(add)(&rem[j], &rem[j], &quot[j]);
    }
}

