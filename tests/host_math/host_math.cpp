// tests/host_math/host_math.cpp — compiles the DEVICE math headers (field.cuh, curve.cuh) for the host, where the
// PTX carry-chain primitives are emulated instruction by instruction, so the exact mad.lo.cc/madc.hi.cc sequences
// the GPU will execute can be checked against the CPU oracle on a machine without a GPU.  Test infrastructure only.
#include <cstring>
#include "../../proof_systems_b200/csrc/curve.cuh"
#include "../../proof_systems_b200/csrc/ntt_butterfly.cuh"
#include <vector>
using namespace zkb;

template <class F> static void mul_(const uint32_t* a, const uint32_t* b, uint32_t* r) {
    fe x, y; memcpy(x.v, a, 32); memcpy(y.v, b, 32);
    fe z = fe_mul<F>(x, y); memcpy(r, z.v, 32);
}
#define BIN(name, fn)                                                                                   \
    extern "C" void hm_##name(int fid, const uint32_t* a, const uint32_t* b, uint32_t* r, size_t n) {   \
        for (size_t i = 0; i < n; i++) {                                                                \
            fe x, y; memcpy(x.v, a + 8 * i, 32); memcpy(y.v, b + 8 * i, 32);                            \
            fe z = fid == 0 ? fn<FpParams>(x, y) : fn<FqParams>(x, y);                                  \
            memcpy(r + 8 * i, z.v, 32);                                                                 \
        }                                                                                               \
    }
BIN(mul, fe_mul)
BIN(add, fe_add)
BIN(sub, fe_sub)
#define UN(name, fn)                                                                       \
    extern "C" void hm_##name(int fid, const uint32_t* a, uint32_t* r, size_t n) {         \
        for (size_t i = 0; i < n; i++) {                                                   \
            fe x; memcpy(x.v, a + 8 * i, 32);                                              \
            fe z = fid == 0 ? fn<FpParams>(x) : fn<FqParams>(x);                           \
            memcpy(r + 8 * i, z.v, 32);                                                    \
        }                                                                                  \
    }
UN(neg, fe_neg)
UN(inv, fe_inv)
UN(to_mont, fe_to_mont)
UN(from_mont, fe_from_mont)

extern "C" void hm_consts(int fid, uint32_t* one, uint32_t* r2, uint32_t* root) {
    for (int i = 0; i < 8; i++) {
        one[i] = fid == 0 ? FpParams::R(i) : FqParams::R(i);
        r2[i] = fid == 0 ? FpParams::R2(i) : FqParams::R2(i);
        root[i] = fid == 0 ? FpParams::ROOT(i) : FqParams::ROOT(i);
    }
}

// curve: cid 0 = Pallas (base Fp), 1 = Vesta (base Fq).  op: 0 madd(p xyzz, q affine), 1 add(p, q xyzz), 2 dbl(p)
// points: xyzz = 32 u32, affine = 16 u32.  Output always affine (16 u32) AND xyzz (32 u32).
template <class F> static void curve_op_(int op, const uint32_t* p, const uint32_t* q, uint32_t* out_aff, uint32_t* out_xyzz) {
    xyzz_t P, R; memcpy(&P, p, sizeof P);
    if (op == 0) { affine_t Q; memcpy(&Q, q, sizeof Q); R = xyzz_madd<F>(P, Q); }
    else if (op == 1) { xyzz_t Q; memcpy(&Q, q, sizeof Q); R = xyzz_add<F>(P, Q); }
    else R = xyzz_dbl<F>(P);
    affine_t A = xyzz_to_affine<F>(R);
    memcpy(out_aff, &A, sizeof A); memcpy(out_xyzz, &R, sizeof R);
}
extern "C" void hm_curve_op(int cid, int op, const uint32_t* p, const uint32_t* q, uint32_t* out_aff, uint32_t* out_xyzz) {
    if (cid == 0) curve_op_<FpParams>(op, p, q, out_aff, out_xyzz);
    else curve_op_<FqParams>(op, p, q, out_aff, out_xyzz);
}
// sum of n affine points through repeated madd, result affine
extern "C" void hm_sum_affine(int cid, const uint32_t* pts, size_t n, uint32_t* out_aff) {
    xyzz_t acc = xyzz_identity();
    for (size_t i = 0; i < n; i++) {
        affine_t Q; memcpy(&Q, pts + 16 * i, sizeof Q);
        acc = cid == 0 ? xyzz_madd<FpParams>(acc, Q) : xyzz_madd<FqParams>(acc, Q);
    }
    affine_t A = cid == 0 ? xyzz_to_affine<FpParams>(acc) : xyzz_to_affine<FqParams>(acc);
    memcpy(out_aff, &A, sizeof A);
}


// The layer schedule of k_ntt_pass on one column of S = 2^log_s rows (in place, natural order in, bit-reversed order out):
// radix 4: an odd log_s starts with one radix-2 layer, then radix-2^2 units; radix 2: plain layers.  small = w_1024^i, 512 entries.
template <class F> static void ntt_column_(uint32_t* data, unsigned log_s, const uint32_t* small32, int radix) {
    const unsigned S = 1u << log_s;
    fe* x = reinterpret_cast<fe*>(data);
    const fe* small = reinterpret_cast<const fe*>(small32);
    int l = (int)log_s - 1;
    if (radix == 4 && (log_s & 1) && l >= 0) {
        for (unsigned j = 0; j < S / 2; j++) {
            unsigned i0, tw;
            ntt_index2(j, (unsigned)l, i0, tw);
            ntt_bfly2<F>(x[i0], x[i0 + (1u << l)], l ? &small[tw] : nullptr);
        }
        l--;
    }
    while (l >= 0) {
        if (radix == 4) {
            for (unsigned j = 0; j < S / 4; j++) {
                unsigned e0, q, tA, tB, tC;
                ntt_index4(j, (unsigned)l, e0, q, tA, tB, tC);
                const unsigned h = 1u << l;
                ntt_bfly4<F>(x[e0], x[e0 + q], x[e0 + h], x[e0 + h + q], small[tA], small[tB], l >= 2 ? &small[tC] : nullptr);
            }
            l -= 2;
        } else {
            for (unsigned j = 0; j < S / 2; j++) {
                unsigned i0, tw;
                ntt_index2(j, (unsigned)l, i0, tw);
                ntt_bfly2<F>(x[i0], x[i0 + (1u << l)], l ? &small[tw] : nullptr);
            }
            l--;
        }
    }
}
extern "C" void hm_ntt_column(int fid, uint32_t* data, unsigned log_s, const uint32_t* small, int radix) {
    if (fid == 0) ntt_column_<FpParams>(data, log_s, small, radix);
    else ntt_column_<FqParams>(data, log_s, small, radix);
}

// The decimation-in-time schedule of k_ntt_pass (csrc/ntt.cu) on one column of S = 2^log_s rows: bit-reversed load, then — for
// S >= 64 — the REGISTER stage emulated lane by lane (32 lanes of a warp as arrays, the shuffles as array reads: exactly the
// per-lane functions the kernel calls), then the shared-memory layers; natural order out.  small = w_1024^i, 512 entries.
template <class F> static void ntt_column_dit_(uint32_t* data, unsigned log_s, const uint32_t* small32) {
    const unsigned S = 1u << log_s;
    fe* x = reinterpret_cast<fe*>(data);
    const fe* small = reinterpret_cast<const fe*>(small32);
    std::vector<fe> sm(S);
    for (unsigned r = 0; r < S; r++) {
        unsigned i = 0;
        for (unsigned b = 0; b < log_s; b++) i |= ((r >> b) & 1u) << (log_s - 1 - b);
        sm[i] = x[r];
    }
    unsigned l0 = 0;
    if (log_s >= 6) {
        for (unsigned q = 0; q < (S >> 6); q++) {
            fe a[32], b[32];
            for (unsigned t = 0; t < 32; t++) { a[t] = sm[(q << 6) + t]; b[t] = sm[(q << 6) + 32 + (t ^ 31u)]; }
            for (unsigned l = 0; l < 5; l++) {
                bool a_hi[32];
                fe tx[32], sa[32], sb[32];
                for (unsigned t = 0; t < 32; t++) ntt_lane_pre<F>(t, l, a[t], b[t], small, a_hi[t], tx[t], sa[t], sb[t]);
                for (unsigned t = 0; t < 32; t++) ntt_lane_post<F>(a_hi[t], tx[t], sa[t ^ (1u << l)], sb[t ^ (1u << l)], a[t], b[t]);
            }
            fe bn[32];
            for (unsigned t = 0; t < 32; t++) bn[t] = b[t ^ 31u];
            for (unsigned t = 0; t < 32; t++) {
                ntt_lane_last<F>(t, a[t], bn[t], small);
                sm[(q << 6) + t] = a[t];
                sm[(q << 6) + 32 + t] = bn[t];
            }
        }
        l0 = 6;
    }
    for (unsigned l = l0; l < log_s; l++) {
        for (unsigned j = 0; j < S / 2; j++) {
            unsigned i0, tw;
            ntt_index2_dit(j, l, i0, tw);
            ntt_bfly2_dit<F>(sm[i0], sm[i0 + (1u << l)], l ? &small[tw] : nullptr);
        }
    }
    for (unsigned k = 0; k < S; k++) x[k] = sm[k];
}
extern "C" void hm_ntt_column_dit(int fid, uint32_t* data, unsigned log_s, const uint32_t* small) {
    if (fid == 0) ntt_column_dit_<FpParams>(data, log_s, small);
    else ntt_column_dit_<FqParams>(data, log_s, small);
}
