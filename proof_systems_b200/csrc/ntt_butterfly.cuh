// ntt_butterfly.cuh — the butterflies of the tile pass (ntt.cu) on VALUES, and their index maps; shared by k_ntt_pass and by
// the host-compiled test (tests/host_math/host_math.cpp), so the layer schedule is checked against the oracle without a GPU.
//
// Decimation in TIME inside a column of S = 2^log_s rows: the column is loaded in bit-reversed order and comes out in natural
// order.  Layer l (half size h = 2^l, l = 0 .. log_s-1) pairs rows (i, i + h) of every group of 2h rows:
//     (u, v) -> (u + w v, u - w v),   w = w_{2h}^pos,  pos = i mod h,  w_{2h}^pos = small[pos << (9 - l)]
// with small[i] = w_1024^i (512 entries; no multiplication at l = 0).
//
// Layers 0..5 (h <= 32) never leave a chunk of 64 consecutive rows.  A WARP runs them on one chunk in registers: lane t holds
//     a = chunk[t]               (block A, position t)
//     b = chunk[32 + (t ^ 31)]   (block B, positions COMPLEMENTED)
// so that in every layer h < 32 exactly one of the lane's two elements is the upper ("hi") element of its pair: every lane does
// one multiplication per layer, nothing idles.  The hi element is multiplied by the twiddle in place, the pair partners (lane
// t ^ h, both blocks) swap what the other needs with one shuffle per block, and each side finishes with one addition or
// subtraction.  Before the last layer (h = 32: the pair is (A[p], B[p])) block B is un-complemented with one more shuffle.
// The functions below are the per-lane halves of such a layer; ntt.cu puts warp shuffles between them, the host test arrays.
//
// (The decimation-in-frequency helpers further down are the round-1 schedule, kept for the host test's cross-check.)
#pragma once
#include "field.cuh"

namespace zkb {

// ------------------------------------------------------------------------------------------------ decimation in time
template <class F> ZK_HD void ntt_bfly2_dit(fe& u, fe& v, const fe* tw) {
    const fe t = tw ? fe_mul<F>(v, *tw) : v;
    const fe s = fe_add<F>(u, t);
    v = fe_sub<F>(u, t);
    u = s;
}
// butterfly j (of S/2) of layer l: rows i0, i0 + 2^l and the index of its twiddle in `small` (unused at l = 0)
ZK_HD void ntt_index2_dit(unsigned j, unsigned l, unsigned& i0, unsigned& tw) {
    const unsigned h = 1u << l, grp = j >> l, pos = j & (h - 1);
    i0 = (grp << (l + 1)) + pos;
    tw = pos << (9 - l);
}

// Register stage, layer l < 5 (h = 2^l < 32), lane t.  Before the exchange: which block holds the hi element, that element times
// its twiddle (tx), and what the lane hands to lane t ^ h for each block.
template <class F> ZK_HD void ntt_lane_pre(unsigned t, unsigned l, const fe& a, const fe& b, const fe* small, bool& a_hi, fe& tx, fe& send_a, fe& send_b) {
    const unsigned h = 1u << l;
    a_hi = (t & h) != 0;                                  // block A position t; block B position t ^ 31 has the opposite bit
    const unsigned p = a_hi ? t : (t ^ 31u);              // position of the lane's hi element inside its block of 32
    const fe x = a_hi ? a : b;
    tx = l ? fe_mul<F>(x, small[(p & (h - 1)) << (9 - l)]) : x;
    send_a = a_hi ? tx : a;
    send_b = a_hi ? b : tx;
}
// After the exchange (recv_* = the partner's send_*): lo side u + w v, hi side u - w v.
template <class F> ZK_HD void ntt_lane_post(bool a_hi, const fe& tx, const fe& recv_a, const fe& recv_b, fe& a, fe& b) {
    if (a_hi) {
        a = fe_sub<F>(recv_a, tx);
        b = fe_add<F>(b, recv_b);
    } else {
        a = fe_add<F>(a, recv_a);
        b = fe_sub<F>(recv_b, tx);
    }
}
// Last register layer (l = 5, h = 32) once block B is back in natural order: the pair is (a, b) of the same lane, pos = t.
template <class F> ZK_HD void ntt_lane_last(unsigned t, fe& a, fe& b, const fe* small) {
    ntt_bfly2_dit<F>(a, b, &small[t << 4]);
}

// ------------------------------------------------------------------------------------------------ decimation in frequency
// Natural order in, bit-reversed order out.  Layer l pairs rows (i, i + h):  (u, v) -> (u + v, (u - v) * w_{2h}^pos).
template <class F> ZK_HD void ntt_bfly2(fe& u, fe& v, const fe* tw) {
    const fe s = fe_add<F>(u, v);
    fe d = fe_sub<F>(u, v);
    if (tw) d = fe_mul<F>(d, *tw);
    u = s;
    v = d;
}

// wA = w_{2h}^pos, wB = w_{2h}^(pos + q), wC = w_{h}^pos (nullptr when the second layer is layer 0)
template <class F> ZK_HD void ntt_bfly4(fe& x0, fe& x1, fe& x2, fe& x3, const fe& wA, const fe& wB, const fe* wC) {
    const fe a0 = fe_add<F>(x0, x2), a1 = fe_add<F>(x1, x3);
    const fe b0 = fe_mul<F>(fe_sub<F>(x0, x2), wA), b1 = fe_mul<F>(fe_sub<F>(x1, x3), wB);
    x0 = fe_add<F>(a0, a1);
    x1 = fe_sub<F>(a0, a1);
    x2 = fe_add<F>(b0, b1);
    x3 = fe_sub<F>(b0, b1);
    if (wC) { x1 = fe_mul<F>(x1, *wC); x3 = fe_mul<F>(x3, *wC); }
}

// radix-2 butterfly j (of S/2) of layer l: rows i0, i0 + h and the index of its twiddle in `small`
ZK_HD void ntt_index2(unsigned j, unsigned l, unsigned& i0, unsigned& tw) {
    const unsigned h = 1u << l, grp = j >> l, pos = j & (h - 1);
    i0 = (grp << (l + 1)) + pos;
    tw = pos << (9 - l);
}
// radix-2^2 unit j (of S/4) of layers (l, l-1), l >= 1: first row e0, quarter q, twiddle indices (tC unused when l == 1)
ZK_HD void ntt_index4(unsigned j, unsigned l, unsigned& e0, unsigned& q, unsigned& tA, unsigned& tB, unsigned& tC) {
    q = 1u << (l - 1);
    const unsigned grp = j >> (l - 1), pos = j & (q - 1);
    e0 = (grp << (l + 1)) + pos;
    tA = pos << (9 - l);
    tB = tA + 256;
    tC = pos << (10 - l);
}

}  // namespace zkb
