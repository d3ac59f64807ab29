// ntt_butterfly.cuh — the butterflies of the tile pass (ntt.cu) on VALUES, and their index maps; shared by k_ntt_pass and by
// the host-compiled test (tests/host_math/host_math.cpp), so the layer schedule is checked against the oracle without a GPU.
//
// Decimation in frequency inside a column of S = 2^log_s rows, natural order in, bit-reversed order out.  Layer l (half size
// h = 2^l) pairs rows (i, i + h) of every group of 2h rows:  (u, v) -> (u + v, (u - v) * w_{2h}^pos),  pos = i mod h, and
// w_{2h}^pos = small[pos << (9 - l)] with small[i] = w_1024^i (512 entries; no twiddle at l = 0).
// A radix-2^2 unit does layers l and l-1 (l >= 1) on the four rows e0 + {0, q, h, h + q}, q = h/2, in registers: the same 4
// multiplications and 8 additions as the four radix-2 butterflies it replaces, but one shared-memory round trip and one
// barrier instead of two.
#pragma once
#include "field.cuh"

namespace zkb {

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
