// ntt.cuh — radix-2 number-theoretic transform over the Pasta fields on sm_100a.
//
// Drop-in semantics of ark_poly::Radix2EvaluationDomain::<F>::{fft_in_place, ifft_in_place} as the reference calls
// them (kimchi/src/prover.rs:289,377,907,1163; kimchi/src/circuits/constraints.rs:494;
// poly-commitment/src/utils.rs:195-198 — SURVEY.md §8 row a8):
//     forward   out[i] = sum_j a[j] (g w^i)^j          inverse   a[j] = g^-j n^-1 sum_i out[i] w^-ij
// natural order in and out, w = (5^T)^(2^(32-log n)) (fp.rs:10,21-27), g = 1 (plain domain) or 5 (coset),
// elements in Montgomery form, inputs shorter than the domain zero-padded.
//
// B200 shape: a four-step decomposition n = n1 * n2 (n1, n2 <= 2^10) so that a 2^20 transform is exactly two
// kernel passes; each CTA stages an S x T tile (S-point sub-transforms for T adjacent columns, 128 KiB) in shared
// memory, runs the log2(S) butterfly layers on it, and fuses the inter-pass twiddle (pass 1), the 1/n scaling
// (folded into the twiddle table) and the transposition back to natural order (pass 2) into its store.  The
// inter-pass twiddles w_n^(j2*k1) are produced on the fly from two 1024-entry tables (w_n^lo * w_n^(1024*hi)), never
// from an n-entry table.  Work is 255-bit modular integer arithmetic: no tensor cores.
#pragma once
#include "common.cuh"

namespace zkb {

constexpr unsigned NTT_MAX_LOG_SUB = 10;   // sub-transform size limit (shared-memory tile)
constexpr unsigned NTT_MAX_LOG_N = 20;     // two passes of <= 2^10
constexpr unsigned NTT_TILE_ELEMS = 2048;  // S*T elements per CTA tile (64 KiB of fe): three CTAs per SM overlap each other's barriers
constexpr unsigned NTT_THREADS = 256;

// device tables of one (field, log_n, direction)
struct NttTables {
    fe* lo = nullptr;   // [1024] w_n^(+-i)            (inverse: times n^-1)
    fe* hi = nullptr;   // [1024] w_n^(+-1024 i)
    fe* clo = nullptr;  // [1024] g^(+-i)              coset powers
    fe* chi = nullptr;  // [1024] g^(+-1024 i)
};

struct NttPassParams {
    const fe* in;
    fe* out;
    const fe* small;       // [512] w_1024^(+-i)
    const fe* lo;          // inter-pass twiddle tables (pass 1 only, else null)
    const fe* hi;
    unsigned log_s;        // sub-transform size S = 2^log_s
    unsigned log_t;        // columns per CTA T = 2^log_t
    unsigned ncols;        // number of columns in one matrix
    size_t in_row_stride, in_col_stride;    // element (r, c) of the input lives at in[b*batch_stride + r*rs + c*cs]
    size_t out_row_stride, out_col_stride;  // output (k, c)
    size_t batch_stride;
    size_t in_len;         // elements with linear input index >= in_len read as zero
    const fe* scale;       // optional factor applied at the store (device pointer, e.g. n^-1), or null
    int col_is_poly;       // 1: every column is an independent polynomial (single-pass plan)
    int col_fastest_in;    // 1: adjacent threads walk columns on load (column stride 1), 0: rows
    int col_fastest_out;
};

template <class F> int ntt_build_small_table(fe* d_small, bool inverse, cudaStream_t st);
template <class F> int ntt_build_tables(NttTables& t, unsigned log_n, bool inverse, cudaStream_t st);
void ntt_free_tables(NttTables& t);

// In-place (on d_data) transform of `batch` polynomials of 2^log_n elements each, stored back to back.
// d_tmp: scratch of the same size (used when log_n > NTT_MAX_LOG_SUB).
template <class F>
int ntt_run(fe* d_data, fe* d_tmp, const fe* d_small, const NttTables& tabs, unsigned log_n, size_t batch, size_t in_len,
            bool inverse, bool coset, cudaStream_t st, unsigned* launches);

}  // namespace zkb
