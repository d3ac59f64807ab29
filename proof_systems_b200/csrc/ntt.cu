// ntt.cu — kernels and launch plan of the Pasta-field NTT (see ntt.cuh for semantics and reference call sites).
#include "ntt.cuh"
#include "ntt_butterfly.cuh"

namespace zkb {

// ---------------------------------------------------------------------------------------------- table construction
// bases[0] = w      (n-th root of unity, or its inverse)        bases[1] = w^1024
// bases[2] = g      (coset generator 5, or its inverse)         bases[3] = g^1024
// bases[4] = scale  (1 forward, n^-1 inverse)                   bases[5] = 1
template <class F> __global__ void k_ntt_setup(fe* bases, unsigned log_n, int inverse) {
    if (threadIdx.x != 0 || blockIdx.x != 0) return;
    fe w;
#pragma unroll
    for (int i = 0; i < 8; i++) w.v[i] = F::ROOT(i);
    for (unsigned i = log_n; i < 32; i++) w = fe_sqr<F>(w);  // (5^T)^(2^(32-log_n)); kimchi/src/circuits/domains.rs:40-69
    fe five = fe_zero();
    five.v[0] = 5;
    fe g = fe_to_mont<F>(five);  // multiplicative generator (fp.rs:10 / fq.rs:10) = ark's default coset offset
    fe scale = fe_one<F>();
    if (inverse) {
        w = fe_inv<F>(w);
        g = fe_inv<F>(g);
        fe n = fe_zero();
        n.v[0] = 1u << log_n;  // log_n <= 30
        scale = fe_inv<F>(fe_to_mont<F>(n));
    }
    bases[0] = w;
    bases[1] = fe_pow_u64<F>(w, 1024);
    bases[2] = g;
    bases[3] = fe_pow_u64<F>(g, 1024);
    bases[4] = scale;
    bases[5] = fe_one<F>();
}

// out[i] = base^i * scale, i < count
template <class F> __global__ void k_pow_table(fe* out, const fe* base, const fe* scale, unsigned count) {
    unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    fe b = *base;
    fe r = fe_mul<F>(fe_pow_u64<F>(b, i), *scale);
    store_fe(out + i, r);
}

template <class F> int ntt_build_small_table(fe* d_small, bool inverse, cudaStream_t st) {
    fe* bases;
    ZK_CUDA(cudaMalloc(&bases, 6 * sizeof(fe)));
    k_ntt_setup<F><<<1, 32, 0, st>>>(bases, 10, inverse ? 1 : 0);
    k_pow_table<F><<<2, 256, 0, st>>>(d_small, bases + 0, bases + 5, 512);
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaStreamSynchronize(st));
    ZK_CUDA(cudaFree(bases));
    return ZK_OK;
}

template <class F> int ntt_build_tables(NttTables& t, unsigned log_n, bool inverse, cudaStream_t st) {
    fe* bases;
    ZK_CUDA(cudaMalloc(&bases, 6 * sizeof(fe)));
    fe* blk;
    ZK_CUDA(cudaMalloc(&blk, 4 * 1024 * sizeof(fe)));
    t.lo = blk; t.hi = blk + 1024; t.clo = blk + 2048; t.chi = blk + 3072;
    k_ntt_setup<F><<<1, 32, 0, st>>>(bases, log_n, inverse ? 1 : 0);
    k_pow_table<F><<<4, 256, 0, st>>>(t.lo, bases + 0, bases + 4, 1024);   // w^i * (n^-1 if inverse)
    k_pow_table<F><<<4, 256, 0, st>>>(t.hi, bases + 1, bases + 5, 1024);   // w^(1024 i)
    k_pow_table<F><<<4, 256, 0, st>>>(t.clo, bases + 2, bases + 5, 1024);  // g^i
    k_pow_table<F><<<4, 256, 0, st>>>(t.chi, bases + 3, bases + 5, 1024);  // g^(1024 i)
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaStreamSynchronize(st));
    ZK_CUDA(cudaFree(bases));
    return ZK_OK;
}

void ntt_free_tables(NttTables& t) {
    if (t.lo) cudaFree(t.lo);
    t = NttTables();
}

// ---------------------------------------------------------------------------------------------- the tile pass
// Shared-memory tile, limb-major: word (limb l, column c, row r) at sm[(l*T + c)*(S+1) + r].  Adjacent threads work on
// adjacent rows of one column, so every layer is bank-conflict free; the +1 pitch keeps the column-fastest load/store
// conflict free as well.
template <class F> __global__ void __launch_bounds__(NTT_THREADS, 3) k_ntt_pass(NttPassParams p) {
    extern __shared__ uint32_t sm[];
    const unsigned S = 1u << p.log_s, T = 1u << p.log_t, PITCH = S + 1;
    const unsigned tid = threadIdx.x;
    const size_t col0 = (size_t)blockIdx.x * T;
    const fe* in = p.in + (size_t)blockIdx.y * p.batch_stride;
    fe* out = p.out + (size_t)blockIdx.y * p.batch_stride;

    for (unsigned idx = tid; idx < S * T; idx += NTT_THREADS) {
        unsigned r, c;
        if (p.col_fastest_in) { c = idx & (T - 1); r = idx >> p.log_t; }
        else { r = idx & (S - 1); c = idx >> p.log_s; }
        size_t col = col0 + c;
        fe v = fe_zero();
        if (col < p.ncols) {
            // zero padding is decided by the position inside the polynomial: the row index when every column is a
            // polynomial of its own, the row-major matrix index otherwise
            size_t off = r * p.in_row_stride + col * p.in_col_stride;
            size_t pos = p.col_is_poly ? (size_t)r : off;
            if (pos < p.in_len) v = load_fe(in + off);
        }
        uint32_t* dst = sm + c * PITCH + r;
#pragma unroll
        for (int l = 0; l < 8; l++) dst[l * T * PITCH] = v.v[l];
    }
    __syncthreads();

    // decimation in frequency: natural order in, bit-reversed order out (undone by the store's index map).  The butterfly and
    // its index map live in ntt_butterfly.cuh (also compiled for the host and checked against the oracle there).  A radix-2^2
    // schedule (two layers per shared-memory round trip, same header) was measured: identical times — the pass is bound by
    // the integer pipes, not by shared memory or barriers — so the simpler radix-2 loop stays.
    for (int l = (int)p.log_s - 1; l >= 0; l--) {
        for (unsigned bidx = tid; bidx < (S * T) / 2; bidx += NTT_THREADS) {
            const unsigned c = bidx >> (p.log_s - 1), j = bidx & (S / 2 - 1);
            unsigned i0, tw;
            ntt_index2(j, (unsigned)l, i0, tw);
            uint32_t* a0 = sm + c * PITCH + i0;
            uint32_t* a1 = a0 + (1u << l);
            fe u, v;
#pragma unroll
            for (int k = 0; k < 8; k++) { u.v[k] = a0[k * T * PITCH]; v.v[k] = a1[k * T * PITCH]; }
            if (l != 0) { const fe w = load_fe_nc(p.small + tw); ntt_bfly2<F>(u, v, &w); }   // w_S^(pos*S/2h) = w_1024^(pos*512/h)
            else ntt_bfly2<F>(u, v, nullptr);
#pragma unroll
            for (int k = 0; k < 8; k++) { a0[k * T * PITCH] = u.v[k]; a1[k * T * PITCH] = v.v[k]; }
        }
        __syncthreads();
    }

    for (unsigned idx = tid; idx < S * T; idx += NTT_THREADS) {
        unsigned k, c;
        if (p.col_fastest_out) { c = idx & (T - 1); k = idx >> p.log_t; }
        else { k = idx & (S - 1); c = idx >> p.log_s; }
        size_t col = col0 + c;
        if (col >= p.ncols) continue;
        unsigned r = p.log_s ? (__brev(k) >> (32 - p.log_s)) : 0;
        const uint32_t* src = sm + c * PITCH + r;
        fe v;
#pragma unroll
        for (int l = 0; l < 8; l++) v.v[l] = src[l * T * PITCH];
        if (p.lo) {
            // inter-pass twiddle w_n^(col*k) (times n^-1 on the inverse transform), built from the two small tables
            unsigned e = (unsigned)col * k;
            fe tw = load_fe_nc(p.lo + (e & 1023));
            if (e >> 10) tw = fe_mul<F>(tw, load_fe_nc(p.hi + (e >> 10)));
            v = fe_mul<F>(v, tw);
        }
        if (p.scale) v = fe_mul<F>(v, load_fe_nc(p.scale));
        store_fe(out + k * p.out_row_stride + col * p.out_col_stride, v);
    }
}

// x[j] *= g^(+-j) for j < len (forward coset: before the transform; inverse coset: after it)
template <class F> __global__ void k_coset_scale(fe* data, size_t n, size_t len, const fe* clo, const fe* chi) {
    size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= len) return;
    fe* p = data + (size_t)blockIdx.y * n + j;
    fe tw = load_fe_nc(clo + (j & 1023));
    if (j >> 10) tw = fe_mul<F>(tw, load_fe_nc(chi + (j >> 10)));
    store_fe(p, fe_mul<F>(load_fe(p), tw));
}

template <class F> static int launch_pass(const NttPassParams& p, size_t batch_y, cudaStream_t st) {
    const unsigned S = 1u << p.log_s, T = 1u << p.log_t;
    size_t smem = (size_t)8 * T * (S + 1) * sizeof(uint32_t);
    // the attribute is per (function, device): set it on every launch (a host-side table write) instead of caching a
    // per-process flag that a second context on another device would never see
    ZK_CUDA(cudaFuncSetAttribute(k_ntt_pass<F>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    dim3 grid((unsigned)((p.ncols + T - 1) / T), (unsigned)batch_y);
    k_ntt_pass<F><<<grid, NTT_THREADS, smem, st>>>(p);
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}

// Columns per CTA: as many as the 128 KiB tile allows, but never so many that the grid cannot fill the 148 SMs twice
// (a 2^16 transform is only 256 columns; 16-column tiles would leave 132 SMs idle).  Keeps >= 4 columns (128 B runs).
static unsigned pick_log_t(unsigned log_s, size_t ncols, size_t batch) {
    size_t t = NTT_TILE_ELEMS >> log_s;
    if (t > ncols) t = ncols;
    if (t < 1) t = 1;
    while (t > 2 && ((ncols + t - 1) / t) * batch < 3 * 148) t /= 2;
    unsigned l = 0;
    while (((size_t)1 << (l + 1)) <= t) l++;
    return l;
}

static unsigned floor_log2(size_t x) {
    unsigned l = 0;
    while ((x >> (l + 1)) != 0) l++;
    return l;
}

template <class F>
int ntt_run(fe* d_data, fe* d_tmp, const fe* d_small, const NttTables& tabs, unsigned log_n, size_t batch, size_t in_len,
            bool inverse, bool coset, cudaStream_t st, unsigned* launches) {
    if (log_n > NTT_MAX_LOG_N) { zk_set_error("ntt: log_n %u > %u not supported", log_n, NTT_MAX_LOG_N); return ZK_ERR_INVALID; }
    const size_t n = (size_t)1 << log_n;
    if (in_len == 0 || in_len > n) in_len = n;
    if (batch == 0) return ZK_OK;
    unsigned nl = 0;

    if (coset && !inverse) {
        dim3 grid((unsigned)((in_len + 255) / 256), (unsigned)batch);
        k_coset_scale<F><<<grid, 256, 0, st>>>(d_data, n, in_len, tabs.clo, tabs.chi);
        nl++;
    }
    if (log_n == 0) {
        // n = 1: both transforms are the identity (n^-1 = 1)
    } else if (log_n <= NTT_MAX_LOG_SUB) {
        // one pass, the polynomials of the batch are the tile columns
        NttPassParams p{};
        p.in = d_data; p.out = d_data; p.small = d_small; p.lo = nullptr; p.hi = nullptr;
        p.log_s = log_n;
        unsigned t = NTT_TILE_ELEMS >> log_n;
        if (t > 64) t = 64;  // pitch overhead of very short columns
        while (t > 1 && t / 2 >= batch) t /= 2;  // do not stage more columns than there are polynomials
        p.log_t = floor_log2(t);
        p.ncols = (unsigned)batch;
        p.in_row_stride = 1; p.in_col_stride = n; p.out_row_stride = 1; p.out_col_stride = n;
        p.batch_stride = 0; p.in_len = in_len;
        p.col_is_poly = 1;
        p.col_fastest_in = 0; p.col_fastest_out = 0;
        p.scale = inverse ? tabs.lo : nullptr;  // entry 0 of the inverse lo table is n^-1
        int rc = launch_pass<F>(p, 1, st);
        if (rc) return rc;
        nl++;
    } else {
        const unsigned log_n1 = (log_n + 1) / 2, log_n2 = log_n - log_n1;
        const size_t n1 = (size_t)1 << log_n1, n2 = (size_t)1 << log_n2;
        // pass 1: n2 column transforms of size n1 (stride n2), times w_n^(j2 k1); data -> tmp, same layout
        NttPassParams p{};
        p.in = d_data; p.out = d_tmp; p.small = d_small; p.lo = tabs.lo; p.hi = tabs.hi;
        p.log_s = log_n1;
        p.log_t = pick_log_t(log_n1, n2, batch);
        p.ncols = (unsigned)n2;
        p.in_row_stride = n2; p.in_col_stride = 1; p.out_row_stride = n2; p.out_col_stride = 1;
        p.batch_stride = n; p.in_len = in_len; p.scale = nullptr; p.col_is_poly = 0;
        p.col_fastest_in = 1; p.col_fastest_out = 1;
        int rc = launch_pass<F>(p, batch, st);
        if (rc) return rc;
        // pass 2: n1 row transforms of size n2 (contiguous), output transposed to natural order; tmp -> data
        NttPassParams q{};
        q.in = d_tmp; q.out = d_data; q.small = d_small; q.lo = nullptr; q.hi = nullptr;
        q.log_s = log_n2;
        q.log_t = pick_log_t(log_n2, n1, batch);
        q.ncols = (unsigned)n1;
        q.in_row_stride = 1; q.in_col_stride = n2; q.out_row_stride = n1; q.out_col_stride = 1;
        q.batch_stride = n; q.in_len = n; q.scale = nullptr; q.col_is_poly = 0;
        q.col_fastest_in = 0; q.col_fastest_out = 1;
        rc = launch_pass<F>(q, batch, st);
        if (rc) return rc;
        nl += 2;
    }
    if (coset && inverse) {
        dim3 grid((unsigned)((n + 255) / 256), (unsigned)batch);
        k_coset_scale<F><<<grid, 256, 0, st>>>(d_data, n, n, tabs.clo, tabs.chi);
        nl++;
    }
    ZK_CUDA(cudaGetLastError());
    if (launches) *launches += nl;
    return ZK_OK;
}

template int ntt_build_small_table<FpParams>(fe*, bool, cudaStream_t);
template int ntt_build_small_table<FqParams>(fe*, bool, cudaStream_t);
template int ntt_build_tables<FpParams>(NttTables&, unsigned, bool, cudaStream_t);
template int ntt_build_tables<FqParams>(NttTables&, unsigned, bool, cudaStream_t);
template int ntt_run<FpParams>(fe*, fe*, const fe*, const NttTables&, unsigned, size_t, size_t, bool, bool, cudaStream_t, unsigned*);
template int ntt_run<FqParams>(fe*, fe*, const fe*, const NttTables&, unsigned, size_t, size_t, bool, bool, cudaStream_t, unsigned*);

}  // namespace zkb
