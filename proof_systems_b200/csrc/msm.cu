// msm.cu — kernels and pipeline of the Pallas/Vesta MSM (see msm.cuh for semantics and reference call sites).
#include "msm.cuh"

#include <algorithm>

namespace zkb {

unsigned msm_num_windows(unsigned c) { return (256 + c - 1) / c; }  // top signed digit cannot carry out (scalars < 2^254 + 2^127)

int msm_default_window(size_t n, bool precomputed) {
    // accumulation costs nwin(c) * n mixed additions, the bit-sliced bucket reduction ~ c * 2^(c-2) full additions
    unsigned l = 0;
    while (((size_t)1 << (l + 1)) <= n) l++;
    int c = (int)l - (precomputed ? 3 : 4);
    if (c < 4) c = 4;
    if (c > (int)MSM_MAX_WINDOW_BITS) c = MSM_MAX_WINDOW_BITS;
    return c;
}

// ---------------------------------------------------------------------------------------------- table of 2^(cw) P
template <class F> __global__ void k_build_table(affine_t* table, size_t n, unsigned c, unsigned nwin) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    affine_t p = load_affine_nc(table + i);
    xyzz_t acc = xyzz_from_affine<F>(p);
    for (unsigned w = 1; w < nwin; w++) {
        for (unsigned k = 0; k < c; k++) acc = xyzz_dbl<F>(acc);
        store_affine(table + (size_t)w * n + i, xyzz_to_affine<F>(acc));
    }
}

template <class F> int msm_bases_create(MsmBases& b, const affine_t* pts, bool pts_on_device, size_t n, unsigned c_table, cudaStream_t st) {
    if (c_table > MSM_MAX_WINDOW_BITS) { zk_set_error("msm: table window %u > %u", c_table, MSM_MAX_WINDOW_BITS); return ZK_ERR_INVALID; }
    b.n = n;
    b.c = c_table;
    b.nwin = c_table ? msm_num_windows(c_table) : 0;
    size_t rows = c_table ? b.nwin : 1;
    ZK_CUDA(cudaMalloc(&b.d_points, std::max<size_t>(rows * n, 1) * sizeof(affine_t)));
    if (n == 0) return ZK_OK;
    ZK_CUDA(cudaMemcpyAsync(b.d_points, pts, n * sizeof(affine_t), pts_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice, st));
    if (c_table) {
        k_build_table<F><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(b.d_points, n, c_table, b.nwin);
        ZK_CUDA(cudaGetLastError());
    }
    ZK_CUDA(cudaStreamSynchronize(st));
    return ZK_OK;
}

void msm_bases_free(MsmBases& b) {
    if (b.d_points) cudaFree(b.d_points);
    b = MsmBases();
}

// ---------------------------------------------------------------------------------------------- recode + histogram
// One thread per scalar: Montgomery -> canonical if asked (VariableBaseMSM::msm == into_bigint + msm_bigint), then
// signed base-2^c digits d_w in (-2^(c-1), 2^(c-1)];  digit w of scalar i at digits[w*n + i]; per-bucket counts.
template <class FS>
__global__ void k_recode(const fe* scalars, int scalars_mont, size_t n, unsigned c, unsigned nwin, unsigned groups_per_window,
                         int32_t* digits, uint32_t* counts) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    fe s = load_fe_nc(scalars + i);
    if (scalars_mont) s = fe_from_mont<FS>(s);
    const uint32_t half = 1u << (c - 1), mask = (1u << c) - 1;
    const uint32_t B = half;
    uint32_t carry = 0;
    for (unsigned w = 0; w < nwin; w++) {
        unsigned off = w * c, limb = off >> 5, sh = off & 31;
        uint32_t v = 0;
        if (limb < 8) {
            v = s.v[limb] >> sh;
            if (sh + c > 32 && limb + 1 < 8) v |= s.v[limb + 1] << (32 - sh);
        }
        uint32_t d = (v & mask) + carry;
        int32_t sd;
        if (d > half) { sd = (int32_t)d - (int32_t)(1u << c); carry = 1; }
        else { sd = (int32_t)d; carry = 0; }
        digits[(size_t)w * n + i] = sd;
        if (sd != 0) {
            uint32_t mag = (uint32_t)(sd < 0 ? -sd : sd);
            atomicAdd(&counts[(groups_per_window ? w : 0) * B + (mag - 1)], 1u);
        }
    }
}

// Single-CTA exclusive scan of counts[0..nb) -> offsets[0..nb]; clears counts (re-used as scatter cursors).
__global__ void k_scan(uint32_t* counts, uint32_t* offsets, uint32_t nb, uint32_t* total) {
    __shared__ uint32_t warp_sums[32];
    __shared__ uint32_t carry_s;
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    if (tid == 0) carry_s = 0;
    __syncthreads();
    for (uint32_t base = 0; base < nb; base += 1024) {
        uint32_t idx = base + tid;
        uint32_t v = idx < nb ? counts[idx] : 0;
        if (idx < nb) counts[idx] = 0;
        uint32_t x = v;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (unsigned)d) x += y;
        }
        if (lane == 31) warp_sums[wid] = x;
        __syncthreads();
        if (wid == 0) {
            uint32_t ws = warp_sums[lane];
            uint32_t z = ws;
#pragma unroll
            for (int d = 1; d < 32; d <<= 1) {
                uint32_t y = __shfl_up_sync(0xffffffffu, z, d);
                if (lane >= (unsigned)d) z += y;
            }
            warp_sums[lane] = z - ws;  // exclusive
        }
        __syncthreads();
        uint32_t carry = carry_s;
        uint32_t excl = carry + warp_sums[wid] + (x - v);
        if (idx < nb) offsets[idx] = excl;
        __syncthreads();
        if (tid == 1023) carry_s = excl + v;
        __syncthreads();
    }
    if (tid == 0) { offsets[nb] = carry_s; *total = carry_s; }
}

// Counting-sort scatter: entry (point index | sign) of every non-zero digit goes to its bucket's range.
__global__ void k_scatter(const int32_t* digits, size_t n, unsigned c, unsigned nwin, unsigned groups_per_window, size_t base_off,
                          size_t table_stride, int use_table, const uint32_t* offsets, uint32_t* cursors, uint32_t* entries,
                          uint32_t* keys) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (id >= n * nwin) return;
    int32_t sd = digits[id];
    if (sd == 0) return;
    unsigned w = (unsigned)(id / n);
    size_t i = id - (size_t)w * n;
    const uint32_t B = 1u << (c - 1);
    uint32_t mag = (uint32_t)(sd < 0 ? -sd : sd);
    uint32_t key = (groups_per_window ? w : 0) * B + (mag - 1);
    uint32_t pos = offsets[key] + atomicAdd(&cursors[key], 1u);
    size_t pidx = (use_table ? (size_t)w * table_stride : 0) + base_off + i;
    entries[pos] = (uint32_t)pidx | (sd < 0 ? 0x80000000u : 0u);
    keys[pos] = key;
}

// ---------------------------------------------------------------------------------------------- accumulation, level 0
// Thread t sums sorted entries [t*K, (t+1)*K).  A run of equal keys that lies strictly inside the chunk is a whole
// bucket: stored directly.  A run cut by the chunk boundary goes to partial slot 2t (run touching the chunk start) or
// 2t+1 (run touching the chunk end); a run covering the whole chunk uses slot 2t and an identity in 2t+1 so that equal
// keys stay contiguous in the partial list.
template <class F>
__global__ void __launch_bounds__(128) k_accumulate(const affine_t* __restrict__ points, const uint32_t* __restrict__ entries,
                                                    const uint32_t* __restrict__ keys, const uint32_t* __restrict__ total_ptr,
                                                    xyzz_t* buckets, uint32_t* pkeys, xyzz_t* ppts, uint32_t nthreads) {
    uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= nthreads) return;
    const uint32_t M = *total_ptr;
    const uint64_t lo64 = (uint64_t)t * MSM_CHUNK;
    uint32_t head_key = MSM_KEY_EMPTY, tail_key = MSM_KEY_EMPTY;
    if (lo64 < M) {
        const uint32_t lo = (uint32_t)lo64;
        const uint32_t hi = (M - lo > MSM_CHUNK) ? lo + MSM_CHUNK : M;
        const uint32_t key_prev = lo > 0 ? keys[lo - 1] : MSM_KEY_EMPTY;
        const uint32_t key_next = hi < M ? keys[hi] : MSM_KEY_EMPTY;
        uint32_t cur = keys[lo];
        bool first = true;
        xyzz_t acc = xyzz_identity();
        for (uint32_t i = lo; i < hi; i++) {
            uint32_t k = keys[i];
            if (k != cur) {
                if (first && key_prev == cur) { head_key = cur; store_xyzz(ppts + 2 * (size_t)t, acc); }
                else store_xyzz(buckets + cur, acc);
                first = false;
                cur = k;
                acc = xyzz_identity();
            }
            uint32_t e = entries[i];
            affine_t p = load_affine_nc(points + (e & 0x7fffffffu));
            if (e >> 31) p.y = fe_neg<F>(p.y);
            acc = xyzz_madd<F>(acc, p);
        }
        const bool cont_prev = first && key_prev == cur, cont_next = key_next == cur;
        if (!cont_prev && !cont_next) store_xyzz(buckets + cur, acc);
        else if (first) {
            head_key = cur;
            store_xyzz(ppts + 2 * (size_t)t, acc);
            if (cont_next) { tail_key = cur; store_xyzz(ppts + 2 * (size_t)t + 1, xyzz_identity()); }
        } else {
            tail_key = cur;
            store_xyzz(ppts + 2 * (size_t)t + 1, acc);
        }
    }
    pkeys[2 * (size_t)t] = head_key;
    pkeys[2 * (size_t)t + 1] = tail_key;
}

// ---------------------------------------------------------------------------------------------- partial-list levels
// One warp per 32 consecutive partials: segmented inclusive scan by key with warp shuffles; the last lane of a run holds
// the run's sum.  Runs that do not continue into the neighbouring warps are whole buckets (stored); the others go to
// the next, 16x shorter partial list.
template <class F>
__global__ void __launch_bounds__(128) k_segreduce(const uint32_t* __restrict__ pkeys_in, const xyzz_t* __restrict__ ppts_in, uint32_t n_in,
                                                   xyzz_t* buckets, uint32_t* pkeys_out, xyzz_t* ppts_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t lane = threadIdx.x & 31, wg = i >> 5;
    if (wg * 32 >= n_in) return;  // whole warp out of range
    uint32_t key = i < n_in ? pkeys_in[i] : MSM_KEY_EMPTY;
    xyzz_t pt = xyzz_identity();
    if (key != MSM_KEY_EMPTY) pt = load_xyzz(ppts_in + i);
#pragma unroll 1
    for (unsigned d = 1; d < 32; d <<= 1) {
        uint32_t okey = __shfl_up_sync(0xffffffffu, key, d);
        xyzz_t opt = shfl_up_xyzz(pt, d);
        if (lane >= d && okey == key && key != MSM_KEY_EMPTY) pt = xyzz_add<F>(opt, pt);
    }
    const uint32_t prev_key = __shfl_up_sync(0xffffffffu, key, 1);
    const uint32_t next_key = __shfl_down_sync(0xffffffffu, key, 1);
    const bool is_head = lane == 0 || prev_key != key;
    const uint32_t heads = __ballot_sync(0xffffffffu, is_head);
    const bool run_end = lane == 31 || next_key != key;
    if (lane == 0) { pkeys_out[2 * (size_t)wg] = MSM_KEY_EMPTY; pkeys_out[2 * (size_t)wg + 1] = MSM_KEY_EMPTY; }
    __syncwarp();
    if (run_end && key != MSM_KEY_EMPTY) {
        const uint32_t below = heads & (0xffffffffu >> (31 - lane));
        const uint32_t start = 31 - __clz(below);
        const uint32_t w0 = wg * 32;
        const bool cont_prev = start == 0 && w0 > 0 && pkeys_in[w0 - 1] == key;
        const bool cont_next = lane == 31 && w0 + 32 < n_in && pkeys_in[w0 + 32] == key;
        if (!cont_prev && !cont_next) store_xyzz(buckets + key, pt);
        else if (start == 0) {
            pkeys_out[2 * (size_t)wg] = key;
            store_xyzz(ppts_out + 2 * (size_t)wg, pt);
            if (cont_next) { pkeys_out[2 * (size_t)wg + 1] = key; store_xyzz(ppts_out + 2 * (size_t)wg + 1, xyzz_identity()); }
        } else {
            pkeys_out[2 * (size_t)wg + 1] = key;
            store_xyzz(ppts_out + 2 * (size_t)wg + 1, pt);
        }
    }
}

// ---------------------------------------------------------------------------------------------- bucket reduction
// sum_b (b+1) * B[b] = sum_t 2^t * T_t,  T_t = sum of B[b] over the b with bit t of (b+1) set.
// grid (blocks_per_bit, c, G); every CTA tree-sums its slice of one group's buckets for one bit.
constexpr unsigned BITSUM_THREADS = 128;
template <class F> __device__ __forceinline__ xyzz_t block_sum(xyzz_t acc, xyzz_t* sm) {
    const unsigned lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll 1
    for (unsigned d = 16; d >= 1; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, d);
        if (lane < d) acc = xyzz_add<F>(acc, o);
    }
    if (lane == 0) sm[wid] = acc;
    __syncthreads();
    if (wid == 0) {
        const unsigned nw = blockDim.x >> 5;
        acc = lane < nw ? sm[lane] : xyzz_identity();
#pragma unroll 1
        for (unsigned d = 16; d >= 1; d >>= 1) {
            xyzz_t o = shfl_down_xyzz(acc, d);
            if (lane < d && d < 2 * nw) acc = xyzz_add<F>(acc, o);
        }
    }
    return acc;  // valid in thread 0
}

template <class F>
__global__ void __launch_bounds__(BITSUM_THREADS) k_bitsum(const xyzz_t* __restrict__ buckets, uint32_t B, xyzz_t* partial) {
    __shared__ xyzz_t sm[BITSUM_THREADS / 32];
    const unsigned t = blockIdx.y, g = blockIdx.z, nblk = gridDim.x;
    const uint32_t per = (B + nblk - 1) / nblk;
    const uint32_t b0 = blockIdx.x * per, b1 = min(B, b0 + per);
    xyzz_t acc = xyzz_identity();
    for (uint32_t b = b0 + threadIdx.x; b < b1; b += blockDim.x)
        if (((b + 1) >> t) & 1u) acc = xyzz_add<F>(acc, load_xyzz(buckets + (size_t)g * B + b));
    acc = block_sum<F>(acc, sm);
    if (threadIdx.x == 0) store_xyzz(partial + ((size_t)g * gridDim.y + t) * nblk + blockIdx.x, acc);
}

// second stage: one warp per (group, bit) sums the <= 32 CTA partials
template <class F> __global__ void k_bitsum_final(const xyzz_t* __restrict__ partial, unsigned nblk, xyzz_t* out) {
    const unsigned lane = threadIdx.x;
    xyzz_t acc = lane < nblk ? load_xyzz(partial + (size_t)blockIdx.x * nblk + lane) : xyzz_identity();
#pragma unroll 1
    for (unsigned d = 16; d >= 1; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, d);
        if (lane < d) acc = xyzz_add<F>(acc, o);
    }
    if (lane == 0) store_xyzz(out + blockIdx.x, acc);
}

// ---------------------------------------------------------------------------------------------- workspace
void msm_workspace_free(MsmWorkspace& ws) {
    void* ptrs[] = {ws.d_digits, ws.d_counts, ws.d_offsets, ws.d_entries, ws.d_keys, ws.d_buckets,
                    ws.d_pkeys[0], ws.d_pkeys[1], ws.d_ppts[0], ws.d_ppts[1], ws.d_bitsums, ws.d_total};
    for (void* p : ptrs) if (p) cudaFree(p);
    if (ws.h_bitsums) cudaFreeHost(ws.h_bitsums);
    if (ws.h_total) cudaFreeHost(ws.h_total);
    for (auto& e : ws.ev) if (e) cudaEventDestroy(e);
    ws = MsmWorkspace();
}

template <class F, class FS>
int msm_run(const MsmBases& b, size_t off, size_t n, const fe* d_scalars_in, bool scalars_mont, unsigned c, MsmWorkspace& ws,
            cudaStream_t st, MsmResultShape* shape, unsigned* launches) {
    if (off > b.n || n > b.n - off) { zk_set_error("msm: slice [%zu, %zu) outside the %zu resident bases", off, off + n, b.n); return ZK_ERR_INVALID; }
    shape->c = 0; shape->groups = 0;
    if (n == 0) return ZK_OK;
    const bool use_table = b.c != 0;
    if (use_table) c = b.c;
    else if (c == 0) c = (unsigned)msm_default_window(n, false);
    if (c < 2 || c > MSM_MAX_WINDOW_BITS) { zk_set_error("msm: window bits %u outside [2, %u]", c, MSM_MAX_WINDOW_BITS); return ZK_ERR_INVALID; }
    const unsigned nwin = msm_num_windows(c);
    const unsigned G = use_table ? 1 : nwin;          // bucket groups
    const uint32_t B = 1u << (c - 1);                 // buckets per group
    const size_t NB = (size_t)G * B;
    const size_t Mmax = n * nwin;
    if (Mmax >= 0x7fffffffull || b.n * (size_t)std::max(1u, b.nwin) >= 0x7fffffffull) { zk_set_error("msm: %zu x %u entries exceed the 31-bit index space", n, nwin); return ZK_ERR_INVALID; }
    const uint32_t T0 = (uint32_t)((Mmax + MSM_CHUNK - 1) / MSM_CHUNK);
    const unsigned nblk = (unsigned)std::min<size_t>(32, std::max<size_t>(1, B / (BITSUM_THREADS * 2)));

    // scratch (grown on demand, reused across calls)
    static_assert(sizeof(xyzz_t) == 128 && sizeof(affine_t) == 64 && sizeof(fe) == 32, "layout");
    // buffers are grouped by what sizes them: the entry list (n * nwin), the bucket array (G * B), the bit sums (G * c)
    {
        size_t need_digits = Mmax * sizeof(int32_t), need_counts = NB * sizeof(uint32_t), need_offsets = (NB + 1) * sizeof(uint32_t);
        size_t need_entries = Mmax * sizeof(uint32_t), need_buckets = NB * sizeof(xyzz_t);
        size_t need_pk = 2 * (size_t)T0 * sizeof(uint32_t), need_pp = 2 * (size_t)T0 * sizeof(xyzz_t);
        size_t need_bits = (size_t)G * c * nblk * sizeof(xyzz_t) + (size_t)G * c * sizeof(xyzz_t);
        if (ws.cap_entries < need_entries) {
            if (ws.d_digits) cudaFree(ws.d_digits);
            if (ws.d_entries) cudaFree(ws.d_entries);
            if (ws.d_keys) cudaFree(ws.d_keys);
            for (int k = 0; k < 2; k++) { if (ws.d_pkeys[k]) cudaFree(ws.d_pkeys[k]); if (ws.d_ppts[k]) cudaFree(ws.d_ppts[k]); }
            ws.d_digits = nullptr; ws.d_entries = ws.d_keys = nullptr; ws.d_pkeys[0] = ws.d_pkeys[1] = nullptr; ws.d_ppts[0] = ws.d_ppts[1] = nullptr;
            ws.cap_entries = 0;
            ZK_CUDA(cudaMalloc(&ws.d_digits, need_digits));
            ZK_CUDA(cudaMalloc(&ws.d_entries, need_entries));
            ZK_CUDA(cudaMalloc(&ws.d_keys, need_entries));
            ZK_CUDA(cudaMalloc(&ws.d_pkeys[0], need_pk));
            ZK_CUDA(cudaMalloc(&ws.d_ppts[0], need_pp));
            size_t n1 = 2 * (((size_t)2 * T0 + 31) / 32);
            ZK_CUDA(cudaMalloc(&ws.d_pkeys[1], std::max<size_t>(n1, 2) * sizeof(uint32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_ppts[1], std::max<size_t>(n1, 2) * sizeof(xyzz_t)));
            ws.cap_entries = need_entries;
        }
        if (ws.cap_buckets < need_buckets) {
            if (ws.d_counts) cudaFree(ws.d_counts);
            if (ws.d_offsets) cudaFree(ws.d_offsets);
            if (ws.d_buckets) cudaFree(ws.d_buckets);
            ws.d_counts = ws.d_offsets = nullptr; ws.d_buckets = nullptr; ws.cap_buckets = 0;
            ZK_CUDA(cudaMalloc(&ws.d_counts, need_counts));
            ZK_CUDA(cudaMalloc(&ws.d_offsets, need_offsets));
            ZK_CUDA(cudaMalloc(&ws.d_buckets, need_buckets));
            ws.cap_buckets = need_buckets;
        }
        if (ws.cap_partials < need_bits) {
            if (ws.d_bitsums) cudaFree(ws.d_bitsums);
            ws.d_bitsums = nullptr; ws.cap_partials = 0;
            ZK_CUDA(cudaMalloc(&ws.d_bitsums, need_bits));
            ws.cap_partials = need_bits;
        }
        if (ws.cap_hbits < (size_t)G * c) {   // sized by G*c alone: few wide groups and many narrow ones differ
            if (ws.h_bitsums) cudaFreeHost(ws.h_bitsums);
            ws.h_bitsums = nullptr; ws.cap_hbits = 0;
            ZK_CUDA(cudaMallocHost(&ws.h_bitsums, (size_t)G * c * sizeof(xyzz_t)));
            ws.cap_hbits = (size_t)G * c;
        }
        if (!ws.d_total) {
            ZK_CUDA(cudaMalloc(&ws.d_total, sizeof(uint32_t)));
            ZK_CUDA(cudaMallocHost(&ws.h_total, sizeof(uint32_t)));
        }
    }
    unsigned nl = 0;
    if (ws.profile && !ws.ev[0])
        for (int k = 0; k <= MSM_ST_COUNT; k++) ZK_CUDA(cudaEventCreate(&ws.ev[k]));
#define STAGE_MARK(k) do { if (ws.profile) ZK_CUDA(cudaEventRecord(ws.ev[k], st)); } while (0)

    // 1. digits + histogram
    ZK_CUDA(cudaMemsetAsync(ws.d_counts, 0, NB * sizeof(uint32_t), st));
    ZK_CUDA(cudaMemsetAsync(ws.d_buckets, 0, NB * sizeof(xyzz_t), st));  // all-zero XYZZ == identity
    STAGE_MARK(0);
    k_recode<FS><<<(unsigned)((n + 127) / 128), 128, 0, st>>>(d_scalars_in, scalars_mont ? 1 : 0, n, c, nwin, use_table ? 0 : 1, ws.d_digits, ws.d_counts);
    STAGE_MARK(1);
    // 2. bucket offsets
    k_scan<<<1, 1024, 0, st>>>(ws.d_counts, ws.d_offsets, (uint32_t)NB, ws.d_total);
    STAGE_MARK(2);
    // 3. scatter (counting sort by bucket)
    k_scatter<<<(unsigned)((Mmax + 255) / 256), 256, 0, st>>>(ws.d_digits, n, c, nwin, use_table ? 0 : 1, off, b.n, use_table ? 1 : 0,
                                                            ws.d_offsets, ws.d_counts, ws.d_entries, ws.d_keys);
    STAGE_MARK(3);
    // 4. balanced accumulation over the sorted list
    k_accumulate<F><<<(T0 + 127) / 128, 128, 0, st>>>(b.d_points, ws.d_entries, ws.d_keys, ws.d_total, ws.d_buckets, ws.d_pkeys[0], ws.d_ppts[0], T0);
    nl += 4;
    STAGE_MARK(4);
    // 5. partial-list levels
    uint32_t n_in = 2 * T0;
    int cur = 0;
    for (;;) {
        uint32_t warps = (n_in + 31) / 32;
        k_segreduce<F><<<(warps * 32 + 127) / 128, 128, 0, st>>>(ws.d_pkeys[cur], ws.d_ppts[cur], n_in, ws.d_buckets, ws.d_pkeys[cur ^ 1], ws.d_ppts[cur ^ 1]);
        nl++;
        if (warps == 1) break;
        n_in = 2 * warps;
        cur ^= 1;
    }
    STAGE_MARK(5);
    // 6. bit-sliced bucket sums
    xyzz_t* d_partial = ws.d_bitsums;
    xyzz_t* d_T = ws.d_bitsums + (size_t)G * c * nblk;
    k_bitsum<F><<<dim3(nblk, c, G), BITSUM_THREADS, 0, st>>>(ws.d_buckets, B, d_partial);
    k_bitsum_final<F><<<G * c, 32, 0, st>>>(d_partial, nblk, d_T);
    nl += 2;
    STAGE_MARK(6);
    ZK_CUDA(cudaGetLastError());
    ZK_CUDA(cudaMemcpyAsync(ws.h_bitsums, d_T, (size_t)G * c * sizeof(xyzz_t), cudaMemcpyDeviceToHost, st));
    ZK_CUDA(cudaStreamSynchronize(st));
    if (ws.profile)
        for (int k = 0; k < MSM_ST_COUNT; k++) ZK_CUDA(cudaEventElapsedTime(&ws.stage_ms[k], ws.ev[k], ws.ev[k + 1]));
#undef STAGE_MARK
    if (launches) *launches += nl;
    // the O(c) serial tail (c doublings per group) is finished on the host from ws.h_bitsums (api.cu: msm_finish)
    shape->c = c; shape->groups = G;
    return ZK_OK;
}

#define INST(F, FS)                                                                                                             \
    template int msm_bases_create<F>(MsmBases&, const affine_t*, bool, size_t, unsigned, cudaStream_t);                          \
    template int msm_run<F, FS>(const MsmBases&, size_t, size_t, const fe*, bool, unsigned, MsmWorkspace&, cudaStream_t, MsmResultShape*, unsigned*);
INST(FpParams, FqParams)  // Pallas: coordinates Fp, scalars Fq
INST(FqParams, FpParams)  // Vesta:  coordinates Fq, scalars Fp
#undef INST

}  // namespace zkb
