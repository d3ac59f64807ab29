// msm.cu — kernels and pipeline of the Pallas/Vesta MSM (see msm.cuh for semantics and reference call sites).
#include "msm.cuh"
#include "quad.cuh"

#include <algorithm>

namespace zkb {

unsigned msm_num_windows(unsigned c) { return (256 + c - 1) / c; }  // top signed digit cannot carry out (scalars < 2^254 + 2^127)

int msm_default_window(size_t n, bool precomputed) {
    // accumulation costs ceil(256/c) * n mixed additions; the bucket side grows with 2^(c-1) (one bucket set with a
    // table, one per window without).  Measured on B200 (tools/sweep.py): 2^16 points -> c = 15 with a table, 12 without.
    unsigned l = 0;
    while (((size_t)1 << (l + 1)) <= n) l++;
    int c = (int)l - (precomputed ? 1 : 4);
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
// One thread per scalar of one MSM of the batch (blockIdx.y): Montgomery -> canonical if asked (VariableBaseMSM::msm ==
// into_bigint + msm_bigint), then signed base-2^c digits d_w in (-2^(c-1), 2^(c-1)]; digit w of scalar i of MSM j at
// digits[(j*nwin + w)*n + i]; per-bucket counts.  Bucket group of (MSM j, window w): j*gpm + (per_window ? w : 0).
template <class FS>
__global__ void k_recode(MsmScalarSet sc, int scalars_mont, size_t n, unsigned c, unsigned nwin, unsigned gpm, int per_window,
                         int32_t* digits, uint32_t* counts, uint32_t* meta) {
    if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) { meta[2] = 0; meta[4] = 0; }   // giants, plan tiles done
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const unsigned j = blockIdx.y;
    fe s = load_fe_nc(sc.p[j] + i);
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
        digits[((size_t)j * nwin + w) * n + i] = sd;
        // warp-aggregated histogram update: lanes that hit the same bucket (kimchi's all-ones columns: all of them) elect
        // one lane to add their count — one atomic per distinct bucket per warp instead of one per lane
        const uint32_t mag = (uint32_t)(sd < 0 ? -sd : sd);
        const uint32_t key = sd != 0 ? (j * gpm + (per_window ? w : 0)) * B + (mag - 1) : 0xffffffffu;
        const uint32_t peers = __match_any_sync(__activemask(), key);
        if (sd != 0 && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&counts[key], (uint32_t)__popc(peers));
    }
}

// Planning pass over the histogram: exclusive scans of the bucket counts (-> offsets of the sorted entry list) and of the
// per-bucket task counts s_b = ceil(n_b / K) (-> task_off).  One CTA per tile of 4096 buckets; the tiles are chained (tile t
// spins on tile t-1's published inclusive total — tiles are dispatched in order, so the predecessor is always resident), which
// keeps the scan a single launch for any batch size.  64-bit lanes carry (entry count << 32 | task count): one scan yields both.
// Buckets with more than smax tasks ("giant": the all-ones witness columns of kimchi put ~n entries in one bucket,
// SURVEY.md §3.1) are listed; the last tile to finish lays out their side areas behind the main slots.  Clears counts
// (re-used as scatter cursors).
__device__ __forceinline__ uint64_t block_excl_scan_1024(uint64_t v, uint64_t* warp_sums, uint64_t* tile_total) {
    const unsigned tid = threadIdx.x, lane = tid & 31, wid = tid >> 5;
    uint64_t x = v;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
        uint64_t y = __shfl_up_sync(0xffffffffu, x, d);
        if (lane >= (unsigned)d) x += y;
    }
    if (lane == 31) warp_sums[wid] = x;
    __syncthreads();
    if (wid == 0) {
        uint64_t ws = warp_sums[lane], z = ws;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint64_t y = __shfl_up_sync(0xffffffffu, z, d);
            if (lane >= (unsigned)d) z += y;
        }
        warp_sums[lane] = z - ws;
        if (lane == 31) *tile_total = z;
    }
    __syncthreads();
    uint64_t excl = warp_sums[wid] + (x - v);
    __syncthreads();
    return excl;
}

constexpr unsigned PLAN_PER_THREAD = 4, PLAN_TILE = 1024 * PLAN_PER_THREAD;
__global__ void __launch_bounds__(1024) k_plan(uint32_t* counts, uint32_t* offsets, uint32_t* task_off, uint32_t nb, uint32_t K, uint32_t smax,
                                                 uint32_t* meta, uint32_t* giants, uint64_t* chain, uint32_t* chain_flag, uint32_t epoch) {
    __shared__ uint64_t warp_sums[32];
    __shared__ uint64_t tile_total, prefix_s;
    __shared__ uint32_t last_s;
    const unsigned tid = threadIdx.x, tile = blockIdx.x;
    const uint32_t i0 = tile * PLAN_TILE + tid * PLAN_PER_THREAD;
    uint32_t v[PLAN_PER_THREAD];
    uint64_t pk[PLAN_PER_THREAD], sum = 0;
#pragma unroll
    for (unsigned k = 0; k < PLAN_PER_THREAD; k++) {
        v[k] = i0 + k < nb ? counts[i0 + k] : 0;
        if (i0 + k < nb) counts[i0 + k] = 0;
        pk[k] = ((uint64_t)v[k] << 32) | ((v[k] + K - 1) / K);
        sum += pk[k];
    }
    uint64_t ex = block_excl_scan_1024(sum, warp_sums, &tile_total);
    if (tid == 0) {
        uint64_t prefix = 0;
        if (tile > 0) {
            volatile uint32_t* f = chain_flag + (tile - 1);
            while (*f != epoch) { }
            __threadfence();
            prefix = *(volatile uint64_t*)(chain + (tile - 1));
        }
        *(volatile uint64_t*)(chain + tile) = prefix + tile_total;
        __threadfence();
        *(volatile uint32_t*)(chain_flag + tile) = epoch;
        prefix_s = prefix;
    }
    __syncthreads();
    ex += prefix_s;
#pragma unroll
    for (unsigned k = 0; k < PLAN_PER_THREAD; k++) {
        if (i0 + k < nb) {
            offsets[i0 + k] = (uint32_t)(ex >> 32);
            task_off[i0 + k] = (uint32_t)ex;
            if ((uint32_t)pk[k] > smax) {
                uint32_t gi = atomicAdd(&meta[2], 1u);
                if (gi < MSM_MAX_GIANTS) { giants[gi] = i0 + k; giants[MSM_MAX_GIANTS + gi] = (uint32_t)pk[k]; }
            }
        }
        ex += pk[k];
    }
    if (tile == gridDim.x - 1 && tid == 0) {
        const uint64_t tot = prefix_s + tile_total;
        offsets[nb] = (uint32_t)(tot >> 32); task_off[nb] = (uint32_t)tot;
        meta[0] = (uint32_t)(tot >> 32); meta[1] = (uint32_t)tot;
    }
    // the last tile to finish sees the complete giant list: main slots first, then one side area per listed giant
    __threadfence();
    __syncthreads();
    if (tid == 0) last_s = atomicAdd(&meta[4], 1u) == gridDim.x - 1 ? 1u : 0u;
    __syncthreads();
    if (last_s && tid == 0) {
        __threadfence();
        const uint32_t ng = min(*(volatile uint32_t*)(meta + 2), MSM_MAX_GIANTS);
        const uint32_t nt = (uint32_t)(*(volatile uint64_t*)(chain + (gridDim.x - 1)));
        uint32_t excess = 0;
        for (uint32_t r = 0; r < ng; r++) excess += *(volatile uint32_t*)(giants + MSM_MAX_GIANTS + r) - 1;
        uint32_t base = nt - excess;
        meta[3] = base;
        for (uint32_t r = 0; r < ng; r++) { giants[2 * MSM_MAX_GIANTS + r] = base; base += *(volatile uint32_t*)(giants + MSM_MAX_GIANTS + r); }
    }
}

// Counting-sort scatter: entry (point index | sign) of every non-zero digit goes to its bucket's range.
__global__ void k_scatter(const int32_t* digits, size_t n, unsigned c, unsigned nwin, unsigned batch, unsigned gpm, int per_window, size_t base_off,
                          size_t table_stride, int use_table, size_t n_main, size_t n_extra, uint32_t main_count, const uint32_t* offsets,
                          uint32_t* cursors, uint32_t* entries) {
    size_t id = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const bool in_range = id < n * nwin * batch;
    int32_t sd = in_range ? digits[id] : 0;
    const unsigned jw = (unsigned)(id / n), j = jw / nwin, w = jw - j * nwin;
    size_t i = id - (size_t)jw * n;
    const uint32_t B = 1u << (c - 1);
    uint32_t mag = (uint32_t)(sd < 0 ? -sd : sd);
    uint32_t key = sd != 0 ? (j * gpm + (per_window ? w : 0)) * B + (mag - 1) : 0xffffffffu;
    // warp-aggregated cursor bump (see k_recode): the leader reserves a run, every peer takes its rank inside it
    const uint32_t peers = __match_any_sync(0xffffffffu, key);
    const unsigned lane = threadIdx.x & 31, leader = (unsigned)(__ffs(peers) - 1);
    uint32_t base = 0;
    if (sd != 0 && lane == leader) base = atomicAdd(&cursors[key], (uint32_t)__popc(peers));
    base = __shfl_sync(0xffffffffu, base, leader);
    if (sd == 0) return;
    uint32_t pos = offsets[key] + base + (uint32_t)__popc(peers & ((1u << lane) - 1));
    // scalars [0, n_main) belong to the resident bases, [n_main, n) to the call's extra points (their rows behind the table)
    size_t pidx = i < n_main ? (use_table ? (size_t)w * table_stride : 0) + base_off + i
                             : (size_t)main_count + (use_table ? (size_t)w * n_extra : 0) + (i - n_main);
    entries[pos] = (uint32_t)pidx | (sd < 0 ? 0x80000000u : 0u);
}

// ---------------------------------------------------------------------------------------------- slots of the partial list
// Task t of bucket b leaves its partial in a SLOT.  Main slots follow bucket order: bucket b owns [slot_of(b), slot_of(b+1));
// that is its task range, except that a listed giant owns ONE main slot (its total, written by k_giant_finish) and keeps its
// task partials in a side area behind the main slots.  With no giants (uniform scalars) slot == task index.
struct GiantList {
    const uint32_t* bucket;   // [ng]
    const uint32_t* tasks;    // [ng]
    const uint32_t* area;     // [ng] first slot of the side area
    uint32_t ng;
};
__device__ __forceinline__ GiantList giant_list(const uint32_t* __restrict__ meta, const uint32_t* __restrict__ giants) {
    GiantList g;
    g.ng = min(__ldg(meta + 2), MSM_MAX_GIANTS);
    g.bucket = giants; g.tasks = giants + MSM_MAX_GIANTS; g.area = giants + 2 * MSM_MAX_GIANTS;
    return g;
}
__device__ __forceinline__ uint32_t slot_of(const uint32_t* __restrict__ task_off, uint32_t b, const GiantList& g) {
    uint32_t s = __ldg(task_off + b);
    for (uint32_t r = 0; r < g.ng; r++)
        if (__ldg(g.bucket + r) < b) s -= __ldg(g.tasks + r) - 1;
    return s;
}

// ---------------------------------------------------------------------------------------------- accumulation
// Task t belongs to the bucket b with task_off[b] <= t < task_off[b+1]; the bucket's n_b sorted entries are cut into
// s_b = ceil(n_b / K) nearly equal parts, so no task crosses a bucket boundary and every thread sums <= K points with
// XYZZ mixed additions and leaves one partial in its slot.
template <class F>
__global__ void __launch_bounds__(128) k_accumulate(const affine_t* __restrict__ points, const uint32_t* __restrict__ entries,
                                                    const uint32_t* __restrict__ offsets, const uint32_t* __restrict__ task_off, uint32_t nb,
                                                    uint32_t K, const uint32_t* __restrict__ meta, const uint32_t* __restrict__ giants,
                                                    const affine_t* __restrict__ extra, uint32_t main_count, xyzz_t* partials) {
    const uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= __ldg(meta + 1)) return;
    // upper_bound(task_off[0..nb], t) - 1
    uint32_t lo = 0, hi = nb;  // invariant: task_off[lo] <= t < task_off[hi]
    while (hi - lo > 1) {
        uint32_t mid = (lo + hi) >> 1;
        if (__ldg(task_off + mid) <= t) lo = mid; else hi = mid;
    }
    const uint32_t b = lo;
    const uint32_t e0 = __ldg(offsets + b), nbk = __ldg(offsets + b + 1) - e0;
    const uint32_t sb = (nbk + K - 1) / K, sub = t - __ldg(task_off + b);
    const uint32_t base = nbk / sb, rem = nbk - base * sb;
    uint32_t i = e0 + sub * base + min(sub, rem);
    const uint32_t end = i + base + (sub < rem ? 1u : 0u);
    xyzz_t acc = xyzz_identity();
    // software pipeline: the gather of point i+1 is in flight while point i is added
    // (entries >= main_count address the call's extra points: h and U of an IPA round, ipa.rs:944,954)
    uint32_t e = __ldg(entries + i);
    uint32_t idx = e & 0x7fffffffu;
    affine_t p = load_affine_nc(idx < main_count ? points + idx : extra + (idx - main_count));
    for (; i < end; i++) {
        affine_t q = p;
        const uint32_t sign = e >> 31;
        if (i + 1 < end) {
            e = __ldg(entries + i + 1);
            idx = e & 0x7fffffffu;
            p = load_affine_nc(idx < main_count ? points + idx : extra + (idx - main_count));
        }
        if (sign) q.y = fe_neg<F>(q.y);
        acc = xyzz_madd<F>(acc, q);
    }
    uint32_t dest = t;
    const GiantList g = giant_list(meta, giants);
    if (g.ng) {
        uint32_t excess = 0, mine = 0xffffffffu;
        for (uint32_t r = 0; r < g.ng; r++) {
            const uint32_t gb = __ldg(g.bucket + r);
            if (gb < b) excess += __ldg(g.tasks + r) - 1;
            else if (gb == b) mine = r;
        }
        dest = mine != 0xffffffffu ? __ldg(g.area + mine) + sub : t - excess;
    }
    store_xyzz(partials + dest, acc);
}

// ---------------------------------------------------------------------------------------------- sums of partials
constexpr unsigned TREE_THREADS = 256;              // 64 quads per CTA in the giant / final kernels
constexpr unsigned TREE_QUADS = TREE_THREADS / 4;
constexpr unsigned ROWCOL_THREADS = 128;            // 32 quads per row / column CTA: the whole grid of a single MSM is one wave

// Block-wide sum of one value per QUAD (replicated in its four lanes; quad.cuh): three shuffle levels inside every warp, one
// shared-memory hand-over, then the warps' sums in warp 0.  One barrier instead of one per level.  blockDim.x a multiple of 32,
// <= 1024; sm holds blockDim.x / 32 points.  Result valid in quad 0 of warp 0.
template <class F> __device__ __forceinline__ xyzz_t block_sum_quads(xyzz_t acc, xyzz_t* sm) {
    const unsigned tid = threadIdx.x, lane = tid & 31, qw = lane >> 2, wid = tid >> 5, nw = blockDim.x >> 5;
#pragma unroll 1
    for (unsigned d = 4; d >= 1; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, 4 * d);
        if (qw >= d) o = xyzz_identity();
        acc = xyzz_add_quad<F>(acc, o);
    }
    if (nw == 1) return acc;
    if (lane == 0) store_xyzz(sm + wid, acc);
    __syncthreads();
    if (wid != 0) return acc;
    acc = qw < nw ? load_xyzz(sm + qw) : xyzz_identity();
    unsigned top = 1;
    while (top < nw) top <<= 1;      // nw <= 8 here; a 32-warp block would need a second hand-over
#pragma unroll 1
    for (unsigned d = top >> 1; d >= 1; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, 4 * d);
        if (qw >= d) o = xyzz_identity();
        acc = xyzz_add_quad<F>(acc, o);
    }
    return acc;
}

// GIANT_SLICES CTAs per listed giant bucket: each sums a contiguous slice of the giant's side area (quads stride over it, then
// the block sum); the last CTA to arrive (ticket counter) adds the slice sums and writes the giant's main slot.
constexpr unsigned GIANT_SLICES = 16, GIANT_GRID = 32;
template <class F>
__global__ void __launch_bounds__(TREE_THREADS) k_giant_finish(const uint32_t* __restrict__ giants, const uint32_t* __restrict__ meta,
                                                               const uint32_t* __restrict__ task_off, xyzz_t* partials, xyzz_t* slice_sums,
                                                               uint32_t* tickets) {
    __shared__ xyzz_t sm_tree[TREE_THREADS / 32];
    __shared__ uint32_t ticket_s;
    const GiantList g = giant_list(meta, giants);
    const unsigned qd = threadIdx.x >> 2, nq = blockDim.x >> 2;
    for (uint32_t r = blockIdx.x; r < g.ng; r += gridDim.x) {
        const uint32_t sb = __ldg(g.tasks + r), t0 = __ldg(g.area + r);
        const uint32_t per = (sb + GIANT_SLICES - 1) / GIANT_SLICES;
        const uint32_t j_lo = min(sb, blockIdx.y * per), j_hi = min(sb, j_lo + per);
        xyzz_t acc = xyzz_identity();
        for (uint32_t j0 = j_lo; j0 < j_hi; j0 += nq) {
            xyzz_t o = j0 + qd < j_hi ? load_xyzz(partials + t0 + j0 + qd) : xyzz_identity();
            acc = xyzz_add_quad<F>(acc, o);
        }
        acc = block_sum_quads<F>(acc, sm_tree);
        xyzz_t* mine = slice_sums + (size_t)r * GIANT_SLICES;
        if (threadIdx.x == 0) {
            store_xyzz(mine + blockIdx.y, acc);
            __threadfence();
            ticket_s = atomicAdd(&tickets[r], 1u);
        }
        __syncthreads();
        if (ticket_s == GIANT_SLICES - 1) {
            __threadfence();
            acc = qd < GIANT_SLICES ? load_xyzz(mine + qd) : xyzz_identity();
            __syncthreads();   // sm_tree is reused
            acc = block_sum_quads<F>(acc, sm_tree);
            if (threadIdx.x == 0) { store_xyzz(partials + slot_of(task_off, __ldg(g.bucket + r), g), acc); tickets[r] = 0; }
        }
        __syncthreads();
    }
}

// Row and column sums of the (B/W + 1) x W grid of bucket weights i = b + 1 = hi * W + lo, W = 2^w_lo, straight from the
// partial list:   sum_i i * B[i] = W * sum_hi hi * R[hi] + sum_lo lo * C[lo],   R[hi] = sum of the partials of row hi's buckets (one
// contiguous slot range), C[lo] = sum of the partials of the buckets hi * W + lo over hi (their slot ranges walked as one flat
// list, so the quads of a CTA share the work evenly whatever the bucket sizes).  One CTA per row and per column and group;
// about 2 additions per partial in total, <= ceil(slots / quads) dependent additions per quad plus the block sum.
constexpr unsigned ROWCOL_MAX_ROWS = 260;           // B / W + 1 <= 257 for c <= 17
template <class F>
__global__ void __launch_bounds__(ROWCOL_THREADS) k_rowcol(const xyzz_t* __restrict__ partials, const uint32_t* __restrict__ task_off,
                                                           const uint32_t* __restrict__ meta, const uint32_t* __restrict__ giants, uint32_t B,
                                                           unsigned w_lo, xyzz_t* rc) {
    __shared__ xyzz_t sm_tree[ROWCOL_THREADS / 32];
    __shared__ uint32_t s_start[ROWCOL_MAX_ROWS], s_pref[ROWCOL_MAX_ROWS + 1];
    __shared__ uint32_t s_warp[ROWCOL_THREADS / 32];
    const uint32_t W = 1u << w_lo, nrows = (B >> w_lo) + 1;
    const unsigned g = blockIdx.z, tid = threadIdx.x, qd = tid >> 2, nq = blockDim.x >> 2;
    const uint32_t gb0 = g * B;
    const GiantList gl = giant_list(meta, giants);
    xyzz_t acc = xyzz_identity();
    if (blockIdx.x < nrows) {
        const uint32_t hi = blockIdx.x;
        const uint32_t i_first = max(hi * W, 1u), i_last = min(hi * W + W - 1, B);
        const uint32_t s0 = slot_of(task_off, gb0 + i_first - 1, gl), s1 = slot_of(task_off, gb0 + i_last, gl);
        // software pipeline: the next partial is in flight while the current one is added
        xyzz_t cur = s0 + qd < s1 ? load_xyzz(partials + s0 + qd) : xyzz_identity();
        for (uint32_t s = s0; s < s1; s += nq) {
            const uint32_t nx = s + nq + qd;
            xyzz_t nxt = nx < s1 ? load_xyzz(partials + nx) : xyzz_identity();
            acc = xyzz_add_quad<F>(acc, cur);
            cur = nxt;
        }
    } else {
        const uint32_t lo = blockIdx.x - nrows;
        // slot range of every bucket of the column, then an exclusive scan of the range lengths (nrows <= 3 * blockDim.x)
        uint32_t cnt[3], mysum = 0;
#pragma unroll
        for (unsigned k = 0; k < 3; k++) {
            const uint32_t hi = tid * 3 + k;
            cnt[k] = 0;
            if (hi < nrows) {
                const uint32_t i = hi * W + lo;
                uint32_t st = 0;
                if (i >= 1 && i <= B) {
                    st = slot_of(task_off, gb0 + i - 1, gl);
                    cnt[k] = slot_of(task_off, gb0 + i, gl) - st;
                }
                s_start[hi] = st;
            }
            mysum += cnt[k];
        }
        uint32_t x = mysum;
        const unsigned lane = tid & 31, wid = tid >> 5;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
            uint32_t y = __shfl_up_sync(0xffffffffu, x, d);
            if (lane >= (unsigned)d) x += y;
        }
        if (lane == 31) s_warp[wid] = x;
        __syncthreads();
        uint32_t wbase = 0;
        for (unsigned k = 0; k < wid; k++) wbase += s_warp[k];
        uint32_t ex = wbase + x - mysum;
#pragma unroll
        for (unsigned k = 0; k < 3; k++) {
            const uint32_t hi = tid * 3 + k;
            if (hi < nrows) s_pref[hi] = ex;
            ex += cnt[k];
        }
        if (tid == blockDim.x - 1) s_pref[nrows] = ex;
        __syncthreads();
        const uint32_t total = s_pref[nrows];
        auto fetch = [&](uint32_t f) -> xyzz_t {
            if (f >= total) return xyzz_identity();
            uint32_t l = 0, h = nrows;      // s_pref[l] <= f < s_pref[h]
            while (h - l > 1) {
                const uint32_t m = (l + h) >> 1;
                if (s_pref[m] <= f) l = m; else h = m;
            }
            return load_xyzz(partials + s_start[l] + (f - s_pref[l]));
        };
        xyzz_t cur = fetch(qd);
        for (uint32_t f0 = 0; f0 < total; f0 += nq) {
            xyzz_t nxt = fetch(f0 + nq + qd);
            acc = xyzz_add_quad<F>(acc, cur);
            cur = nxt;
        }
    }
    acc = block_sum_quads<F>(acc, sm_tree);
    if (tid == 0) store_xyzz(rc + (size_t)g * (nrows + W) + blockIdx.x, acc);
}

// One CTA per output bit and group: T_t = sum of C[lo] over bit t of lo (t < w_lo), of R[hi] over bit t - w_lo of hi (t >= w_lo).
// c points per group; the group's value is sum_t 2^t T_t (finished on the host: c doublings).
template <class F>
__global__ void __launch_bounds__(TREE_THREADS) k_bit_slices(const xyzz_t* __restrict__ rc, uint32_t B, unsigned w_lo, unsigned c, xyzz_t* out) {
    __shared__ xyzz_t sm_tree[TREE_THREADS / 32];
    const uint32_t W = 1u << w_lo, nrows = (B >> w_lo) + 1;
    const unsigned t = blockIdx.x, g = blockIdx.y;
    const bool cols = t < w_lo;
    const xyzz_t* src = rc + (size_t)g * (nrows + W) + (cols ? nrows : 0);
    const uint32_t count = cols ? W : nrows;
    const unsigned bit = cols ? t : t - w_lo;
    const unsigned qd = threadIdx.x >> 2, nq = blockDim.x >> 2;
    xyzz_t acc = xyzz_identity();
    // the j-th index with `bit` set: e = (j >> bit) << (bit + 1) | 1 << bit | (j & (2^bit - 1))
    for (uint32_t j0 = 0;; j0 += nq) {
        const uint32_t first = ((j0 >> bit) << (bit + 1)) | (1u << bit) | (j0 & ((1u << bit) - 1));
        if (first >= count) break;     // uniform over the block
        const uint32_t j = j0 + qd, e = ((j >> bit) << (bit + 1)) | (1u << bit) | (j & ((1u << bit) - 1));
        xyzz_t o = e < count ? load_xyzz(src + e) : xyzz_identity();
        acc = xyzz_add_quad<F>(acc, o);
    }
    acc = block_sum_quads<F>(acc, sm_tree);
    if (threadIdx.x == 0) store_xyzz(out + (size_t)g * c + t, acc);
}

// ---------------------------------------------------------------------------------------------- workspace
static void free_dev(void* p) { if (p) cudaFree(p); }

void msm_workspace_free(MsmWorkspace& ws) {
    free_dev(ws.d_digits); free_dev(ws.d_entries); free_dev(ws.d_partials);
    free_dev(ws.d_counts); free_dev(ws.d_offsets); free_dev(ws.d_task_off); free_dev(ws.d_chain); free_dev(ws.d_chain_flag);
    free_dev(ws.d_bitsums); free_dev(ws.d_meta); free_dev(ws.d_giants); free_dev(ws.d_giant_slices); free_dev(ws.d_giant_tickets);
    if (ws.h_bitsums) cudaFreeHost(ws.h_bitsums);
    for (auto& e : ws.ev) if (e) cudaEventDestroy(e);
    ws = MsmWorkspace();
}

template <class F, class FS>
int msm_run(const MsmBases& b, size_t off, size_t n_main, const fe* const* d_scalars, unsigned k, bool scalars_mont, unsigned c,
            MsmWorkspace& ws, cudaStream_t st, MsmResultShape* shape, unsigned* launches, const affine_t* d_extra, size_t n_extra) {
    const size_t n = n_main + n_extra;                // scalars per MSM
    if (n_extra && !d_extra) { zk_set_error("msm: extra points missing"); return ZK_ERR_INVALID; }
    if (off > b.n || n_main > b.n - off) { zk_set_error("msm: slice [%zu, %zu) outside the %zu resident bases", off, off + n_main, b.n); return ZK_ERR_INVALID; }
    if (k == 0 || k > MSM_MAX_BATCH) { zk_set_error("msm: batch of %u outside [1, %u]", k, MSM_MAX_BATCH); return ZK_ERR_INVALID; }
    shape->c = 0; shape->groups = 0; shape->batch = k;
    if (n == 0) return ZK_OK;
    const bool use_table = b.c != 0;
    if (use_table) c = b.c;
    else if (c == 0) c = (unsigned)msm_default_window(n, false);
    if (c < 2 || c > MSM_MAX_WINDOW_BITS) { zk_set_error("msm: window bits %u outside [2, %u]", c, MSM_MAX_WINDOW_BITS); return ZK_ERR_INVALID; }
    const unsigned nwin = msm_num_windows(c);
    const unsigned gpm = use_table ? 1 : nwin;        // bucket groups per MSM
    const unsigned G = k * gpm;                       // bucket groups of the batch
    const uint32_t B = 1u << (c - 1);                 // buckets per group
    const size_t NB = (size_t)G * B;
    const size_t Mmax = n * nwin * (size_t)k;
    const size_t main_count = b.n * (size_t)std::max(1u, b.nwin);   // table entries; the extra points' rows are addressed behind them
    if (Mmax >= 0x7fffffffull || NB >= 0x7fffffffull || main_count + n_extra * nwin >= 0x7fffffffull) {
        zk_set_error("msm: %zu x %u x %u entries exceed the 31-bit index space", n, nwin, k);
        return ZK_ERR_INVALID;
    }
    // Entries per accumulation task: the tasks fill the machine ONCE (wave_threads accumulation threads per SM; the mixed
    // addition rate saturates from ~8 warps per SM, tools/microbench.py) — every further task is one more partial for the
    // latency-bound reduction behind the accumulation.  Very large inputs get several waves of <= 256-entry tasks.
    const size_t wave_threads = ws.wave_threads ? ws.wave_threads : 384;
    const size_t capacity = (size_t)ws.sm_count * wave_threads;
    uint32_t K = ws.chunk;
    if (K == 0) {
        const size_t slack = std::min<size_t>(NB / 2, capacity / 4);     // sum_b ceil(n_b/K) ~ M/K + (non-empty buckets)/2
        size_t waves = (Mmax + 256 * capacity - 1) / (256 * capacity);
        if (waves == 0) waves = 1;
        K = (uint32_t)((Mmax + waves * capacity - slack - 1) / (waves * capacity - slack));
        if (K < 4) K = 4;
    }
    const size_t NTmax = Mmax / K + NB + 1;           // sum_b ceil(n_b / K) <= M / K + (number of non-empty buckets)
    const uint32_t smax = 8 * (ROWCOL_THREADS / 4);   // more partials than this in one bucket: "giant" (side area + k_giant_finish)
    const unsigned w_lo = (c - 1) / 2;
    const uint32_t W = 1u << w_lo, nrows = (B >> w_lo) + 1;
    const size_t ntiles = (NB + PLAN_TILE - 1) / PLAN_TILE;

    static_assert(sizeof(xyzz_t) == 128 && sizeof(affine_t) == 64 && sizeof(fe) == 32, "layout");
    if (nrows > ROWCOL_MAX_ROWS || nrows > 3 * ROWCOL_THREADS) { zk_set_error("msm: window %u has too many bucket rows", c); return ZK_ERR_INVALID; }
    // scratch, grouped by what sizes it: the entry list (k * n * nwin), the bucket counters (G * B), the slice sums (G * c)
    {
        const size_t need_entries = Mmax * sizeof(uint32_t), need_partials = (NTmax + MSM_MAX_GIANTS) * sizeof(xyzz_t);
        if (ws.cap_entries < need_entries) {
            free_dev(ws.d_digits); free_dev(ws.d_entries);
            ws.d_digits = nullptr; ws.d_entries = nullptr; ws.cap_entries = 0;
            ZK_CUDA(cudaMalloc(&ws.d_digits, Mmax * sizeof(int32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_entries, need_entries));
            ws.cap_entries = need_entries;
        }
        if (ws.cap_partials < need_partials) {
            free_dev(ws.d_partials);
            ws.d_partials = nullptr; ws.cap_partials = 0;
            ZK_CUDA(cudaMalloc(&ws.d_partials, need_partials));
            ws.cap_partials = need_partials;
        }
        if (ws.cap_buckets < NB) {
            free_dev(ws.d_counts); free_dev(ws.d_offsets); free_dev(ws.d_task_off); free_dev(ws.d_chain); free_dev(ws.d_chain_flag);
            ws.d_counts = ws.d_offsets = ws.d_task_off = ws.d_chain_flag = nullptr; ws.d_chain = nullptr; ws.cap_buckets = 0;
            ZK_CUDA(cudaMalloc(&ws.d_counts, NB * sizeof(uint32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_offsets, (NB + 1) * sizeof(uint32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_task_off, (NB + 1) * sizeof(uint32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_chain, ntiles * sizeof(uint64_t)));
            ZK_CUDA(cudaMalloc(&ws.d_chain_flag, ntiles * sizeof(uint32_t)));
            ZK_CUDA(cudaMemsetAsync(ws.d_chain_flag, 0, ntiles * sizeof(uint32_t), st));
            ws.epoch = 0;
            ws.cap_buckets = NB;
        }
        const size_t need_bits = (size_t)G * (nrows + W) + (size_t)G * c;
        if (ws.cap_bits < need_bits) {
            free_dev(ws.d_bitsums);
            ws.d_bitsums = nullptr; ws.cap_bits = 0;
            ZK_CUDA(cudaMalloc(&ws.d_bitsums, need_bits * sizeof(xyzz_t)));
            ws.cap_bits = need_bits;
        }
        if (ws.cap_hbits < (size_t)G * c) {   // sized by G*c alone: few wide groups and many narrow ones differ
            if (ws.defer_sync) ZK_CUDA(cudaStreamSynchronize(st));   // an earlier deferred result may still be in flight
            if (ws.h_bitsums) cudaFreeHost(ws.h_bitsums);
            ws.h_bitsums = nullptr; ws.cap_hbits = 0;
            ZK_CUDA(cudaMallocHost(&ws.h_bitsums, 2 * (size_t)G * c * sizeof(xyzz_t)));
            ws.cap_hbits = (size_t)G * c;
        }
        if (!ws.d_meta) {
            ZK_CUDA(cudaMalloc(&ws.d_meta, 8 * sizeof(uint32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_giants, 3 * MSM_MAX_GIANTS * sizeof(uint32_t)));
            ZK_CUDA(cudaMalloc(&ws.d_giant_slices, (size_t)MSM_MAX_GIANTS * GIANT_SLICES * sizeof(xyzz_t)));
            ZK_CUDA(cudaMalloc(&ws.d_giant_tickets, MSM_MAX_GIANTS * sizeof(uint32_t)));
            ZK_CUDA(cudaMemsetAsync(ws.d_giant_tickets, 0, MSM_MAX_GIANTS * sizeof(uint32_t), st));
        }
    }
    unsigned nl = 0;
    if (ws.profile && !ws.ev[0])
        for (int s = 0; s <= MSM_ST_COUNT; s++) ZK_CUDA(cudaEventCreate(&ws.ev[s]));
#define STAGE_MARK(s) do { if (ws.profile) ZK_CUDA(cudaEventRecord(ws.ev[s], st)); } while (0)

    MsmScalarSet sc{};
    for (unsigned j = 0; j < k; j++) sc.p[j] = d_scalars[j];
    const uint32_t epoch = ++ws.epoch;
    ZK_CUDA(cudaMemsetAsync(ws.d_counts, 0, NB * sizeof(uint32_t), st));
    STAGE_MARK(0);
    // 1. digits + histogram
    k_recode<FS><<<dim3((unsigned)((n + 127) / 128), k), 128, 0, st>>>(sc, scalars_mont ? 1 : 0, n, c, nwin, gpm, use_table ? 0 : 1, ws.d_digits,
                                                                        ws.d_counts, ws.d_meta);
    STAGE_MARK(1);
    // 2. plan: bucket offsets, task offsets, giant list and side areas
    k_plan<<<(unsigned)ntiles, 1024, 0, st>>>(ws.d_counts, ws.d_offsets, ws.d_task_off, (uint32_t)NB, K, smax, ws.d_meta, ws.d_giants, ws.d_chain,
                                               ws.d_chain_flag, epoch);
    STAGE_MARK(2);
    // 3. scatter (counting sort by bucket)
    k_scatter<<<(unsigned)((Mmax + 255) / 256), 256, 0, st>>>(ws.d_digits, n, c, nwin, k, gpm, use_table ? 0 : 1, off, b.n, use_table ? 1 : 0,
                                                            n_main, n_extra, (uint32_t)main_count, ws.d_offsets, ws.d_counts, ws.d_entries);
    STAGE_MARK(3);
    // 4. accumulation: one task per <= K sorted entries of one bucket
    k_accumulate<F><<<(unsigned)((NTmax + 127) / 128), 128, 0, st>>>(b.d_points, ws.d_entries, ws.d_offsets, ws.d_task_off, (uint32_t)NB, K, ws.d_meta,
                                                                   ws.d_giants, d_extra, (uint32_t)main_count, ws.d_partials);
    STAGE_MARK(4);
    // 5. giants to one slot each, then row / column sums straight from the partial list
    k_giant_finish<F><<<dim3(GIANT_GRID, GIANT_SLICES), TREE_THREADS, 0, st>>>(ws.d_giants, ws.d_meta, ws.d_task_off, ws.d_partials, ws.d_giant_slices,
                                                                                 ws.d_giant_tickets);
    xyzz_t* d_rc = ws.d_bitsums;
    xyzz_t* d_T = ws.d_bitsums + (size_t)G * (nrows + W);
    k_rowcol<F><<<dim3(nrows + W, 1, G), ROWCOL_THREADS, 0, st>>>(ws.d_partials, ws.d_task_off, ws.d_meta, ws.d_giants, B, w_lo, d_rc);
    STAGE_MARK(5);
    // 6. bit slices of the row / column sums
    unsigned ft = TREE_THREADS;                 // half of the elements carry a given bit
    while (ft > 32 && ft / 4 >= std::max(W, nrows)) ft /= 2;
    k_bit_slices<F><<<dim3(c, G), ft, 0, st>>>(d_rc, B, w_lo, c, d_T);
    nl += 7;
    STAGE_MARK(6);
    ZK_CUDA(cudaGetLastError());
    shape->c = c; shape->groups = gpm;
    if (launches) *launches += nl;
    if (ws.d_T_out) {
        if ((size_t)G * c > ws.d_T_cap) { zk_set_error("msm: %u slice sums do not fit the caller's buffer of %zu points", G * c, ws.d_T_cap); return ZK_ERR_INVALID; }
        ZK_CUDA(cudaMemcpyAsync(ws.d_T_out, d_T, (size_t)G * c * sizeof(xyzz_t), cudaMemcpyDeviceToDevice, st));
        return ZK_OK;
    }
    ZK_CUDA(cudaMemcpyAsync(ws.h_bitsums + (size_t)(ws.h_slot & 1) * ws.cap_hbits, d_T, (size_t)G * c * sizeof(xyzz_t), cudaMemcpyDeviceToHost, st));
    if (ws.defer_sync) return ZK_OK;
    ZK_CUDA(cudaStreamSynchronize(st));
    if (ws.profile)
        for (int s = 0; s < MSM_ST_COUNT; s++) ZK_CUDA(cudaEventElapsedTime(&ws.stage_ms[s], ws.ev[s], ws.ev[s + 1]));
#undef STAGE_MARK
    // the O(c) serial tail (c doublings per group) is finished on the host from ws.h_bitsums (api.cu: msm_finish)
    return ZK_OK;
}

// one warp per output point: its 8 quads stride over the ranks, then a shuffle tree
template <class F> __global__ void __launch_bounds__(32) k_sum_partials(const xyzz_t* __restrict__ all, uint32_t world, uint32_t count, xyzz_t* out) {
    const unsigned i = blockIdx.x, qd = threadIdx.x >> 2;
    xyzz_t acc = xyzz_identity();
    for (uint32_t r0 = 0; r0 < world; r0 += 8) {
        const uint32_t r = r0 + qd;
        xyzz_t o = r < world ? load_xyzz(all + (size_t)r * count + i) : xyzz_identity();
        acc = xyzz_add_quad<F>(acc, o);
    }
#pragma unroll 1
    for (unsigned d = 4; d >= 1; d >>= 1) {
        xyzz_t o = shfl_down_xyzz(acc, 4 * d);
        if (qd >= d) o = xyzz_identity();
        acc = xyzz_add_quad<F>(acc, o);
    }
    if (threadIdx.x == 0) store_xyzz(out + i, acc);
}
template <class F> int msm_sum_partials(const xyzz_t* d_all, size_t world, size_t count, xyzz_t* d_out, cudaStream_t st) {
    if (count) k_sum_partials<F><<<(unsigned)count, 32, 0, st>>>(d_all, (uint32_t)world, (uint32_t)count, d_out);
    ZK_CUDA(cudaGetLastError());
    return ZK_OK;
}
template int msm_sum_partials<FpParams>(const xyzz_t*, size_t, size_t, xyzz_t*, cudaStream_t);
template int msm_sum_partials<FqParams>(const xyzz_t*, size_t, size_t, xyzz_t*, cudaStream_t);

#define INST(F, FS)                                                                                                             \
    template int msm_bases_create<F>(MsmBases&, const affine_t*, bool, size_t, unsigned, cudaStream_t);                          \
    template int msm_run<F, FS>(const MsmBases&, size_t, size_t, const fe* const*, unsigned, bool, unsigned, MsmWorkspace&, cudaStream_t, MsmResultShape*, unsigned*, \
                                const affine_t*, size_t);
INST(FpParams, FqParams)  // Pallas: coordinates Fp, scalars Fq
INST(FqParams, FpParams)  // Vesta:  coordinates Fq, scalars Fp
#undef INST

}  // namespace zkb
