// ctx.hpp — the objects behind the opaque handles of include/zkb200.h.
#pragma once
#include <map>
#include <mutex>

#include "msm.cuh"
#include "ntt.cuh"

struct zk_ctx {
    int device = 0;
    cudaStream_t own_stream = nullptr, stream = nullptr;
    std::mutex mu;                       // a context serialises its calls (SRS: Sync + Send, SURVEY.md §8b "Threading")
    zkb::MsmWorkspace ws;
    zkb::fe* d_scalars = nullptr;        // staging for host-pointer MSM calls
    size_t cap_scalars = 0;
    zkb::fe* d_ntt = nullptr;            // staging for host-pointer NTT calls
    size_t cap_ntt = 0;
    zkb::fe* d_ntt_tmp = nullptr;        // second buffer of the two-pass plan
    size_t cap_ntt_tmp = 0;
    zkb::fe* ntt_small[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};  // [field][inverse] w_1024^(+-i)
    std::map<unsigned, zkb::NttTables> ntt_tables;                         // key: field | inverse << 1 | log_n << 2
    uint64_t launches = 0;
    bool profile = false;                // per-stage device timing (zk_ctx_set_profile)
    cudaEvent_t ev_ntt[2] = {nullptr, nullptr};
    float ntt_ms = 0;                    // device time of the last profiled NTT call (all its kernels)
};

struct zk_bases {
    zk_ctx* ctx = nullptr;
    zkb::MsmBases b;
};

namespace zkb {
int ctx_msm_device(zk_ctx* ctx, const zk_bases* bases, size_t off, size_t n, const fe* d_scalars, int mont, int window_bits,
                   uint64_t out_xyz[12]);
int ctx_ensure(void** p, size_t* cap, size_t bytes);
}  // namespace zkb
