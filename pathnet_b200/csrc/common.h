// Shared host-side helpers of libpathnet_b200.so: error TLS, CUDA checks,
// launch counter.
#pragma once
#include <cuda_runtime.h>

#include <atomic>
#include <cstdarg>
#include <cstdint>
#include <cstdio>

#include "../../include/pathnet_b200.h"

namespace pn {

void set_error(const char *fmt, ...);
extern std::atomic<uint64_t> g_launches;

inline void count_launch(uint64_t k = 1) { g_launches.fetch_add(k, std::memory_order_relaxed); }

#define PN_CUDA(expr)                                                                   \
    do {                                                                                \
        cudaError_t _e = (expr);                                                        \
        if (_e != cudaSuccess) {                                                        \
            pn::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e),      \
                          __FILE__, __LINE__);                                          \
            return PN_ERR_CUDA;                                                         \
        }                                                                               \
    } while (0)

#define PN_REQUIRE(cond, ...)                 \
    do {                                      \
        if (!(cond)) {                        \
            pn::set_error(__VA_ARGS__);       \
            return PN_ERR_INVALID;            \
        }                                     \
    } while (0)

// after a kernel launch
#define PN_LAUNCH_CHECK()                     \
    do {                                      \
        pn::count_launch();                   \
        PN_CUDA(cudaGetLastError());          \
    } while (0)

// NVTX ranges (header-only nvtx3: a no-op unless a profiler injects itself): one per C-ABI entry point
// (sampler S1/S2, PAGG K0-K4) and one per dominant kernel, so nsys / ncu timelines carry the reference-side names.
void nvtx_push(const char *name);
void nvtx_pop();
struct NvtxScope {
    explicit NvtxScope(const char *name) { nvtx_push(name); }
    ~NvtxScope() { nvtx_pop(); }
};

// in-stream timing of named kernels (pn_profile_enable / pn_profile_read)
void prof_begin(const char *slot, cudaStream_t st);
void prof_end(const char *slot, cudaStream_t st);
struct ProfScope {
    const char *slot; cudaStream_t st;
    ProfScope(const char *s, cudaStream_t t) : slot(s), st(t) { nvtx_push(slot); prof_begin(slot, st); }
    ~ProfScope() { prof_end(slot, st); nvtx_pop(); }
};

inline int64_t ceil_div(int64_t a, int64_t b) { return (a + b - 1) / b; }

}  // namespace pn
