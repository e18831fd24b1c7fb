// Counter-based dropout mask: one 32-bit hash per element, keep iff
// hash >= floor(p * 2^32).  Stateless, so forward and backward regenerate the
// same mask without storing it (F.dropout call sites: PathNet_run.py:194,209,
// 264,276).  The reference draws its masks from torch's Philox stream, which
// no other implementation can reproduce; parity tests therefore feed the
// oracle the masks computed from this very function (pn_dropout_hash).
#pragma once
#include <cstdint>

#if defined(__CUDACC__)
#define PN_HD __host__ __device__ __forceinline__
#else
#define PN_HD inline
#endif

namespace pn {

PN_HD uint32_t fmix32(uint32_t h)
{
    h ^= h >> 16; h *= 0x85ebca6bu;
    h ^= h >> 13; h *= 0xc2b2ae35u;
    h ^= h >> 16;
    return h;
}

PN_HD uint32_t dropout_hash(uint64_t seed, uint32_t stream_id, uint64_t idx)
{
    uint32_t h = fmix32((uint32_t)idx ^ (uint32_t)seed);
    h = fmix32(h ^ ((uint32_t)(idx >> 32) * 0x9e3779b9u) ^ (uint32_t)(seed >> 32) ^ (stream_id * 0x7f4a7c15u));
    return fmix32(h + 0x6a09e667u);
}

// threshold for drop probability p: drop iff hash < thr
PN_HD uint32_t dropout_threshold(float p)
{
    double t = (double)p * 4294967296.0;
    if (t <= 0.0) return 0u;
    if (t >= 4294967295.0) return 0xFFFFFFFFu;
    return (uint32_t)t;
}

}  // namespace pn
