// Counter-based dropout mask: 16 hash bits per element, keep iff
// hash16 >= floor(p * 65536).  Stateless, so forward and backward regenerate the
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

// one 32-bit hash serves the element PAIR (idx>>1): even idx uses the low 16
// bits, odd idx the high 16 bits.  key = dropout_key(seed, stream) is computed
// once per kernel; per pair the cost is one multiply + one murmur3 finalizer.
PN_HD uint32_t dropout_key(uint64_t seed, uint32_t stream_id)
{
    return fmix32((uint32_t)seed ^ 0x6a09e667u) ^ fmix32((uint32_t)(seed >> 32) + stream_id * 0x7f4a7c15u + 0x3c6ef372u);
}
PN_HD uint32_t dropout_hash_pair_k(uint32_t key, uint64_t pair)
{
    return fmix32(((uint32_t)pair * 0x9e3779b1u) ^ key ^ ((uint32_t)(pair >> 32) * 0x85ebca77u));
}
PN_HD uint32_t dropout_hash_pair(uint64_t seed, uint32_t stream_id, uint64_t pair)
{
    return dropout_hash_pair_k(dropout_key(seed, stream_id), pair);
}

// 16-bit value in [0, 65536) for element idx
PN_HD uint32_t dropout_hash(uint64_t seed, uint32_t stream_id, uint64_t idx)
{
    const uint32_t h = dropout_hash_pair(seed, stream_id, idx >> 1);
    return (idx & 1) ? (h >> 16) : (h & 0xFFFFu);
}

// threshold for drop probability p: drop iff hash16 < thr  (p resolved to 1/65536)
PN_HD uint32_t dropout_threshold(float p)
{
    double t = (double)p * 65536.0;
    if (t <= 0.0) return 0u;
    if (t >= 65536.0) return 65536u;
    return (uint32_t)t;
}

}  // namespace pn
