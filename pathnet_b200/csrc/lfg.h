// glibc srand()/rand() (TYPE_3 additive feedback, x^31 = x^28 + 1 over Z/2^32)
// with polynomial jump-ahead.  The reference sampler draws every random number
// from one sequential rand() stream (gen_merw.cpp:88-89, :161); to reproduce it
// bit-exactly on thousands of threads each thread must start at an arbitrary
// offset of that stream.
//
// Sequence convention: s_0..s_30 = (r_3, .., r_33) of the seeded state (r_31..33
// are copies of r_0..2), s_{j+31} = s_j + s_{j+28}; the k-th rand() is
// s_{341+k} >> 1.  For any k,  s_{k+j} = sum_i c_i s_{i+j}  with
// sum c_i x^i = x^k mod (x^31 - x^28 - 1).
#pragma once
#include <cstdint>

namespace pn {
namespace lfg {

constexpr int DEG = 31;
constexpr int64_t FIRST_OUTPUT = 341;  // index in s of the first rand() output

struct Poly {
    uint32_t c[DEG];
};

// srand(seed) -> window s_0..s_30 (glibc __srandom_r, TYPE_3)
inline void seed_window(uint32_t seed, uint32_t s[DEG])
{
    uint32_t r[34];
    if (seed == 0) seed = 1;
    int32_t word = (int32_t)seed;
    r[0] = (uint32_t)word;
    for (int i = 1; i < 31; i++) {
        long hi = word / 127773, lo = word % 127773;
        word = (int32_t)(16807 * lo - 2836 * hi);
        if (word < 0) word += 2147483647;
        r[i] = (uint32_t)word;
    }
    r[31] = r[0]; r[32] = r[1]; r[33] = r[2];
    for (int j = 0; j < DEG; j++) s[j] = r[3 + j];
}

inline Poly poly_one() { Poly p{}; p.c[0] = 1; return p; }
inline Poly poly_x()   { Poly p{}; p.c[1] = 1; return p; }

// a*b mod (x^31 - x^28 - 1), coefficients mod 2^32
inline Poly poly_mul(const Poly &a, const Poly &b)
{
    uint32_t t[2 * DEG - 1] = {0};
    for (int i = 0; i < DEG; i++) {
        uint32_t ai = a.c[i];
        if (!ai) continue;
        for (int j = 0; j < DEG; j++) t[i + j] += ai * b.c[j];
    }
    for (int d = 2 * DEG - 2; d >= DEG; d--) {  // x^d = x^(d-3) + x^(d-31)
        uint32_t v = t[d];
        t[d - 3] += v;
        t[d - DEG] += v;
    }
    Poly r;
    for (int i = 0; i < DEG; i++) r.c[i] = t[i];
    return r;
}

// x^k
inline Poly poly_xpow(uint64_t k)
{
    Poly result = poly_one(), base = poly_x();
    while (k) {
        if (k & 1) result = poly_mul(result, base);
        k >>= 1;
        if (k) base = poly_mul(base, base);
    }
    return result;
}

// window at offset +k: out_j = sum_i c_i s_{i+j}
inline void window_apply(const Poly &p, const uint32_t s[DEG], uint32_t out[DEG])
{
    uint32_t e[2 * DEG - 1];
    for (int j = 0; j < DEG; j++) e[j] = s[j];
    for (int d = 0; d < DEG - 1; d++) e[DEG + d] = e[d] + e[d + 28];
    for (int j = 0; j < DEG; j++) {
        uint32_t acc = 0;
        for (int i = 0; i < DEG; i++) acc += p.c[i] * e[i + j];
        out[j] = acc;
    }
}

}  // namespace lfg
}  // namespace pn
