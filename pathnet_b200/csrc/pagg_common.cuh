// Shared pieces of the PAGG kernels: workspace layout, generic fp32 GEMM,
// small elementwise kernels.  See pagg.cu for the operator itself.
#pragma once
#include <algorithm>

#include "common.h"
#include "dropout_hash.h"

namespace pn {

// ---------------------------------------------------------------------------
// workspace carving (all regions 256-byte aligned)
// ---------------------------------------------------------------------------
struct PaggWs {
    // forward products kept for backward
    float *Z;        // [N,H]   fc0 output (homo: after ReLU)
    float *T;        // [N,L,H] per-(node,type) projection table (homo: after ReLU)
    float *Wp;       // [2H][4H] packed LSTM weights, column = unit*4+gate
    float *WpT;      // [4H][2H]
    float *bp;       // [4H]     b_ih+b_hh, packed
    float *acts;     // [L][P][5][H]  i,f,g,o,c of every step
    float *hfin;     // [P,H]
    float *score;    // [P] raw attention score
    float *att;      // [P] pooling weight before the 1/nw (softmax prob or 1+score)
    float *cat;      // [S,2H] fc2 input (after dropout)
    // backward scratch
    float *dcat;     // [S,2H]
    float *dhfin;    // [P,H]
    float *dZ;       // [N,H]
    float *dT;       // [N,L,H]
    float *dA;       // [L*P,4H] pre-activation gate grads, packed columns
    float *XH;       // [L*P,2H] LSTM step inputs [x_t | h_{t-1}]
    float *dWp;      // [2H][4H]
    float *dbp;      // [4H]
    float *dattw;    // [2H+1]
    float *gscale;   // [8] gradient scales {s, 1/s} of the tensor-core backward: [0,1] LSTM (max |dh_final|), [2,3] dT, [4,5] dZ
    // tensor-core mode only (fp16 storage)
    float *x16;      // [N][F16] fp16 copy of X (unless the caller supplied cfg.x16)
    float *z16;      // [N][H]   fp16 copy of Z written by the fc0 epilogue
    float *w16;      // fc0_w [H][F16] then nets_w [L*H][H], fp16
    float *t16;      // [N][L*H] fp16 copy of T times the dropout scale (what the LSTM kernel gathers); split-precision mode: two
                     // such arrays, fp16(T s) then fp16(T s - hi)
    float *dmask;    // [tile][L][4][128] uint32 dropout-mask words of the LSTM input (forward -> backward)
    size_t bytes;
};

inline size_t ws_align(size_t x) { return (x + 255) & ~(size_t)255; }

// rows per pass over the per-sequence kernels (pn_pagg_cfg.row_chunk; 0 = the whole call)
inline int64_t chunk_rows(const pn_pagg_cfg &c)
{
    return (c.row_chunk > 0 && c.row_chunk < c.S) ? (int64_t)c.row_chunk : (int64_t)std::max(c.S, 1);
}

// one pass: output rows [s0, s0 + Sc) of the call
struct PaggChunk {
    int64_t s0, Sc;
    bool save;    // keep activations / gradient operands for this chunk's backward
    bool first;   // first chunk of the call: weight packing happens here
};

inline PaggWs carve_workspace(const pn_pagg_cfg &c, void *base)
{
    PaggWs w{};
    char *p = (char *)base;
    size_t off = 0;
    // per-node regions are whole-call; per-sequence regions are sized for one row chunk
    const size_t N = c.N, H = c.H, L = c.L, S = (size_t)chunk_rows(c), P = S * c.nw;
    const size_t P128 = (P + 127) / 128 * 128;  // the tensor-core kernels work on whole 128-sequence tiles
    auto take = [&](float *&dst, size_t nfloats) {
        dst = (float *)(p + off);
        off += ws_align(nfloats * sizeof(float));
    };
    take(w.Z, N * H);
    take(w.T, N * L * H);
    take(w.Wp, 2 * H * 4 * H);
    take(w.WpT, 4 * H * 2 * H);
    take(w.bp, 4 * H);
    take(w.acts, L * P128 * 5 * H);
    take(w.hfin, P * H);
    take(w.score, P);
    take(w.att, P);
    take(w.cat, S * 2 * H);
    take(w.dcat, S * 2 * H);
    take(w.dhfin, P * H);
    take(w.dZ, N * H);
    take(w.dT, N * L * H);
    take(w.dA, L * P128 * 4 * H);
    take(w.XH, L * P128 * 2 * H);
    take(w.dWp, 2 * H * 4 * H);
    take(w.dbp, 4 * H);
    take(w.dattw, 2 * H + 1);
    take(w.gscale, 8);
    // tensor-core mode: 16-bit copies feeding the TMA-loaded projection GEMMs (see pagg_tc.cu)
    const size_t F16 = ((size_t)c.F + 7) / 8 * 8;
    take(w.x16, c.math_mode == PN_MATH_TC && !c.x16 ? (N * F16 + 1) / 2 : 0);
    take(w.z16, c.math_mode == PN_MATH_TC ? (N * H + 1) / 2 : 0);
    take(w.w16, c.math_mode == PN_MATH_TC ? (H * F16 + L * H * H + 1) / 2 : 0);
    take(w.t16, c.math_mode == PN_MATH_TC ? (N * L * H + 1) / 2 : c.math_mode == PN_MATH_TC3 ? N * L * H : 0);   // TC3: hi then lo term
    take(w.dmask, c.math_mode == PN_MATH_TC || c.math_mode == PN_MATH_TC3 ? L * P128 * 4 : 0);
    w.bytes = off;
    return w;
}

// ---------------------------------------------------------------------------
// generic strided fp32 GEMM on CUDA cores:
//   C[m,n] (+)= epi( sum_k A(m,k) * B(k,n) ),  A(m,k) = A[m*sAm + k*sAk],
//   B(k,n) = B[k*sBk + n*sBn], C row-major with leading dimension ldc.
// 64x64 tile, 16-deep k-chunks, 256 threads x (4x4).  gridDim.z > 1 splits K
// and accumulates with atomics (C must be pre-zeroed / hold the addend).
// ---------------------------------------------------------------------------
struct GemmArgs {
    const float *A, *B;
    float *C;
    const float *bias;  // per n, may be null
    int64_t M, N, K;
    int64_t sAm, sAk, sBk, sBn, ldc;
    int relu;        // epilogue ReLU
    int accumulate;  // C += result
};

static __global__ void __launch_bounds__(256) k_gemm_f32(GemmArgs g)
{
    __shared__ float As[16][64 + 4];
    __shared__ float Bs[16][64 + 4];
    const int tid = threadIdx.x;
    const int64_t m0 = (int64_t)blockIdx.x * 64, n0 = (int64_t)blockIdx.y * 64;   // row tiles on grid.x (no 65 535 limit)
    const int tm = (tid / 16) * 4, tn = (tid % 16) * 4;
    // split-K range
    const int64_t kchunks = (g.K + 15) / 16;
    const int64_t per = (kchunks + gridDim.z - 1) / gridDim.z;
    const int64_t kbeg = (int64_t)blockIdx.z * per * 16;
    const int64_t kend = min(g.K, kbeg + per * 16);
    float acc[4][4] = {};
    // loader mapping: make the unit-stride dimension the fastest across threads
    const bool a_m_fast = (g.sAm == 1), b_n_fast = (g.sBn == 1);
    for (int64_t k0 = kbeg; k0 < kend; k0 += 16) {
#pragma unroll
        for (int r = 0; r < 4; r++) {
            int e = tid + r * 256;  // 0..1023
            int am = a_m_fast ? (e & 63) : (e >> 4), ak = a_m_fast ? (e >> 6) : (e & 15);
            int64_t gm = m0 + am, gk = k0 + ak;
            As[ak][am] = (gm < g.M && gk < kend) ? g.A[gm * g.sAm + gk * g.sAk] : 0.f;
            int bn = b_n_fast ? (e & 63) : (e >> 4), bk = b_n_fast ? (e >> 6) : (e & 15);
            int64_t gn = n0 + bn, gk2 = k0 + bk;
            Bs[bk][bn] = (gn < g.N && gk2 < kend) ? g.B[gk2 * g.sBk + gn * g.sBn] : 0.f;
        }
        __syncthreads();
#pragma unroll
        for (int kk = 0; kk < 16; kk++) {
            float4 a = *reinterpret_cast<const float4 *>(&As[kk][tm]);
            float4 b = *reinterpret_cast<const float4 *>(&Bs[kk][tn]);
            float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
            for (int i = 0; i < 4; i++)
#pragma unroll
                for (int j = 0; j < 4; j++) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        __syncthreads();
    }
    const bool split = gridDim.z > 1;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int64_t m = m0 + tm + i;
        if (m >= g.M) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int64_t n = n0 + tn + j;
            if (n >= g.N) continue;
            float v = acc[i][j];
            float *c = g.C + m * g.ldc + n;
            if (split) {
                if (g.bias && blockIdx.z == 0) v += g.bias[n];
                atomicAdd(c, v);
            } else {
                if (g.bias) v += g.bias[n];
                if (g.relu) v = fmaxf(v, 0.f);
                *c = g.accumulate ? *c + v : v;
            }
        }
    }
}

inline int launch_gemm(cudaStream_t st, const float *A, int64_t sAm, int64_t sAk, const float *B,
                       int64_t sBk, int64_t sBn, float *C, int64_t ldc, int64_t M, int64_t N,
                       int64_t K, const float *bias, bool relu, bool accumulate, int splitk = 1)
{
    if (M <= 0 || N <= 0) return PN_OK;
    GemmArgs g{A, B, C, bias, M, N, K, sAm, sAk, sBk, sBn, ldc, relu ? 1 : 0, accumulate ? 1 : 0};
    if (splitk > 1) {
        // atomics accumulate: caller guarantees C holds zeros (or the addend); relu unsupported
        if (relu) { set_error("launch_gemm: split-K with ReLU"); return PN_ERR_INVALID; }
    }
    dim3 grid((unsigned)ceil_div(M, 64), (unsigned)ceil_div(N, 64), (unsigned)splitk);
    k_gemm_f32<<<grid, 256, 0, st>>>(g);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

// max |v| of a block's values -> atomicMax on the float's bit pattern (non-negative floats order like unsigned ints)
__device__ __forceinline__ void block_absmax(float m, unsigned int *absmax_bits)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m > 0.f) atomicMax(absmax_bits, __float_as_uint(m));
}

// column sums: out[n] = sum_m A[m*ld + n]   (out pre-zeroed; atomics across row blocks);
// optionally max |A| -> *absmax_bits (pre-zeroed)
static __global__ void k_colsum(const float *__restrict__ A, int64_t M, int64_t N, int64_t ld,
                         float *__restrict__ out, unsigned int *absmax_bits)
{
    int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f, mx = 0.f;
    if (n < N) {
        int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
        int64_t m0 = (int64_t)blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
        for (int64_t m = m0; m < m1; m++) { const float v = A[m * ld + n]; s += v; mx = fmaxf(mx, fabsf(v)); }
        atomicAdd(out + n, s);
    }
    if (absmax_bits) block_absmax(mx, absmax_bits);
}

inline int launch_colsum(cudaStream_t st, const float *A, int64_t M, int64_t N, int64_t ld, float *out, bool accumulate = false,
                         unsigned int *absmax_bits = nullptr)
{
    if (N <= 0) return PN_OK;
    if (!accumulate) PN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, st));
    if (M <= 0) return PN_OK;
    int64_t yb = std::max<int64_t>(1, std::min<int64_t>(1024, M / 256));
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)yb);
    k_colsum<<<grid, 128, 0, st>>>(A, M, N, ld, out, absmax_bits);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

// ReLU backward and the bias gradient in one pass over a dense [M][N] gradient: g = (y > 0) ? g : 0, out[n] = sum_m g[m][n].
// (The two used to be separate sweeps of the same 40 MB table gradient.)
static __global__ void k_relu_mask_colsum(float *__restrict__ g, const float *__restrict__ y, int64_t M, int64_t N,
                                          float *__restrict__ out, unsigned int *absmax_bits)
{
    const int64_t n = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    float s = 0.f, mx = 0.f;
    if (n < N) {
        const int64_t rows_per = (M + gridDim.y - 1) / gridDim.y;
        const int64_t m0 = (int64_t)blockIdx.y * rows_per, m1 = min(M, m0 + rows_per);
        // batches of 8 rows: all 16 loads are issued before the first (possibly aliasing, as far as the compiler knows) store
        for (int64_t mb = m0; mb < m1; mb += 8) {
            float v[8], a[8];
#pragma unroll
            for (int i = 0; i < 8; i++) {
                const bool in = mb + i < m1;
                v[i] = in ? g[(mb + i) * N + n] : 0.f;
                a[i] = in ? y[(mb + i) * N + n] : 1.f;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                if (a[i] > 0.f) { s += v[i]; mx = fmaxf(mx, fabsf(v[i])); }
                else g[(mb + i) * N + n] = 0.f;   // never taken for the padding entries (a = 1)
            }
        }
        atomicAdd(out + n, s);
    }
    if (absmax_bits) block_absmax(mx, absmax_bits);   // max |g| after masking: the tensor-core GEMMs scale their fp16 operand by it
}

inline int launch_relu_mask_colsum(cudaStream_t st, float *g, const float *y, int64_t M, int64_t N, float *out,
                                   unsigned int *absmax_bits = nullptr)
{
    if (N <= 0) return PN_OK;
    PN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * (size_t)N, st));
    if (M <= 0) return PN_OK;
    const int64_t yb = std::max<int64_t>(1, std::min<int64_t>(2048, M / 64));
    dim3 grid((unsigned)ceil_div(N, 128), (unsigned)yb);
    k_relu_mask_colsum<<<grid, 128, 0, st>>>(g, y, M, N, out, absmax_bits);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

}  // namespace pn
