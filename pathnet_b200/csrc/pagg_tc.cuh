// Declarations shared by the tensor-core LSTM kernels (pagg_tc.cu: fp16 single-pass mode; pagg_tc3.cu: split-precision mode).
#pragma once
#include "pagg_common.cuh"
#include "tc_util.cuh"

namespace pn {

// LSTM weights -> fp16, canonical layout, cut into 16 K-chunks of 16 so each chunk ([512 cols][16 k], 16 KB) is one
// contiguous bulk copy.  column = unit*4 + gate (i,f,g,o), k < H: W_ih, k >= H: W_hh      (H = 128)
constexpr int TC_H = 128, TC_G4 = 512, TC_K2 = 256, TC_KC = 16, TC_NCHUNK = TC_K2 / TC_KC;
constexpr int TC_CHUNK_BYTES = TC_G4 * TC_KC * 2;  // 16 KB

struct LstmTcArgs {
    const __half *T16;        // [N,L,H] fp16 table, pre-multiplied by the dropout scale (null: gather the fp32 table below)
    const float *T;           // [N,L,H] fp32 table
    const int32_t *seq_node;  // [P,L]
    const uint8_t *seq_type;  // [P,L]
    const unsigned char *Wh;  // packed fp16 weights (16 chunks x 16 KB)
    const float *bp;          // [512]
    __half *acts;             // tile images: block (tile, t) = [5 (i,f,g,o,c)][16 unit8][128 rows][8] fp16
    float *hfin;              // [P,H]
    int64_t P;
    int L;
    uint32_t drop_thr; float drop_scale; uint64_t drop_seed;
    const uint64_t *seed_dev; // optional device word added to drop_seed (CUDA-graph replay)
    int64_t p_base;           // global index of this call's first sequence (row-chunked calls)
    unsigned char *xh_img;    // wgrad operand images [x_t | h_{t-1}] (see LstmBwdTcArgs), written here
    uint32_t *dmask;          // [tile][t][cq][128 rows]: bit u of a word = "x_t[32 cq + u] was kept" — the backward's dx scatter
                              // reads 4 bytes instead of recomputing 16 hashes (null: no dropout, or nothing saved)
    int save;                 // 0: inference only, skip activations and operand images
    long long *trace;         // optional clock64() trace of CTA `trace_block` (debug)
    int trace_block;
    const __half *T16lo;      // split-precision mode (pagg_tc3.cu): fp16(T * scale - T16), the second term of the table
    unsigned char *xh_img_lo; // split-precision mode: second fp16 term of the [x_t | h_{t-1}] images
    float *acts32;            // split-precision mode: activations saved as fp32, block (tile, t) = [5][32 unit4][128 rows][4]
};

constexpr int ACT_BLOCK_HALFS = 5 * 16 * 128 * 8;  // 160 KB per (tile, step)
constexpr int DA_BLOCK_BYTES = 64 * 128 * 16, XH_BLOCK_BYTES = 32 * 128 * 16;  // wgrad operand images per (tile, step)
constexpr int TC_NSTAGE = 7;                        // 112 KB weight ring (largest that fits beside the operand tiles)
constexpr int TC_XBYTES = 128 * TC_H * 2;           // 32 KB: one [128][128] fp16 operand half (x_t or h_{t-1})
constexpr int TC_SMEM = 3 * TC_XBYTES + TC_NSTAGE * TC_CHUNK_BYTES + 2048 /*bias*/ + 128 * 8 * 4 /*row offsets*/ + 1024 /*align*/;
constexpr int TC_MAX_L = 8;

__device__ __forceinline__ float tanh_fast(float x)
{
    float y;
    asm("tanh.approx.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float sigmoid_fast(float x) { return fmaf(0.5f, tanh_fast(0.5f * x), 0.5f); }
// fp32-grade pieces of the split-precision mode: ex2.approx (2 ulp) and rcp.approx (1 ulp) instead of tanh.approx (2^-11)
__device__ __forceinline__ float ex2_approx(float x)
{
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
__device__ __forceinline__ float rcp_approx(float x)
{
    float y;
    asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// e^{-k x} for k = 1 (sigmoid) or 2 (tanh), capped at 2^40 so that products of three (1 + e) terms stay finite
__device__ __forceinline__ float exp_neg(float x) { return ex2_approx(fminf(-1.4426950408889634f * x, 40.f)); }
__device__ __forceinline__ float exp_neg2(float x) { return ex2_approx(fminf(-2.8853900817779268f * x, 40.f)); }
__device__ __forceinline__ float tanh_acc(float x)
{
    const float e = exp_neg2(x);
    return (1.f - e) * rcp_approx(1.f + e);
}

constexpr int TC_CWARPS = 16;                    // compute warps: 4 TMEM lane quarters x 4 column quarters
constexpr int TC_THREADS = (TC_CWARPS + 2) * 32;  // forward: + MMA warp + weight-producer warp
// backward: + a whole service warpgroup (MMA warp, W' producer, two idle warps) so that setmaxnreg -- which works on warpgroups
// of 4 consecutive warps, inside the register pool the CTA was given at launch -- can move registers: 20 warps = 5 per SM
// sub-partition -> 96 per thread at launch; the service warpgroup shrinks to 32 and the 16 compute warps grow to 112.
constexpr int TCB_THREADS = (TC_CWARPS + 4) * 32;
__device__ __forceinline__ bool elect_one()   // true in exactly one lane of a converged warp
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
template <int N> __device__ __forceinline__ void reg_inc() { asm volatile("setmaxnreg.inc.sync.aligned.u32 %0;" ::"n"(N)); }
template <int N> __device__ __forceinline__ void reg_dec() { asm volatile("setmaxnreg.dec.sync.aligned.u32 %0;" ::"n"(N)); }

// W' = [W_ih | W_hh] as B operand of the dgrad GEMM: [256 rows = input index k][K = 512 gate columns],
// fp16 canonical, cut into 16 chunks of 32 columns (16 KB each).
constexpr int TCB_KC = 32, TCB_NCHUNK = TC_G4 / TCB_KC, TCB_CHUNK_BYTES = TC_K2 * TCB_KC * 2;

struct LstmBwdTcArgs {
    const float *T;
    const int32_t *seq_node;
    const uint8_t *seq_type;
    const unsigned char *WhT;  // 16 chunks x 16 KB
    const __half *acts;        // tile images (see LstmTcArgs)
    const float *dhfin;        // [P,H], or null: dh_final[p] = dh_k[p] * gdrop[p / nw] + dh_s[p] * attw_h (see PoolArgs)
    const float *dh_k, *dh_s, *gdrop, *attw_h;
    int64_t gdrop_ld;
    int nw;
    const float *gscale;       // {scale, 1/scale}
    float *dT;                 // [N,L,H]
    unsigned char *da_img;     // block b = t*ntiles + tile: [64 col8][128 r][8] fp16   (128 KB)
    const unsigned char *xh_img;  // block b: [32 k8][128 r][8] fp16 (64 KB): k8 < 16 x_t, k8 >= 16 h_{t-1}; written by the FORWARD kernel
    const uint32_t *dmask;        // the forward's dropout-mask words (see LstmTcArgs), or null: recompute the hashes
    int64_t P, ntiles;
    int L;
    uint32_t drop_thr; float drop_scale; uint64_t drop_seed;
    const uint64_t *seed_dev;
    int64_t p_base;
    long long *trace;
    int trace_block;
    const float *acts32;       // split-precision mode: activations as fp32 tile images (see k_lstm_fwd_tc3)
    unsigned char *da_img_lo;  // split-precision mode: second fp16 term of the da images (same block layout as da_img)
    int l2pf;                  // k_lstm_bwd_tc<PF = true>: prefetch.global.L2 of the next step's activations (off by default; PN_BWD_L2PF=1 turns it on)
};
constexpr int TCB_NSTAGE = 4;   // 64 KB of W' in flight beside the 128 KB operand tile: the deepest ring that stays inside the 196 KB
                                // shared-memory carve-out (6 stages -> 228 KB carve-out, the smaller L1 costs more than the depth wins)
constexpr int TCB_A_BYTES = 128 * TC_G4 * 2;  // 128 KB: da tile [128][512] fp16 canonical
constexpr int TCB_SMEM_FIXED = TCB_A_BYTES + TCB_NSTAGE * TCB_CHUNK_BYTES + 128 /*align*/;   // + 128 * L * 4 bytes of row offsets

__device__ __forceinline__ void prefetch_l2(const void *p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }

}  // namespace pn
