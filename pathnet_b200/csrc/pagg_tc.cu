// Tensor-core (tcgen05 / TMEM) math modes of PAGG: fp16 operands, fp32 accumulation in TMEM, fp32 gate math.
// See tc_util.cuh for the operand layout, pagg_tc.cuh for the shared declarations and pagg_tc3.cu for the
// split-precision ("tc3") LSTM kernels.
//
//   k_gemm_tc      C (+)= act(op(A) op(B)^T + bias), fp32 operands in global memory converted to fp16 while staged into
//                  the canonical UMMA layout (K- or MN-major, split-K, scaled gradients): the backward projection
//                  GEMMs and the stand-alone tcgen05 unit test.  <..., BK = 32, SPLIT>: both operands as two fp16
//                  terms, three MMAs per K step (fp32-grade products; the projections of the tc3 mode).
//   k_lstm_fwd_tc  fused gather + dropout + LSTM forward.  One CTA = 128 sequences (one TMEM lane each), all L steps:
//                    16 compute warps (4 TMEM lane quarters x 4 column quarters): gather x_{t+1} (fp16 table rows,
//                               dropout as a bit mask) into the other x buffer while the tensor core works, read the
//                               128 x 512 fp32 gate pre-activations from TMEM, gate math (c in registers), write h_t
//                               back as the next step's operand, save activations and the wgrad operand image;
//                    MMA warp   converged warp, one elected lane issues tcgen05.mma (M128 N256 K16, 32 per step);
//                    producer   one thread streams the packed fp16 weights (256 KB per step, L2-resident) through a
//                               7-stage shared-memory ring with cp.async.bulk + mbarriers.
//   k_lstm_bwd_tc  BPTT: gate gradients -> da (dgrad operand in smem + wgrad image), [dx | dh] = da W' pipelined in
//                  K-quarters, dx scattered into dT with vector REDs, dc / dh carried in TMEM.
//   k_wgrad_tc     dW = sum over blocks [x|h]^T da: 4-CTA clusters, multicast operand quarters.
//   pagg_tc_project / _backward: the projections around the LSTM (TMA-fed k_gemm_tma forward, k_gemm_tc backward).
#include "pagg_common.cuh"
#include <cuda_bf16.h>

#include <cstdlib>

#include "tc_util.cuh"
#include "pagg_tc.cuh"

namespace pn {
using namespace tc;

// ---------------------------------------------------------------------------
// generic tcgen05 GEMM:  C[M,N] (+)= act( op(A) op(B)^T + bias )
//   A_MN = false: A is [M][K] row-major (K contiguous)  -> K-major operand
//   A_MN = true : A is [K][M] row-major (M contiguous)  -> MN-major operand
//   likewise B ([N][K] or [K][N]).  fp32 sources are converted to fp16 (or bf16,
//   for raw gradients whose magnitude would underflow fp16) while being staged.
//   gridDim.z > 1 splits K; split or `accumulate` results are added atomically.
// Both layouts use 16-byte chunks of 8 elements, LBO = 128, SBO = 1024 for a
// 64-deep K tile, and advance 256 bytes per K = 16 step.
// ---------------------------------------------------------------------------
struct GemmTcArgs {
    const float *A, *B;
    float *C;            // [M,N] ldc
    const float *bias;   // [N] or null
    int64_t M, N, K, lda, ldb, ldc;
    int relu, accumulate;
    const float *a_scale;   // optional device {s, 1/s}: A is staged as fp16(A * s) and the result multiplied by 1/s (raw
                            // gradients would underflow fp16; s is a power of two taken from max |A|, see k_absmax_scale)
};

template <bool BF16>
__device__ __forceinline__ uint4 pack8_half_or_bf16(const float (&f)[8])
{
    if (!BF16) return pack8_f16(f);
    uint4 pk;
    __nv_bfloat162 b0 = __floats2bfloat162_rn(f[0], f[1]), b1 = __floats2bfloat162_rn(f[2], f[3]);
    __nv_bfloat162 b2 = __floats2bfloat162_rn(f[4], f[5]), b3 = __floats2bfloat162_rn(f[6], f[7]);
    pk.x = *reinterpret_cast<uint32_t *>(&b0); pk.y = *reinterpret_cast<uint32_t *>(&b1);
    pk.z = *reinterpret_cast<uint32_t *>(&b2); pk.w = *reinterpret_cast<uint32_t *>(&b3);
    return pk;
}

// stage a [ROWS x 64] operand tile.  MN = false: src[row][k]; MN = true: src[k][row].
// Two phases so that all of a thread's global loads are in flight together.
// BK: depth of the K tile (64, or 32 for the split-precision variant); SPLIT: also write the second fp16 term of every value
// (lo = fp16(v - hi)) as a second tile image at dst_lo.
template <int ROWS, int NTHREADS, bool MN, bool BF16, int BK = 64, bool SPLIT = false>
__device__ __forceinline__ void stage_tile(unsigned char *dst, const float *__restrict__ src, int64_t ld, int64_t row0,
                                           int64_t nrows, int64_t k0, int64_t kend, int tid, float mul = 1.f, unsigned char *dst_lo = nullptr)
{
    constexpr int CHUNKS = ROWS * BK / 8;  // 16-byte chunks (8 elements)
    constexpr int PER = CHUNKS / NTHREADS;
    static_assert(CHUNKS % NTHREADS == 0, "tile must divide evenly among the threads");
    const bool vec_ok = (ld % 4 == 0) && ((reinterpret_cast<uintptr_t>(src) & 15) == 0);
    float4 lo[PER], hi[PER];
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int c = tid + i * NTHREADS;
        int64_t base;   // element offset of the chunk's first value; its 8 values are contiguous in src
        bool full, any;
        if (!MN) {
            const int r = c % ROWS, k8 = c / ROWS;
            const int64_t gr = row0 + r, gk = k0 + k8 * 8;
            base = gr * ld + gk;
            any = gr < nrows && gk < kend;
            full = gr < nrows && gk + 8 <= kend;
        } else {
            const int k = c % BK, r8 = c / BK;
            const int64_t gk = k0 + k, gr = row0 + r8 * 8;
            base = gk * ld + gr;
            any = gk < kend && gr < nrows;
            full = gk < kend && gr + 8 <= nrows;
        }
        if (full && vec_ok) {
            lo[i] = __ldg(reinterpret_cast<const float4 *>(src + base));
            hi[i] = __ldg(reinterpret_cast<const float4 *>(src + base + 4));
        } else {
            float f[8];
            const int64_t lim = !MN ? (kend - (k0 + (c / ROWS) * 8)) : (nrows - (row0 + (c / BK) * 8));
#pragma unroll
            for (int j = 0; j < 8; j++) f[j] = (any && j < lim) ? __ldg(src + base + j) : 0.f;
            lo[i] = make_float4(f[0], f[1], f[2], f[3]);
            hi[i] = make_float4(f[4], f[5], f[6], f[7]);
        }
    }
#pragma unroll
    for (int i = 0; i < PER; i++) {
        const int c = tid + i * NTHREADS;
        const uint32_t off = !MN ? canon_off(c % ROWS, (c / ROWS) * 8, BK)
                                 : (uint32_t)((c / BK) * (BK / 8 * 128) + ((c % BK) >> 3) * 128 + ((c % BK) & 7) * 16);
        const float f[8] = {lo[i].x * mul, lo[i].y * mul, lo[i].z * mul, lo[i].w * mul, hi[i].x * mul, hi[i].y * mul, hi[i].z * mul, hi[i].w * mul};
        const uint4 pk = pack8_half_or_bf16<BF16>(f);
        *reinterpret_cast<uint4 *>(dst + off) = pk;
        if constexpr (SPLIT) {
            float back[8], rest[8];
            unpack8_f16(pk, back);
#pragma unroll
            for (int j = 0; j < 8; j++) rest[j] = f[j] - back[j];
            *reinterpret_cast<uint4 *>(dst_lo + off) = pack8_f16(rest);
        }
    }
}

// 256 threads: all eight warps stage operand tiles (the kernel is bound by the latency of those loads), warps 0-3 own the
// four TMEM lane quarters in the epilogue.
// SPLIT (split-precision mode "tc3"): both operands are staged as two fp16 terms and every K step is three MMAs,
// A B ~= A_hi B_hi + A_lo B_hi + A_hi B_lo, into the same fp32 accumulator: fp32-grade products on the tensor core.  Its K tile is
// 32 deep so that a stage (four tile images) stays at 32 KB and three CTAs still fit on an SM.
template <int BN, bool A_MN, bool B_MN, bool BF16, int BK = 64, bool SPLIT = false>
__global__ void __launch_bounds__(256) k_gemm_tc(GemmTcArgs g)
{
    constexpr int BM = 128;
    static_assert(!(SPLIT && BF16), "split precision uses fp16 terms");
    extern __shared__ __align__(128) unsigned char smem_raw[];
    constexpr int TERM = (BM + BN) * BK * 2;      // A tile then B tile, 16-bit: one term
    constexpr int STAGE = TERM * (SPLIT ? 2 : 1); // hi term, then lo term; two stages (double buffer)
    __shared__ uint64_t bar[2];
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * BM, n0 = (int64_t)blockIdx.y * BN;   // row tiles on grid.x (no 65 535 limit)
    // split-K range (multiples of BK)
    const int64_t kt = (g.K + BK - 1) / BK, per = (kt + gridDim.z - 1) / gridDim.z;
    const int64_t kbeg = (int64_t)blockIdx.z * per * BK, kend = min(g.K, kbeg + per * BK);
    if (tid == 0) { mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc<BN>(&s_taddr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_taddr;
    constexpr uint32_t idesc = make_idesc_f16(BM, BN) | (BF16 ? ((1u << 7) | (1u << 10)) : 0u) |
                               (A_MN ? (1u << 15) : 0u) | (B_MN ? (1u << 16) : 0u);
    // staging of tile i+1 overlaps the MMAs of tile i; a buffer is reused once the MMAs that
    // read it (two tiles earlier) have committed
    uint32_t phase[2] = {0, 0};
    int it = 0;
    const float a_mul = g.a_scale ? g.a_scale[0] : 1.f, out_mul = g.a_scale ? g.a_scale[1] : 1.f;
    for (int64_t k0 = kbeg; k0 < kend; k0 += BK, it++) {
        const int b = it & 1;
        unsigned char *As = smem_raw + b * STAGE, *Bs = As + BM * BK * 2;
        if (it >= 2) { mbar_wait(&bar[b], phase[b]); phase[b] ^= 1; }
        stage_tile<BM, 256, A_MN, BF16, BK, SPLIT>(As, g.A, g.lda, m0, g.M, k0, kend, tid, a_mul, As + TERM);
        stage_tile<BN, 256, B_MN, BF16, BK, SPLIT>(Bs, g.B, g.ldb, n0, g.N, k0, kend, tid, 1.f, Bs + TERM);
        fence_proxy_async();
        __syncthreads();
        if (tid == 0) {
            tc_fence_after();
            const uint64_t ad = make_smem_desc(smem_u32(As), 128, (BK / 8) * 128);
            const uint64_t bd = make_smem_desc(smem_u32(Bs), 128, (BK / 8) * 128);
#pragma unroll
            for (int kk = 0; kk < BK / 16; kk++) {
                mma_f16(taddr, ad + (uint64_t)(kk * 16), bd + (uint64_t)(kk * 16), idesc, (k0 > kbeg) || (kk > 0));
                if constexpr (SPLIT) {
                    const uint64_t ad_lo = make_smem_desc(smem_u32(As + TERM), 128, (BK / 8) * 128);
                    const uint64_t bd_lo = make_smem_desc(smem_u32(Bs + TERM), 128, (BK / 8) * 128);
                    mma_f16(taddr, ad_lo + (uint64_t)(kk * 16), bd + (uint64_t)(kk * 16), idesc, true);
                    mma_f16(taddr, ad + (uint64_t)(kk * 16), bd_lo + (uint64_t)(kk * 16), idesc, true);
                }
            }
            mma_commit(&bar[b]);
        }
    }
    // drain: the last (up to two) commits
    if (it >= 2) { const int b = it & 1; mbar_wait(&bar[b], phase[b]); }
    if (it >= 1) { const int b = (it - 1) & 1; mbar_wait(&bar[b], phase[b]); }
    tc_fence_after();
    // epilogue: warp w owns TMEM lanes 32w..32w+31 = rows m0+32w..+31.  tcgen05.ld hands each lane one ROW; written out like
    // that every store instruction would touch 32 different lines, so each 32 x 32 block is transposed through shared memory
    // (the operand stages are idle: every MMA has committed) and leaves as whole 128-byte row segments.
    const bool atomic = gridDim.z > 1 || g.accumulate;
    if (kbeg < kend && warp < 4) {
        float *stg = reinterpret_cast<float *>(smem_raw) + warp * (32 * 33);
        for (int c0 = 0; c0 < BN; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) stg[lane * 33 + j] = __uint_as_float(v[j]);
            __syncwarp();
            const int64_t n = n0 + c0 + lane;
            if (n < g.N) {
                const float bn = (g.bias && blockIdx.z == 0) ? g.bias[n] : 0.f;
                const int64_t rmax = min((int64_t)32, g.M - (m0 + warp * 32));
                float *dst = g.C + (m0 + warp * 32) * g.ldc + n;
                for (int rr = 0; rr < rmax; rr++) {
                    const float x = stg[rr * 33 + lane] * out_mul + bn;
                    if (atomic) atomicAdd(dst + (int64_t)rr * g.ldc, x);
                    else dst[(int64_t)rr * g.ldc] = g.relu ? fmaxf(x, 0.f) : x;
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc<BN>(taddr); }
}

// (Measured and dropped: a kind::tf32 variant that cp.async-copies the fp32 operands straight into the UMMA layout -- no
// conversion pass, 4 x 32 KB in flight per CTA.  Its 192 KB ring allows one CTA per SM; this kernel's 64 KB allows three,
// which matters more for these short-K GEMMs: 75 us vs 44 us on the fc0 / table projections.  MN-major tf32 operands
// also failed to validate with the layout that is correct for 16-bit types.)
// a_mn / b_mn: operand given as [K][M] / [K][N]; bf16: 16-bit format of the operands;
// splitk > 1 or accumulate: results are atomically added to C (no ReLU then).
int launch_gemm_tc(cudaStream_t st, const float *A, int64_t lda, const float *B, int64_t ldb, float *C, int64_t ldc,
                   int64_t M, int64_t N, int64_t K, const float *bias, bool relu, bool a_mn, bool b_mn, bool bf16,
                   bool accumulate, int splitk, const float *a_scale, bool split3 = false)
{
    if (M <= 0 || N <= 0) return PN_OK;
    if ((splitk > 1 || accumulate) && relu) { set_error("launch_gemm_tc: ReLU with atomic accumulation"); return PN_ERR_INVALID; }
    GemmTcArgs g{A, B, C, bias, M, N, K, lda, ldb, ldc, relu ? 1 : 0, accumulate ? 1 : 0, a_scale};
    constexpr int BN = 128;
    const size_t smem = 2 * (128 + BN) * 64 * 2;
    dim3 grid((unsigned)ceil_div(M, 128), (unsigned)ceil_div(N, BN), (unsigned)std::max(1, splitk));
#define PN_GEMM_CASE(AM, BM_, BF)                                                                              \
    do {                                                                                                       \
        PN_CUDA(cudaFuncSetAttribute(k_gemm_tc<BN, AM, BM_, BF>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)); \
        k_gemm_tc<BN, AM, BM_, BF><<<grid, 256, smem, st>>>(g);                                                \
    } while (0)
    if (split3) {   // split-precision variant: fp16 hi + lo terms of both operands, K tile 32
        if (bf16) { set_error("launch_gemm_tc: split precision uses fp16 terms"); return PN_ERR_INVALID; }
        const size_t smem3 = 2 * 2 * (128 + BN) * 32 * 2;
#define PN_GEMM_CASE3(AM, BM_)                                                                                 \
    do {                                                                                                       \
        PN_CUDA(cudaFuncSetAttribute(k_gemm_tc<BN, AM, BM_, false, 32, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem3)); \
        k_gemm_tc<BN, AM, BM_, false, 32, true><<<grid, 256, smem3, st>>>(g);                                  \
    } while (0)
        if (!a_mn && !b_mn) PN_GEMM_CASE3(false, false);
        else if (a_mn && b_mn) PN_GEMM_CASE3(true, true);
        else if (!a_mn && b_mn) PN_GEMM_CASE3(false, true);
        else { set_error("launch_gemm_tc: operand layout combination not instantiated (split precision)"); return PN_ERR_UNSUPPORTED; }
#undef PN_GEMM_CASE3
        PN_LAUNCH_CHECK();
        return PN_OK;
    }
    if (!a_mn && !b_mn && !bf16) PN_GEMM_CASE(false, false, false);
    else if (!a_mn && b_mn && bf16) PN_GEMM_CASE(false, true, true);
    else if (a_mn && b_mn && bf16) PN_GEMM_CASE(true, true, true);
    else if (a_mn && b_mn && !bf16) PN_GEMM_CASE(true, true, false);
    else if (!a_mn && b_mn && !bf16) PN_GEMM_CASE(false, true, false);
    else if (!a_mn && !b_mn && bf16) PN_GEMM_CASE(false, false, true);
    else { set_error("launch_gemm_tc: operand layout combination not instantiated"); return PN_ERR_UNSUPPORTED; }
#undef PN_GEMM_CASE
    PN_LAUNCH_CHECK();
    return PN_OK;
}

int launch_f32_to_f16_rows(cudaStream_t st, const float *src, int64_t rows, int64_t K, int64_t ld_src, __half *dst, int64_t K16);   // gemm_tma.cu
int launch_pack_lstm_tc3_bwd(cudaStream_t st, const float *w_ih, const float *w_hh, unsigned char *WhT);   // pagg_tc3.cu
int launch_lstm_bwd_tc3(const LstmBwdTcArgs &a, int64_t ntiles, int smem, int smem_max, cudaStream_t st);    // pagg_tc3.cu
int launch_gemm_tma(cudaStream_t st, const __half *A16, const __half *B16, int64_t M, int64_t N, int64_t K16, const float *bias,
                    bool relu, float *C, int64_t ldc, __half *C16, int64_t ldc16, float c16_scale);                                  // gemm_tma.cu
static bool tc_legacy_gemm()
{
    static const bool legacy = [] { const char *e = getenv("PN_TC_GEMM"); return e && e[0] == 'l'; }();
    return legacy;
}

// ---------------------------------------------------------------------------
// LSTM weights -> fp16, canonical layout, cut into 16 K-chunks of 16 so each
// chunk ([512 cols][16 k], 16 KB) is one contiguous bulk copy.
//   column = unit*4 + gate (i,f,g,o), k < H: W_ih, k >= H: W_hh      (H = 128)
// ---------------------------------------------------------------------------

__global__ void k_pack_lstm_tc(const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                               const float *__restrict__ b_ih, const float *__restrict__ b_hh,
                               unsigned char *__restrict__ Wh, float *__restrict__ bp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < TC_G4 * TC_K2) {
        const int col = i / TC_K2, k = i % TC_K2, unit = col >> 2, gate = col & 3;
        const float v = k < TC_H ? w_ih[(gate * TC_H + unit) * TC_H + k] : w_hh[(gate * TC_H + unit) * TC_H + (k - TC_H)];
        const int chunk = k / TC_KC, kk = k % TC_KC;
        // sigmoid(a) = 0.5 tanh(a/2) + 0.5: the 1/2 of the i, f, o gates is folded into weights and bias
        // (exact in fp16), so the epilogue issues one tanh.approx per gate and no pre-scale
        const float sc = gate == 2 ? 1.f : 0.5f;
        *reinterpret_cast<__half *>(Wh + (size_t)chunk * TC_CHUNK_BYTES + canon_off(col, kk, TC_KC)) = __float2half_rn(sc * v);
    }
    if (i < TC_G4) {
        const int unit = i >> 2, gate = i & 3;
        bp[i] = (gate == 2 ? 1.f : 0.5f) * (b_ih[gate * TC_H + unit] + b_hh[gate * TC_H + unit]);
    }
}


// this thread's quarter (32 values) of table row `row_off` -> dropout -> fp16 -> x tile (canonical, K extent 128).
// With the fp16 table (written by the projection GEMM's epilogue, already times 1/(1-p)) the row is 4 x 16 bytes and
// dropout is a bit mask on the packed halves; the fp32 table costs 8 loads, the scale multiply and the conversion.
__device__ __forceinline__ void tc_gather_quarter(const LstmTcArgs &a, unsigned char *Xs, int r, int cq, int32_t row_off,
                                                   uint64_t elem0 /* (p*L + t) * H */, unsigned char *xh_blk, uint32_t *mask_out)
{
    const bool valid = row_off >= 0;
    const uint32_t key = dropout_key(a.seed_dev ? a.drop_seed + *a.seed_dev : a.drop_seed, 0);
    if (a.T16) {
        const uint4 *src = reinterpret_cast<const uint4 *>(a.T16 + (int64_t)(valid ? row_off : 0) * TC_H + cq * 32);
        uint4 ld[4];
#pragma unroll
        for (int i = 0; i < 4; i++) ld[i] = valid ? __ldg(src + i) : make_uint4(0, 0, 0, 0);
        uint32_t mbits = 0;
#pragma unroll
        for (int kk = 0; kk < 4; kk++) {
            const int k8 = cq * 4 + kk;
            uint32_t wv[4] = {ld[kk].x, ld[kk].y, ld[kk].z, ld[kk].w};
            if (valid && a.drop_thr) {
                const uint64_t pair = (elem0 + k8 * 8) >> 1;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t h = dropout_hash_pair_k(key, pair + q);
                    const uint32_t k0 = (h & 0xFFFFu) < a.drop_thr ? 0u : 1u, k1 = (h >> 16) < a.drop_thr ? 0u : 1u;
                    wv[q] &= (0u - k0) & 0x0000FFFFu | (0u - k1) & 0xFFFF0000u;
                    mbits |= (k0 | (k1 << 1)) << (kk * 8 + q * 2);
                }
            }
            const uint4 pk = make_uint4(wv[0], wv[1], wv[2], wv[3]);
            *reinterpret_cast<uint4 *>(Xs + canon_off(r, k8 * 8, TC_H)) = pk;
            if (a.save) *reinterpret_cast<uint4 *>(xh_blk + (size_t)k8 * 2048 + r * 16) = pk;   // x half of the wgrad operand image
        }
        if (mask_out) *mask_out = mbits;
        return;
    }
    const float *src = a.T + (int64_t)(valid ? row_off : 0) * TC_H;
    float4 ld[8];   // all eight 16-byte loads of the quarter row in flight together
#pragma unroll
    for (int i = 0; i < 8; i++) ld[i] = valid ? __ldg(reinterpret_cast<const float4 *>(src + cq * 32 + i * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int kk = 0; kk < 4; kk++) {
        const int k8 = cq * 4 + kk;
        float f[8] = {ld[2 * kk].x, ld[2 * kk].y, ld[2 * kk].z, ld[2 * kk].w, ld[2 * kk + 1].x, ld[2 * kk + 1].y, ld[2 * kk + 1].z, ld[2 * kk + 1].w};
        if (valid) {
            if (a.drop_thr) {
                const uint64_t pair = (elem0 + k8 * 8) >> 1;
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint32_t h = dropout_hash_pair_k(key, pair + q);
                    f[2 * q] = (h & 0xFFFFu) < a.drop_thr ? 0.f : f[2 * q] * a.drop_scale;
                    f[2 * q + 1] = (h >> 16) < a.drop_thr ? 0.f : f[2 * q + 1] * a.drop_scale;
                }
            }
        }
        const uint4 pk = pack8_f16(f);
        *reinterpret_cast<uint4 *>(Xs + canon_off(r, k8 * 8, TC_H)) = pk;
        if (a.save) *reinterpret_cast<uint4 *>(xh_blk + (size_t)k8 * 2048 + r * 16) = pk;   // x half of the wgrad operand image
    }
}

__global__ void __launch_bounds__(TC_THREADS, 1) k_lstm_fwd_tc(LstmTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *Xs = base;                       // two x_t buffers (double buffer), [128][128] fp16 canonical each
    unsigned char *Hs = base + 2 * TC_XBYTES;       // h_{t-1}
    unsigned char *Ws = base + 3 * TC_XBYTES;       // ring of weight chunks
    float *bias_s = reinterpret_cast<float *>(Ws + TC_NSTAGE * TC_CHUNK_BYTES);
    int32_t *rowoff_s = reinterpret_cast<int32_t *>(bias_s + TC_G4);   // [128][TC_MAX_L]: table row of (sequence, step), -1 = padding
    __shared__ uint64_t full_bar[TC_NSTAGE], empty_bar[TC_NSTAGE], a_ready, acc_full;
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t p0 = (int64_t)blockIdx.x * 128;
    const int L = a.L;

    if (tid == 0) {
        for (int s = 0; s < TC_NSTAGE; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&a_ready, TC_CWARPS * 32);
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == TC_CWARPS) tmem_alloc<512>(&s_taddr);
    for (int i = tid; i < TC_G4; i += TC_THREADS) bias_s[i] = a.bp[i];
    for (int i = tid; i < 128 * L; i += TC_THREADS) {
        const int r = i / L, t = i % L;
        const int64_t p = p0 + r;
        int32_t off = -1;
        if (p < a.P) { const int64_t e = p * L + t; off = a.seq_node[e] * L + a.seq_type[e]; }
        rowoff_s[r * TC_MAX_L + t] = off;
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_taddr;

    if (warp == TC_CWARPS + 1) {
        // ===== weight producer: chunk sequence per step = (t == 0 ? x half : x half + h half) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = 0; t < L; t++) {
                const int nch = t == 0 ? TC_NCHUNK / 2 : TC_NCHUNK;
                for (int c = 0; c < nch; c++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);  // slot free (first pass: passes immediately)
                    mbar_arrive_expect_tx(&full_bar[stage], TC_CHUNK_BYTES);
                    bulk_g2s(Ws + stage * TC_CHUNK_BYTES, a.Wh + (size_t)c * TC_CHUNK_BYTES, TC_CHUNK_BYTES, &full_bar[stage]);
                    if (++stage == TC_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == TC_CWARPS) {
        // ===== MMA issuer: the whole warp runs the loop (uniform control flow keeps the descriptors in uniform registers: no
        //       R2UR / election loop per MMA, tools/probe/ring_probe.cu), one elected lane issues =====
        {
            constexpr uint32_t idesc = make_idesc_f16(128, 256);
            uint32_t stage = 0, phase = 0;
            const bool trm = a.trace && blockIdx.x == a.trace_block && lane == 0;
            for (int t = 0; t < L; t++) {
                mbar_wait(&a_ready, t & 1);  // x_t (and h_{t-1}) staged, TMEM drained
                tc_fence_after();
                if (trm) a.trace[t * 8 + 4] = clock64();
                const int nch = t == 0 ? TC_NCHUNK / 2 : TC_NCHUNK;  // h_0 = 0: skip the W_hh half
                const uint32_t xbase = smem_u32(Xs + (t & 1) * TC_XBYTES), hbase = smem_u32(Hs);
                for (int c = 0; c < nch; c++) {
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t wbase = smem_u32(Ws + stage * TC_CHUNK_BYTES);
                    if (elect_one()) {
                        // chunk c = 16 consecutive k: two core matrices (256 B) of the x half (c < 8) or the h half
                        const uint32_t abase = (c < TC_NCHUNK / 2 ? xbase + (uint32_t)c * 256 : hbase + (uint32_t)(c - TC_NCHUNK / 2) * 256);
                        const uint64_t ad = make_smem_desc(abase, 128, (TC_H / 8) * 128);
#pragma unroll
                        for (int nh = 0; nh < 2; nh++) {
                            const uint64_t bd = make_smem_desc(wbase + (uint32_t)nh * (256 / 8) * ((TC_KC / 8) * 128), 128, (TC_KC / 8) * 128);
                            mma_f16(taddr + (uint32_t)nh * 256, ad, bd, idesc, c > 0);
                        }
                        mma_commit(&empty_bar[stage]);  // slot reusable once these MMAs have read it
                        if (c == nch - 1) mma_commit(&acc_full);
                    }
                    __syncwarp();
                    if (++stage == TC_NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (trm) a.trace[t * 8 + 5] = clock64();
            }
        }
    } else {
        // ===== gather + gate math: warp = (TMEM lane quarter lq, column quarter cq); the thread owns
        //       sequence p0 + 32 lq + lane (= its TMEM lane) and hidden units [32 cq, 32 cq + 32) =====
        const int lq = warp & 3, cq = warp >> 2;
        const int r = lq * 32 + lane;
        const int64_t p = p0 + r;
        const bool valid = p < a.P;
        float c[32];
#pragma unroll
        for (int j = 0; j < 32; j++) c[j] = 0.f;
        const bool tr = a.trace && blockIdx.x == a.trace_block && tid == 0;
        if (tr) a.trace[6] = clock64();
        const int64_t ntiles = gridDim.x;
        // dropout-mask word of (tile, t, cq, row): written only when the backward will need it
        auto mask_slot = [&](int t_) { return (a.dmask && a.save && a.drop_thr) ? a.dmask + (((int64_t)blockIdx.x * L + t_) * 4 + cq) * 128 + r : nullptr; };
        tc_gather_quarter(a, Xs, r, cq, rowoff_s[r * TC_MAX_L + 0], (uint64_t)((p + a.p_base) * L + 0) * TC_H,
                          a.xh_img + ((int64_t)0 * ntiles + blockIdx.x) * XH_BLOCK_BYTES, mask_slot(0));
        // h_{-1} = 0: h half of step 0's image
#pragma unroll
        for (int k8 = 16 + cq * 4; k8 < 16 + cq * 4 + 4; k8++)
            if (a.save) *reinterpret_cast<uint4 *>(a.xh_img + (int64_t)blockIdx.x * XH_BLOCK_BYTES + (size_t)k8 * 2048 + r * 16) = make_uint4(0, 0, 0, 0);
        for (int t = 0; t < L; t++) {
            if (tr) a.trace[t * 8 + 0] = clock64();
            fence_proxy_async();   // generic-proxy smem writes (x_t, h_{t-1}) -> async proxy (tcgen05.mma)
            tc_fence_before();     // order this thread's earlier tcgen05.ld before the MMA that overwrites TMEM
            mbar_arrive(&a_ready);
            // x_{t+1} into the other x buffer while the tensor core works on step t (MMA(t-1), the last
            // reader of that buffer, completed before acc_full(t-1) was observed)
            if (t + 1 < L)
                tc_gather_quarter(a, Xs + ((t + 1) & 1) * TC_XBYTES, r, cq, rowoff_s[r * TC_MAX_L + t + 1],
                                  (uint64_t)((p + a.p_base) * L + t + 1) * TC_H, a.xh_img + ((int64_t)(t + 1) * ntiles + blockIdx.x) * XH_BLOCK_BYTES,
                                  mask_slot(t + 1));
            if (tr) a.trace[t * 8 + 1] = clock64();
            // ---- wait for the gate pre-activations ----
            mbar_wait(&acc_full, t & 1);
            tc_fence_after();
            if (tr) a.trace[t * 8 + 2] = clock64();
            // activations of (tile, t): chunk (q, unit8, row) -> consecutive rows are consecutive 16-byte chunks
            __half *act = a.acts + ((int64_t)blockIdx.x * L + t) * ACT_BLOCK_HALFS + r * 8;
            const bool t_is_last = t == L - 1;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int c0 = cq * 128 + jj * 32;   // packed gate columns of units [c0/4, c0/4 + 8)
                // outputs are packed to fp16 pairs as they are produced (keeps the register footprint small)
                uint32_t pi[4], pf[4], pg[4], po[4], pc[4], ph[4];
                float hq[4];
                uint32_t v[16];
#pragma unroll
                for (int u2 = 0; u2 < 4; u2++) {
                    if ((u2 & 1) == 0) {   // 16 columns = 4 units per TMEM load
                        tmem_ld16(taddr + ((uint32_t)(lq * 32) << 16) + (uint32_t)(c0 + u2 * 8), v);
                        tmem_ld_wait();
                    }
                    float gi2[2], gf2[2], gg2[2], go2[2], cn2[2], hh2[2];
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int u = 2 * u2 + e, unit = jj * 8 + u;   // index into this thread's 32 cell states
                        const int vi = 4 * (u & 3);
                        const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c0 + 4 * u);
                        gi2[e] = fmaf(0.5f, tanh_fast(__uint_as_float(v[vi + 0]) + b4.x), 0.5f);   // pre-activation already halved
                        gf2[e] = fmaf(0.5f, tanh_fast(__uint_as_float(v[vi + 1]) + b4.y), 0.5f);
                        gg2[e] = tanh_fast(__uint_as_float(v[vi + 2]) + b4.z);
                        go2[e] = fmaf(0.5f, tanh_fast(__uint_as_float(v[vi + 3]) + b4.w), 0.5f);
                        const float cn = fmaf(gf2[e], c[unit], gi2[e] * gg2[e]);
                        c[unit] = cn;
                        cn2[e] = cn;
                        hh2[e] = go2[e] * tanh_fast(cn);
                    }
                    __half2 t2;
                    t2 = __floats2half2_rn(gi2[0], gi2[1]); pi[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(gf2[0], gf2[1]); pf[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(gg2[0], gg2[1]); pg[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(go2[0], go2[1]); po[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(cn2[0], cn2[1]); pc[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(hh2[0], hh2[1]); ph[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    hq[(u2 & 1) * 2] = hh2[0]; hq[(u2 & 1) * 2 + 1] = hh2[1];
                    if (t_is_last && (u2 & 1)) {
                        // final h (fp32) -> staging tile in the (now idle) x buffers, 16-byte slots XOR-swizzled by
                        // the row so that 8 consecutive rows hit 8 different bank groups
                        const int j4 = c0 / 16 + (u2 >> 1);   // float4 slot of units [4 j4, 4 j4 + 4)
                        *reinterpret_cast<float4 *>(Xs + r * 512 + ((j4 ^ (r & 7)) << 4)) = make_float4(hq[0], hq[1], hq[2], hq[3]);
                    }
                }
                // h_t -> operand tile for the next step, and h half of the next step's wgrad operand image
                if (!t_is_last) {
                    const uint4 hpk = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4 *>(Hs + canon_off(r, c0 / 4, TC_H)) = hpk;
                    if (a.save) *reinterpret_cast<uint4 *>(a.xh_img + ((int64_t)(t + 1) * ntiles + blockIdx.x) * XH_BLOCK_BYTES + (size_t)(16 + c0 / 32) * 2048 + r * 16) = hpk;
                }
                if (valid && a.save) {
                    const int ub = (c0 / 32) * 1024;  // unit8 block of 128 rows x 8 halves
                    *reinterpret_cast<uint4 *>(act + 0 * 16384 + ub) = make_uint4(pi[0], pi[1], pi[2], pi[3]);
                    *reinterpret_cast<uint4 *>(act + 1 * 16384 + ub) = make_uint4(pf[0], pf[1], pf[2], pf[3]);
                    *reinterpret_cast<uint4 *>(act + 2 * 16384 + ub) = make_uint4(pg[0], pg[1], pg[2], pg[3]);
                    *reinterpret_cast<uint4 *>(act + 3 * 16384 + ub) = make_uint4(po[0], po[1], po[2], po[3]);
                    *reinterpret_cast<uint4 *>(act + 4 * 16384 + ub) = make_uint4(pc[0], pc[1], pc[2], pc[3]);
                }
            }
            if (tr) a.trace[t * 8 + 3] = clock64();
        }
        // ---- final h: staging tile -> hfin[P][H], whole 512-byte rows per warp ----
        asm volatile("bar.sync 1, %0;" ::"n"(TC_CWARPS * 32) : "memory");
        for (int rr = warp * 8; rr < warp * 8 + 8; rr++) {
            const int64_t pr = p0 + rr;
            if (pr < a.P) {
                const float4 hv = *reinterpret_cast<const float4 *>(Xs + rr * 512 + ((lane ^ (rr & 7)) << 4));
                *reinterpret_cast<float4 *>(a.hfin + pr * TC_H + lane * 4) = hv;
            }
        }
        if (tr) a.trace[7] = clock64();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CWARPS) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

// Forward LSTM in tensor-core mode (called from pn_pagg_forward).
int pagg_tc_forward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm,
                         const int32_t *seq_node, const uint8_t *seq_type, const PaggChunk &ck, cudaStream_t st)
{
    if (c.H != TC_H || c.L > TC_MAX_L) {
        set_error("PN_MATH_TC: the tensor-core LSTM is built for hidden size 128 and path_len <= %d (got H=%d L=%d)", TC_MAX_L, c.H, c.L);
        return PN_ERR_UNSUPPORTED;
    }
    // packed fp16 weights live in the (otherwise unused in this mode) Wp/WpT region: 2 x 256 KB fp32 >= 256 KB fp16
    unsigned char *Wh = reinterpret_cast<unsigned char *>(w.Wp);
    if (ck.first) {
        k_pack_lstm_tc<<<(unsigned)ceil_div(TC_G4 * TC_K2, 256), 256, 0, st>>>(prm.w_ih, prm.w_hh, prm.b_ih, prm.b_hh, Wh, w.bp);
        PN_LAUNCH_CHECK();
    }
    uint32_t thr = 0; float scale = 1.f;
    if (c.training && c.dropout_p > 0.f) {
        thr = dropout_threshold(c.dropout_p);
        scale = c.dropout_p >= 1.f ? 0.f : 1.f / (1.f - c.dropout_p);
    }
    const int64_t P = ck.Sc * c.nw;
    LstmTcArgs a{tc_legacy_gemm() ? nullptr : reinterpret_cast<const __half *>(w.t16), w.T, seq_node, seq_type, Wh, w.bp,
                 reinterpret_cast<__half *>(w.acts), w.hfin, P, c.L, thr, scale, c.dropout_seed,
                 c.dropout_seed_dev, ck.s0 * c.nw, reinterpret_cast<unsigned char *>(w.XH),
                 tc_legacy_gemm() ? nullptr : reinterpret_cast<uint32_t *>(w.dmask), ck.save ? 1 : 0, nullptr, 0};
    long long *d_trace = nullptr;
    const char *trace_env = getenv("PN_TRACE_FWD");
    if (trace_env) {
        PN_CUDA(cudaMalloc(&d_trace, sizeof(long long) * 8 * 64));
        PN_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(long long) * 8 * 64, st));
        a.trace = d_trace;
        a.trace_block = atoi(trace_env);
    }
    PN_CUDA(cudaFuncSetAttribute(k_lstm_fwd_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, TC_SMEM));
    {
        ProfScope prof("lstm_fwd", st);
        k_lstm_fwd_tc<<<(unsigned)ceil_div(P, 128), TC_THREADS, TC_SMEM, st>>>(a);
    }
    PN_LAUNCH_CHECK();
    if (d_trace) {  // debug only: synchronous dump of the phase timestamps of one CTA
        long long h[8 * 64];
        PN_CUDA(cudaStreamSynchronize(st));
        PN_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        const long long t00 = h[0];
        for (int t = 0; t < c.L && t < 64; t++)
            fprintf(stderr, "[trace fwd cta %d] t=%d  begin %lld | arrive+prefetch done %lld | mma-thread saw a_ready %lld | mma issued %lld | acc_full seen %lld | epilogue done %lld\n",
                    a.trace_block, t, h[t * 8 + 0] - t00, h[t * 8 + 1] - t00, h[t * 8 + 4] - t00, h[t * 8 + 5] - t00, h[t * 8 + 2] - t00, h[t * 8 + 3] - t00);
        fprintf(stderr, "[trace fwd cta %d] first gather started %lld, hfin written %lld\n", a.trace_block, h[6] - t00, h[7] - t00);
    }
    return PN_OK;
}


// ===========================================================================
// tensor-core backward
// ===========================================================================
// Gradients travel as fp16 multiplied by a power-of-two scale chosen from
// max|dh_final| (taken by k_pool_bwd; the largest gate gradient then sits near 2^4): the same 11-bit
// significand as the forward operands, with 2^12 of head-room against growth
// through the recurrence and 2^-28 relative before flush-to-zero.
__global__ void k_make_scale(float *gscale)
{
    const float m = __uint_as_float(*reinterpret_cast<unsigned int *>(gscale));
    float sc = 1.f;
    if (m > 0.f && isfinite(m)) sc = exp2f(4.f - ceilf(log2f(m)));
    gscale[0] = sc;
    gscale[1] = 1.f / sc;
}

__global__ void k_pack_lstm_tc_bwd(const float *__restrict__ w_ih, const float *__restrict__ w_hh, unsigned char *__restrict__ WhT)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_G4 * TC_K2) return;
    const int col = i / TC_K2, k = i % TC_K2, unit = col >> 2, gate = col & 3;
    const float v = k < TC_H ? w_ih[(gate * TC_H + unit) * TC_H + k] : w_hh[(gate * TC_H + unit) * TC_H + (k - TC_H)];
    const int chunk = col / TCB_KC, kk = col % TCB_KC;
    *reinterpret_cast<__half *>(WhT + (size_t)chunk * TCB_CHUNK_BYTES + canon_off(k, kk, TCB_KC)) = __float2half_rn(v);
}


// (Measured and dropped: pairing neighbouring tiles in a cluster that multicasts the W' stream halves its L2 traffic but
// couples the two CTAs chunk by chunk -- 5 % slower.
//  Round 2, also measured and dropped: two 32 KB quarter buffers for da instead of the whole tile + an 8-stage W' ring
//  (same 192 KB): 1.144 vs 1.072 ms.  In-kernel traces put the wait for acc_full after the last quarter at ~3 k cycles per
//  step with EITHER ring depth -- it is the 16 MMAs of the last quarter plus commit / wake-up latency, not the stream --
//  while the extra "buffer free" waits slowed the gate-gradient loop from 10.5 k to 12.9 k cycles per step.)
// PF = true (default): activations are fetched one half-iteration (4 units, six 8-byte loads) ahead of the gate-gradient
// math; PF = false (PN_BWD_PF=0): the round-1 loop (8 units per iteration, loads consumed where issued).  Same-box A/B,
// pubmed step (profiles/r02/bench_tc_bwd_pf{0,1}_run5.json): 1.072 vs 1.111 ms per launch.
template <bool PF>
__global__ void __launch_bounds__(TCB_THREADS, 1) k_lstm_bwd_tc(LstmBwdTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    // no-swizzle UMMA operands and bulk copies need 16-byte alignment only; 128 keeps the tiles on line boundaries
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    unsigned char *As = base;                 // da_t, [128][512] fp16 canonical K-major
    unsigned char *Ws = base + TCB_A_BYTES;   // ring of W' chunks
    int32_t *rowoff_s = reinterpret_cast<int32_t *>(Ws + TCB_NSTAGE * TCB_CHUNK_BYTES);   // [128][L]: dT row of (sequence, step)
    __shared__ uint64_t full_bar[TCB_NSTAGE], empty_bar[TCB_NSTAGE], a_ready[4], acc_full;
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t tile = blockIdx.x, p0 = tile * 128;
    const int L = a.L;
    for (int i = tid; i < 128 * L; i += TCB_THREADS) {
        const int rr = i / L, tt = i % L;
        const int64_t pp = p0 + rr;
        int32_t off = -1;
        if (pp < a.P) { const int64_t e = pp * L + tt; off = a.seq_node[e] * L + a.seq_type[e]; }
        rowoff_s[rr * L + tt] = off;
    }
    if (tid == 0) {
        for (int s = 0; s < TCB_NSTAGE; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int i = 0; i < 4; i++) mbar_init(&a_ready[i], TC_CWARPS * 32);
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == TC_CWARPS) tmem_alloc<512>(&s_taddr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    // TMEM columns: [0,128): dx of the step, [128,256) / [256,384): dh_{t-1} written by even / odd steps (ping-pong: the
    // tensor core already accumulates this step's dh while the gate-gradient loop still reads the previous one),
    // [384,512): running dc.
    // Pipelining inside a step: the gate-gradient loop produces da in four quarters (iteration i of every warp fills
    // K-chunks {i, 4+i, 8+i, 12+i} of the operand tile); the MMA warp starts on a quarter as soon as it is complete, so
    // the weight streaming and the tensor-core time hide under the remaining gate-gradient work.
    const uint32_t taddr = s_taddr;

    if (warp >= TC_CWARPS) {
    reg_dec<32>();   // all four warps of the service warpgroup (two of them have nothing else to do)
    if (warp == TC_CWARPS + 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int s = 0; s < L; s++)
                for (int j = 0; j < TCB_NCHUNK; j++) {
                    const int c = 4 * (j & 3) + (j >> 2);   // quarter-major order, the order the MMA warp consumes them in
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], TCB_CHUNK_BYTES);
                    bulk_g2s(Ws + stage * TCB_CHUNK_BYTES, a.WhT + (size_t)c * TCB_CHUNK_BYTES, TCB_CHUNK_BYTES, &full_bar[stage]);
                    if (++stage == TCB_NSTAGE) { stage = 0; phase ^= 1; }
                }
        }
    } else if (warp == TC_CWARPS) {
        // the whole warp runs the loop (uniform control flow: descriptors stay in uniform registers, no R2UR / election loop
        // per MMA -- tools/probe/ring_probe.cu: ~15 % less issue overhead per stage), one elected lane issues
        uint32_t stage = 0, phase = 0;
        constexpr uint32_t idesc = make_idesc_f16(128, 128);
        for (int s = 0; s < L; s++) {
            const int t = L - 1 - s;
            const uint32_t dh_col = 128u + (uint32_t)(s & 1) * 128u;
            for (int j = 0; j < TCB_NCHUNK; j++) {
                const int c = 4 * (j & 3) + (j >> 2);
                if ((j & 3) == 0) { mbar_wait(&a_ready[j >> 2], s & 1); tc_fence_after(); }
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t wbase = smem_u32(Ws + stage * TCB_CHUNK_BYTES);
                if (elect_one()) {
#pragma unroll
                    for (int ks = 0; ks < TCB_KC / 16; ks++) {
                        const uint64_t ad = make_smem_desc(smem_u32(As) + (uint32_t)(c * (TCB_KC / 8) + ks * 2) * 128, 128, (TC_G4 / 8) * 128);
                        // rows 0..127 of W' give dx, rows 128..255 (16 row groups further) give dh_{t-1} (unused at t = 0)
                        const uint64_t bd_x = make_smem_desc(wbase + (uint32_t)ks * 256, 128, (TCB_KC / 8) * 128);
                        mma_f16(taddr, ad, bd_x, idesc, (j > 0) || (ks > 0));
                        if (t > 0) {
                            const uint64_t bd_h = make_smem_desc(wbase + 16u * ((TCB_KC / 8) * 128) + (uint32_t)ks * 256, 128, (TCB_KC / 8) * 128);
                            mma_f16(taddr + dh_col, ad, bd_h, idesc, (j > 0) || (ks > 0));
                        }
                    }
                    mma_commit(&empty_bar[stage]);
                    if (j == TCB_NCHUNK - 1) mma_commit(&acc_full);
                }
                __syncwarp();
                if (++stage == TCB_NSTAGE) { stage = 0; phase ^= 1; }
            }
        }
    }
    } else {
        reg_inc<112>();
        // warp = (TMEM lane quarter lq, column quarter cq): sequence p0 + 32 lq + lane, units [32 cq, 32 cq + 32)
        const int lq = warp & 3, cq = warp >> 2;
        const int r = lq * 32 + lane;
        const int64_t p = p0 + r;
        const bool valid = p < a.P;
        const uint32_t lane_base = taddr + ((uint32_t)(lq * 32) << 16);
        const float scale = a.gscale[0], inv_scale = a.gscale[1];
        const uint32_t dkey = dropout_key(a.seed_dev ? a.drop_seed + *a.seed_dev : a.drop_seed, 0);
        // dx of step `ss` (time tt = L-1-ss) -> dT with the forward's dropout mask
        auto scatter_dx = [&](int ss) {
            const int tt = L - 1 - ss;
            // the forward's mask word of (tile, tt, cq, row): issued before the TMEM load so both latencies overlap
            uint32_t mbits = 0xFFFFFFFFu;
            const bool have_mask = a.dmask != nullptr && a.drop_thr != 0;
            if (have_mask) mbits = __ldg(a.dmask + (((int64_t)tile * L + tt) * 4 + cq) * 128 + r);
            // tcgen05.ld is warp-collective (.sync.aligned): every lane loads, only valid rows scatter
            uint32_t v[32];
            tmem_ld32(lane_base + cq * 32, v);   // dx region = columns [0,128)
            tmem_ld_wait();
            if (valid) {
                const int64_t e = p * L + tt;
                float *dst = a.dT + (int64_t)rowoff_s[r * L + tt] * TC_H + cq * 32;
                const float mul = inv_scale * (a.drop_thr ? a.drop_scale : 1.f);
#pragma unroll
                for (int c8 = 0; c8 < 4; c8++) {
                    if (a.drop_thr && !have_mask) {   // no saved mask (legacy gather path): recompute this chunk's 8 bits
                        const uint64_t pair = ((uint64_t)(e + a.p_base * L) * TC_H + cq * 32 + c8 * 8) >> 1;
                        uint32_t b = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t h = dropout_hash_pair_k(dkey, pair + q);
                            b |= (((h & 0xFFFFu) < a.drop_thr ? 0u : 1u) | ((h >> 16) < a.drop_thr ? 0u : 2u)) << (2 * q);
                        }
                        mbits = (mbits & ~(0xFFu << (c8 * 8))) | (b << (c8 * 8));
                    }
                    if (a.drop_thr) {
                        // most elements are dropped (p = 0.7 upstream): only touch the 16-byte groups that carry anything.  The
                        // scatter is bound by the SM's RED issue rate per active lane (~1.3 cycles each), so the unit is the
                        // widest vector: P(a quad has a kept element) = 0.76 vs 2 x 0.51 for pairs at p = 0.7.
#pragma unroll
                        for (int q = 0; q < 2; q++) {
                            const uint32_t b = (mbits >> (c8 * 8 + 4 * q)) & 15u;
                            if (b) atomicAdd(reinterpret_cast<float4 *>(dst + c8 * 8 + 4 * q),
                                             make_float4((b & 1u) ? __uint_as_float(v[c8 * 8 + 4 * q]) * mul : 0.f,
                                                         (b & 2u) ? __uint_as_float(v[c8 * 8 + 4 * q + 1]) * mul : 0.f,
                                                         (b & 4u) ? __uint_as_float(v[c8 * 8 + 4 * q + 2]) * mul : 0.f,
                                                         (b & 8u) ? __uint_as_float(v[c8 * 8 + 4 * q + 3]) * mul : 0.f));
                        }
                    } else {
                        atomicAdd(reinterpret_cast<float4 *>(dst + c8 * 8), make_float4(__uint_as_float(v[c8 * 8]) * mul, __uint_as_float(v[c8 * 8 + 1]) * mul,
                                                                                     __uint_as_float(v[c8 * 8 + 2]) * mul, __uint_as_float(v[c8 * 8 + 3]) * mul));
                        atomicAdd(reinterpret_cast<float4 *>(dst + c8 * 8 + 4), make_float4(__uint_as_float(v[c8 * 8 + 4]) * mul, __uint_as_float(v[c8 * 8 + 5]) * mul,
                                                                                         __uint_as_float(v[c8 * 8 + 6]) * mul, __uint_as_float(v[c8 * 8 + 7]) * mul));
                    }
                }
            }
            __syncwarp();
        };
        const bool tr = a.trace && blockIdx.x == a.trace_block && tid == 0;
        // factored dh_final: this sequence's two scalars and its output row
        float dh_kk = 1.f, dh_ds = 0.f;
        int64_t grow = 0;
        if (!a.dhfin && valid) { dh_kk = a.dh_k[p]; dh_ds = a.dh_s[p]; grow = p / a.nw; }
        if constexpr (PF) {
        // The saved activations of a half-iteration (4 hidden units x {i, f, g, o, c_t, c_{t-1}}: six 8-byte loads, plus
        // dh_final at the first step) are fetched one half-iteration AHEAD of the gate-gradient math that consumes them —
        // across the K-quarter boundaries and across the step boundary (where the dx scatter sits in between) — so the
        // L2 round trip that used to be exposed once per quarter (the top stall of this kernel in round 1's ncu source
        // view: 24 % of the samples on the first unpack after the loads) overlaps arithmetic.
        struct BwdRaw { uint2 gi, gf, gg, go, cn, cp; float4 dh0; };
        auto load_raw = [&](int s_, int hi_, BwdRaw &x) {
            const int t_ = L - 1 - s_;
            const int cb_ = cq * 4 + (hi_ >> 1), sub = (hi_ & 1) * 4;
            const __half *act_ = a.acts + ((int64_t)tile * L + t_) * ACT_BLOCK_HALFS + r * 8 + cb_ * 1024 + sub;
            x.gi = x.gf = x.gg = x.go = x.cn = x.cp = make_uint2(0u, 0u);
            x.dh0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                x.gi = __ldg(reinterpret_cast<const uint2 *>(act_ + 0 * 16384));
                x.gf = __ldg(reinterpret_cast<const uint2 *>(act_ + 1 * 16384));
                x.gg = __ldg(reinterpret_cast<const uint2 *>(act_ + 2 * 16384));
                x.go = __ldg(reinterpret_cast<const uint2 *>(act_ + 3 * 16384));
                x.cn = __ldg(reinterpret_cast<const uint2 *>(act_ + 4 * 16384));
                if (t_ > 0) x.cp = __ldg(reinterpret_cast<const uint2 *>(act_ - ACT_BLOCK_HALFS + 4 * 16384));
                if (s_ == 0) x.dh0 = a.dhfin ? __ldg(reinterpret_cast<const float4 *>(a.dhfin + p * TC_H + cb_ * 8 + sub))
                                             : __ldg(reinterpret_cast<const float4 *>(a.gdrop + grow * a.gdrop_ld + cb_ * 8 + sub));
            }
        };
        // two half-iterations of activations in flight (n1: the next one, n2: the one after it)
        BwdRaw n1, n2;
        load_raw(0, 0, n1);
        load_raw(0, 1, n2);
        for (int s = 0; s < L; s++) {
            const int t = L - 1 - s;
            if (tr) a.trace[s * 8 + 0] = clock64();
            const int64_t blk = (int64_t)t * a.ntiles + tile;
            unsigned char *da_blk = a.da_img + blk * DA_BLOCK_BYTES;
            // ---- optionally pull the next step's activations towards L2 a whole step ahead (PN_BWD_L2PF=1).  Off by default since the
            //      register prefetch below exists: same-box A/B on the pubmed step, on / off: 1.006 / 0.967 ms per launch, step 2.66 / 2.61 ms ----
            if (valid && t > 0 && a.l2pf) {
                const __half *actp = a.acts + ((int64_t)tile * L + t - 1) * ACT_BLOCK_HALFS + r * 8;
#pragma unroll
                for (int cb = cq * 4; cb < cq * 4 + 4; cb++) {
#pragma unroll
                    for (int q = 0; q < 5; q++) prefetch_l2(actp + q * 16384 + cb * 1024);
                    if (t > 1) prefetch_l2(actp - ACT_BLOCK_HALFS + 4 * 16384 + cb * 1024);
                }
            }
            // ---- the previous step's dx (complete since acc_full) -> dT, before this step's MMAs overwrite the region ----
            if (s > 0) scatter_dx(s - 1);
            if (tr) a.trace[s * 8 + 1] = clock64();
            const uint32_t dh_rd = 128u + (uint32_t)((s - 1) & 1) * 128u;
            // dh / dc of a half-iteration are fetched from TMEM during the previous one (inside a step: the next step's dh
            // is complete only at acc_full), so one TMEM round trip per step is exposed instead of eight
            uint32_t v0n[4] = {0u, 0u, 0u, 0u}, v1n[4] = {0u, 0u, 0u, 0u};
            if (s > 0) {
                tmem_ld4(lane_base + dh_rd + cq * 32, v0n);
                tmem_ld4(lane_base + 384 + cq * 32, v1n);
            }
            // ---- gate gradients: 8 half-iterations of 4 units; two of them complete one K-quarter of the operand tile ----
#pragma unroll 2
            for (int hi = 0; hi < 8; hi++) {
                const BwdRaw cur = n1;
                n1 = n2;
                if (hi < 6) load_raw(s, hi + 2, n2);
                else if (s + 1 < L) load_raw(s + 1, hi - 6, n2);
                const int cb = cq * 4 + (hi >> 1), u0 = cb * 8 + (hi & 1) * 4;
                float dh[4], dc[4];
                if (s == 0) {
                    float4 d0 = cur.dh0;
                    if (!a.dhfin) {
                        const float4 wh = __ldg(reinterpret_cast<const float4 *>(a.attw_h + u0));   // same address across the warp
                        d0 = make_float4(fmaf(dh_kk, d0.x, dh_ds * wh.x), fmaf(dh_kk, d0.y, dh_ds * wh.y), fmaf(dh_kk, d0.z, dh_ds * wh.z),
                                         fmaf(dh_kk, d0.w, dh_ds * wh.w));
                        if (!valid) d0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    dh[0] = d0.x * scale; dh[1] = d0.y * scale; dh[2] = d0.z * scale; dh[3] = d0.w * scale;
                    dc[0] = dc[1] = dc[2] = dc[3] = 0.f;
                } else {
                    tmem_ld_wait_dep(v0n, v1n);
#pragma unroll
                    for (int j = 0; j < 4; j++) { dh[j] = __uint_as_float(v0n[j]); dc[j] = __uint_as_float(v1n[j]); }
                    if (hi < 7) {
                        const int u1 = (cq * 4 + ((hi + 1) >> 1)) * 8 + ((hi + 1) & 1) * 4;
                        tmem_ld4(lane_base + dh_rd + u1, v0n);
                        tmem_ld4(lane_base + 384 + u1, v1n);
                    }
                }
                float gi[4], gf[4], gg[4], go[4], cn[4], cp[4];
                unpack4_f16(cur.gi, gi); unpack4_f16(cur.gf, gf); unpack4_f16(cur.gg, gg);
                unpack4_f16(cur.go, go); unpack4_f16(cur.cn, cn); unpack4_f16(cur.cp, cp);
                float da[2][8];  // [q][.]: packed columns (u0 + 2q + {0,1}) * 4 + gate
                uint32_t dcn_bits[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float tcn = tanh_fast(cn[j]);
                    const float dcn = fmaf(dh[j] * go[j], 1.f - tcn * tcn, dc[j]);
                    const float d_i = dcn * gg[j] * gi[j] * (1.f - gi[j]);
                    const float d_f = dcn * cp[j] * gf[j] * (1.f - gf[j]);
                    const float d_g = dcn * gi[j] * (1.f - gg[j] * gg[j]);
                    const float d_o = dh[j] * tcn * go[j] * (1.f - go[j]);
                    dcn_bits[j] = __float_as_uint(dcn * gf[j]);
                    da[j >> 1][(j & 1) * 4 + 0] = d_i; da[j >> 1][(j & 1) * 4 + 1] = d_f;
                    da[j >> 1][(j & 1) * 4 + 2] = d_g; da[j >> 1][(j & 1) * 4 + 3] = d_o;
                }
                tmem_st4(lane_base + 384 + u0, dcn_bits);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    const int chunk = cb * 4 + (hi & 1) * 2 + q;   // 8 packed gate columns = 16 bytes
                    const uint4 pk = pack8_f16(da[q]);
                    *reinterpret_cast<uint4 *>(As + canon_off(r, chunk * 8, TC_G4)) = pk;          // dgrad operand
                    *reinterpret_cast<uint4 *>(da_blk + (size_t)chunk * 2048 + r * 16) = pk;       // wgrad operand image
                }
                if (hi & 1) {   // this quarter of the operand tile is written: let the tensor core start on it
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(&a_ready[hi >> 1]);
                }
            }
            tmem_st_wait();
            if (tr) a.trace[s * 8 + 2] = clock64();
            mbar_wait(&acc_full, s & 1);
            tc_fence_after();
            if (tr) a.trace[s * 8 + 3] = clock64();
            if (tr) a.trace[s * 8 + 4] = clock64();
        }
        } else {
        for (int s = 0; s < L; s++) {
            const int t = L - 1 - s;
            if (tr) a.trace[s * 8 + 0] = clock64();
            const int64_t blk = (int64_t)t * a.ntiles + tile;
            unsigned char *da_blk = a.da_img + blk * DA_BLOCK_BYTES;
            const __half *act = a.acts + ((int64_t)tile * L + t) * ACT_BLOCK_HALFS + r * 8;
            const __half *actp = act - ACT_BLOCK_HALFS;  // step t-1 (only read when t > 0)
            // ---- pull the next step's activations towards L2 (a whole step ahead) ----
            if (valid && t > 0) {
#pragma unroll
                for (int cb = cq * 4; cb < cq * 4 + 4; cb++) {
#pragma unroll
                    for (int q = 0; q < 5; q++) prefetch_l2(actp + q * 16384 + cb * 1024);
                    if (t > 1) prefetch_l2(actp - ACT_BLOCK_HALFS + 4 * 16384 + cb * 1024);
                }
            }
            // ---- the previous step's dx (complete since acc_full) -> dT, before this step's MMAs overwrite the region ----
            if (s > 0) scatter_dx(s - 1);
            if (tr) a.trace[s * 8 + 1] = clock64();
            const uint32_t dh_rd = 128u + (uint32_t)((s - 1) & 1) * 128u;
            // ---- gate gradients, one K-quarter of the operand tile per iteration ----
#pragma unroll 1
            for (int cb = cq * 4; cb < cq * 4 + 4; cb++) {
                const int u0 = cb * 8;
                float dh[8], dc[8];
                if (s == 0) {
#pragma unroll
                    for (int j = 0; j < 8; j++) { dh[j] = 0.f; dc[j] = 0.f; }
                    if (valid) {
                        float4 x0, x1;
                        if (a.dhfin) {
                            x0 = __ldg(reinterpret_cast<const float4 *>(a.dhfin + p * TC_H + u0));
                            x1 = __ldg(reinterpret_cast<const float4 *>(a.dhfin + p * TC_H + u0 + 4));
                        } else {   // dh_final = k * (row's pooled-half gradient) + dscore * attw_w[0:H]  (see PoolArgs)
                            const float4 g0 = __ldg(reinterpret_cast<const float4 *>(a.gdrop + grow * a.gdrop_ld + u0));
                            const float4 g1 = __ldg(reinterpret_cast<const float4 *>(a.gdrop + grow * a.gdrop_ld + u0 + 4));
                            const float4 w0 = __ldg(reinterpret_cast<const float4 *>(a.attw_h + u0));
                            const float4 w1 = __ldg(reinterpret_cast<const float4 *>(a.attw_h + u0 + 4));
                            x0 = make_float4(fmaf(dh_kk, g0.x, dh_ds * w0.x), fmaf(dh_kk, g0.y, dh_ds * w0.y), fmaf(dh_kk, g0.z, dh_ds * w0.z), fmaf(dh_kk, g0.w, dh_ds * w0.w));
                            x1 = make_float4(fmaf(dh_kk, g1.x, dh_ds * w1.x), fmaf(dh_kk, g1.y, dh_ds * w1.y), fmaf(dh_kk, g1.z, dh_ds * w1.z), fmaf(dh_kk, g1.w, dh_ds * w1.w));
                        }
                        dh[0] = x0.x * scale; dh[1] = x0.y * scale; dh[2] = x0.z * scale; dh[3] = x0.w * scale;
                        dh[4] = x1.x * scale; dh[5] = x1.y * scale; dh[6] = x1.z * scale; dh[7] = x1.w * scale;
                    }
                } else {
                    uint32_t v0[8], v1[8];
                    tmem_ld8(lane_base + dh_rd + u0, v0);
                    tmem_ld8(lane_base + 384 + u0, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 8; j++) { dh[j] = __uint_as_float(v0[j]); dc[j] = __uint_as_float(v1[j]); }
                }
                float gi[8], gf[8], gg[8], go[8], cn[8], cp[8];
#pragma unroll
                for (int j = 0; j < 8; j++) { gi[j] = gf[j] = gg[j] = go[j] = cn[j] = cp[j] = 0.f; }
                if (valid) {
                    const int ub = cb * 1024;
                    unpack8_f16(__ldg(reinterpret_cast<const uint4 *>(act + 0 * 16384 + ub)), gi);
                    unpack8_f16(__ldg(reinterpret_cast<const uint4 *>(act + 1 * 16384 + ub)), gf);
                    unpack8_f16(__ldg(reinterpret_cast<const uint4 *>(act + 2 * 16384 + ub)), gg);
                    unpack8_f16(__ldg(reinterpret_cast<const uint4 *>(act + 3 * 16384 + ub)), go);
                    unpack8_f16(__ldg(reinterpret_cast<const uint4 *>(act + 4 * 16384 + ub)), cn);
                    if (t > 0) unpack8_f16(__ldg(reinterpret_cast<const uint4 *>(actp + 4 * 16384 + ub)), cp);
                }
                float da[4][8];  // [q][.]: packed columns (u0 + 2q + {0,1}) * 4 + gate
                uint32_t dcn_bits[8];
#pragma unroll
                for (int j = 0; j < 8; j++) {
                    const float tcn = tanh_fast(cn[j]);
                    const float dcn = fmaf(dh[j] * go[j], 1.f - tcn * tcn, dc[j]);
                    const float d_i = dcn * gg[j] * gi[j] * (1.f - gi[j]);
                    const float d_f = dcn * cp[j] * gf[j] * (1.f - gf[j]);
                    const float d_g = dcn * gi[j] * (1.f - gg[j] * gg[j]);
                    const float d_o = dh[j] * tcn * go[j] * (1.f - go[j]);
                    dcn_bits[j] = __float_as_uint(dcn * gf[j]);
                    da[j >> 1][(j & 1) * 4 + 0] = d_i; da[j >> 1][(j & 1) * 4 + 1] = d_f;
                    da[j >> 1][(j & 1) * 4 + 2] = d_g; da[j >> 1][(j & 1) * 4 + 3] = d_o;
                }
                tmem_st8(lane_base + 384 + u0, dcn_bits);
#pragma unroll
                for (int q = 0; q < 4; q++) {
                    const uint4 pk = pack8_f16(da[q]);
                    *reinterpret_cast<uint4 *>(As + canon_off(r, cb * 32 + q * 8, TC_G4)) = pk;          // dgrad operand
                    *reinterpret_cast<uint4 *>(da_blk + (size_t)(cb * 4 + q) * 2048 + r * 16) = pk;       // wgrad operand image
                }
                // this quarter of the operand tile is written: let the tensor core start on it
                fence_proxy_async();
                tc_fence_before();
                mbar_arrive(&a_ready[cb - cq * 4]);
            }
            tmem_st_wait();
            if (tr) a.trace[s * 8 + 2] = clock64();
            mbar_wait(&acc_full, s & 1);
            tc_fence_after();
            if (tr) a.trace[s * 8 + 3] = clock64();
            if (tr) a.trace[s * 8 + 4] = clock64();
        }
        }
        scatter_dx(L - 1);   // the last step's dx (nothing left to overlap it with)
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CWARPS) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

// ---------------------------------------------------------------------------
// weight gradients: dWp[k][col] (+)= sum over blocks  XH_blk^T . dA_blk
// CTA = one 128 (k) x 256 (col) output tile x one split of the block list;
// operands are bulk-copied images (MN-major UMMA).  The four tiles of a split
// form a thread-block cluster (1 x 4): tile (mt, nt) needs XH half mt and dA
// half nt, so each CTA fetches ONE quarter of the block (16 KB of XH, 32 KB of
// dA) and multicasts it to the two CTAs that use it -- every byte of the
// images is read from HBM/L2 once per block instead of twice.
// ---------------------------------------------------------------------------
struct WgradTcArgs {
    const unsigned char *xh_img, *da_img;
    int64_t nblocks;
    const float *gscale;
    float *dWp;  // [256][512] fp32, atomically accumulated
    float *dbp;  // [512]      fp32, atomically accumulated (bias gradient = column sums of da)
    int do_bias; // 0: skip the bias rows (split-precision mode: the pass that pairs the lo term of [x|h] with da again)
};
constexpr int WG_STAGE_BYTES = 32768 + 65536;
constexpr int WG_ONES_BYTES = 32768;   // constant A operand [128 (m)][128 (r)] with row m = 0 all ones
constexpr int WG_SMEM = 2 * WG_STAGE_BYTES + WG_ONES_BYTES + 1024;

__global__ void __launch_bounds__(192, 1) k_wgrad_tc(WgradTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *ones = base + 2 * WG_STAGE_BYTES;
    __shared__ uint64_t full_bar[2], empty_bar[2], acc_full;
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const uint32_t rank = cluster_ctarank();          // == blockIdx.y (cluster = the 4 tiles of one split)
    const int mt = rank >> 1, nt = rank & 1;
    // who shares what: XH half mt is used by ranks {2mt, 2mt+1}; dA half nt by ranks {nt, 2+nt}
    const uint16_t xh_mask = (uint16_t)(3u << (2 * mt)), da_mask = (uint16_t)(5u << nt);
    const uint16_t peers = (uint16_t)(xh_mask | da_mask);   // self + the two CTAs whose copies land in my stages
    if (tid == 0) {
        for (int s = 0; s < 2; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 3); }
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 4) tmem_alloc<512>(&s_taddr);
    // ones operand, MN-major image like the XH blocks: chunk (m8, r) at m8*2048 + r*16 holds m = 8 m8 .. 8 m8 + 7
    for (int i = tid; i < WG_ONES_BYTES / 16; i += 192) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (i < 128) v.x = 0x00003C00u;   // m8 = 0: element m = 0 is fp16 1.0, m = 1..7 are 0
        reinterpret_cast<uint4 *>(ones)[i] = v;
    }
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // every CTA's mbarriers are initialised before any peer's copy or commit can reach them
    tc_fence_after();
    const uint32_t taddr = s_taddr;   // cols [0,256): dWp tile, cols [256,384): bias rows (lane 0 = column sums of 128 da columns)
    const int64_t first = blockIdx.x, step = gridDim.x;
    if (warp == 5) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int64_t b = first; b < a.nblocks; b += step) {
                unsigned char *dst = base + stage * WG_STAGE_BYTES;
                mbar_wait(&empty_bar[stage], phase ^ 1);   // this stage is free in all three CTAs my copies land in
                mbar_arrive_expect_tx(&full_bar[stage], WG_STAGE_BYTES);
                // my quarter of XH half mt (piece nt) -> both CTAs of row mt
                bulk_g2s_mc(dst + nt * 16384, a.xh_img + b * XH_BLOCK_BYTES + (size_t)mt * 32768 + (size_t)nt * 16384, 16384,
                            &full_bar[stage], xh_mask);
                // my quarter of dA half nt (piece mt) -> both CTAs of column nt
                bulk_g2s_mc(dst + 32768 + mt * 32768, a.da_img + b * DA_BLOCK_BYTES + (size_t)nt * 65536 + (size_t)mt * 32768, 32768,
                            &full_bar[stage], da_mask);
                if (++stage == 2) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 4) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16_mn(128, 256), idesc_b = make_idesc_f16_mn(128, 128);
            uint32_t stage = 0, phase = 0;
            bool first_blk = true;
            for (int64_t b = first; b < a.nblocks; b += step) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t abase = smem_u32(base + stage * WG_STAGE_BYTES), bbase = abase + 32768, obase = smem_u32(ones);
#pragma unroll
                for (int ks = 0; ks < 8; ks++) {  // 128 rows of the block = 8 K-steps of 16
                    const uint64_t ad = make_smem_desc(abase + ks * 256, 128, 2048);
                    const uint64_t bd = make_smem_desc(bbase + ks * 256, 128, 2048);
                    mma_f16(taddr, ad, bd, idesc, !first_blk || ks > 0);
                    // bias: ones . dA over this CTA's share of the columns (mt picks which 128 of the 256)
                    if (a.do_bias)
                        mma_f16(taddr + 256, make_smem_desc(obase + ks * 256, 128, 2048),
                                make_smem_desc(bbase + mt * 32768 + ks * 256, 128, 2048), idesc_b, !first_blk || ks > 0);
                }
                first_blk = false;
                mma_commit_mc(&empty_bar[stage], peers);
                if (++stage == 2) { stage = 0; phase ^= 1; }
            }
            mma_commit(&acc_full);
        }
    } else {
        mbar_wait(&acc_full, 0);
        tc_fence_after();
        if (first < a.nblocks) {
            const float inv_scale = a.gscale[1];
            float *dst = a.dWp + (int64_t)(mt * 128 + tid) * TC_G4 + nt * 256;
            for (int c0 = 0; c0 < 256; c0 += 32) {
                uint32_t v[32];
                tmem_ld32(taddr + ((uint32_t)(warp * 32) << 16) + (uint32_t)c0, v);
                tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    atomicAdd(reinterpret_cast<float4 *>(dst + c0 + j),
                              make_float4(__uint_as_float(v[j]) * inv_scale, __uint_as_float(v[j + 1]) * inv_scale,
                                          __uint_as_float(v[j + 2]) * inv_scale, __uint_as_float(v[j + 3]) * inv_scale));
            }
            if (warp == 0 && a.do_bias) {   // TMEM lane 0 of the second accumulator = sum over rows of da[:, col]
                for (int c0 = 0; c0 < 128; c0 += 32) {
                    uint32_t v[32];
                    tmem_ld32(taddr + 256 + (uint32_t)c0, v);
                    tmem_ld_wait();
                    if (lane == 0)
#pragma unroll
                        for (int j = 0; j < 32; j++) atomicAdd(a.dbp + nt * 256 + mt * 128 + c0 + j, __uint_as_float(v[j]) * inv_scale);
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    cluster_sync_all();   // no CTA leaves while a peer's multicast copy or commit may still target it
    if (warp == 4) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

// LSTM backward in tensor-core mode: fills w.dWp [2H][4H] and w.dbp [4H] (packed
// layouts, same as the fp32 path) and scatters dx into w.dT.
int pagg_tc_backward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm,
                          const int32_t *seq_node, const uint8_t *seq_type, const PaggChunk &ck, bool factored_dh, cudaStream_t st,
                          cudaEvent_t after_bptt)
{
    if (c.H != TC_H) {
        set_error("PN_MATH_TC / TC3: the tensor-core LSTM is built for hidden size 128 (got %d)", c.H);
        return PN_ERR_UNSUPPORTED;
    }
    const int64_t P = ck.Sc * c.nw, ntiles = ceil_div(P, 128), nblocks = ntiles * c.L;
    // max |dh_final| was accumulated into gscale[0] (as float bits) by k_pool_bwd
    k_make_scale<<<1, 1, 0, st>>>(w.gscale);
    PN_LAUNCH_CHECK();
    unsigned char *WhT = reinterpret_cast<unsigned char *>(w.WpT);  // 256 KB of the 512 KB fp32 region
    const bool split = c.math_mode == PN_MATH_TC3;   // split-precision mode: two fp16 terms of W' (all 512 KB of the region)
    if (ck.first) {
        if (split) { const int rc = launch_pack_lstm_tc3_bwd(st, prm.w_ih, prm.w_hh, WhT); if (rc) return rc; }
        else {
            k_pack_lstm_tc_bwd<<<(unsigned)ceil_div(TC_G4 * TC_K2, 256), 256, 0, st>>>(prm.w_ih, prm.w_hh, WhT);
            PN_LAUNCH_CHECK();
        }
    }
    uint32_t thr = 0; float scale = 1.f;
    if (c.training && c.dropout_p > 0.f) {
        thr = dropout_threshold(c.dropout_p);
        scale = c.dropout_p >= 1.f ? 0.f : 1.f / (1.f - c.dropout_p);
    }
    unsigned char *da_img = reinterpret_cast<unsigned char *>(w.dA), *xh_img = reinterpret_cast<unsigned char *>(w.XH);
    LstmBwdTcArgs a{w.T, seq_node, seq_type, WhT, reinterpret_cast<const __half *>(w.acts), factored_dh ? nullptr : w.dhfin,
                    w.att, w.score, w.dcat + c.H, prm.attw_w, 2 * (int64_t)c.H, c.nw, w.gscale, w.dT,
                    da_img, xh_img, (tc_legacy_gemm() && !split) ? nullptr : reinterpret_cast<const uint32_t *>(w.dmask), P, ntiles, c.L, thr, scale,
                    c.dropout_seed, c.dropout_seed_dev, ck.s0 * c.nw, nullptr, 0,
                    w.acts, da_img + nblocks * DA_BLOCK_BYTES,   // split-precision mode: fp32 activations, second term of the da images
                    []{ const char *e = getenv("PN_BWD_L2PF"); return (e && e[0] == '1') ? 1 : 0; }()};
    long long *d_trace = nullptr;
    const char *trace_env = getenv("PN_TRACE_BWD");
    if (trace_env) {
        PN_CUDA(cudaMalloc(&d_trace, sizeof(long long) * 8 * 64));
        PN_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(long long) * 8 * 64, st));
        a.trace = d_trace;
        a.trace_block = atoi(trace_env);
    }
    const int bwd_smem = TCB_SMEM_FIXED + 128 * c.L * (int)sizeof(int32_t);
    static const bool bwd_pf = [] { const char *e = getenv("PN_BWD_PF"); return !(e && e[0] == '0'); }();
    constexpr int bwd_smem_max = TCB_SMEM_FIXED + 128 * TC_MAX_L * (int)sizeof(int32_t);
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_max));
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, bwd_smem_max));
    {
        ProfScope prof("lstm_bwd", st);
        if (split) { const int rc = launch_lstm_bwd_tc3(a, ntiles, bwd_smem, bwd_smem_max, st); if (rc) return rc; }
        else if (bwd_pf) k_lstm_bwd_tc<true><<<(unsigned)ntiles, TCB_THREADS, bwd_smem, st>>>(a);
        else k_lstm_bwd_tc<false><<<(unsigned)ntiles, TCB_THREADS, bwd_smem, st>>>(a);
    }
    PN_LAUNCH_CHECK();
    if (d_trace) {  // debug only
        long long h[8 * 64];
        PN_CUDA(cudaStreamSynchronize(st));
        PN_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        const long long t00 = h[0];
        for (int q = 0; q < c.L && q < 64; q++)
            fprintf(stderr, "[trace bwd cta %d] s=%d  begin %lld | prev dx scattered %lld | 4 da quarters arrived %lld | acc_full seen %lld | %lld\n",
                    a.trace_block, q, h[q * 8 + 0] - t00, h[q * 8 + 1] - t00, h[q * 8 + 2] - t00, h[q * 8 + 3] - t00, h[q * 8 + 4] - t00);
    }
    if (after_bptt) PN_CUDA(cudaEventRecord(after_bptt, st));   // dT is complete: the caller forks the projections' backward here
    // dWp / dbp are zeroed once per call by pn_pagg_backward: every chunk adds into them
    WgradTcArgs wa{xh_img, da_img, nblocks, w.gscale, w.dWp, w.dbp, 1};
    PN_CUDA(cudaFuncSetAttribute(k_wgrad_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, WG_SMEM));
    {
        ProfScope prof("wgrad", st);
        // one cluster of 4 CTAs per split; as many splits as clusters fit on the device at once (a second wave would double the time)
        cudaLaunchConfig_t lc = {};
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 1; attr[0].val.clusterDim.y = 4; attr[0].val.clusterDim.z = 1;
        lc.gridDim = dim3(1, 4); lc.blockDim = dim3(192); lc.dynamicSmemBytes = WG_SMEM; lc.stream = st;
        lc.attrs = attr; lc.numAttrs = 1;
        static int max_clusters = 0;
        if (max_clusters == 0) {
            int mc = 0;
            lc.gridDim = dim3(64, 4);
            PN_CUDA(cudaOccupancyMaxActiveClusters(&mc, k_wgrad_tc, &lc));
            max_clusters = std::max(1, std::min(mc, 64));
            if (getenv("PN_TRACE_BWD")) fprintf(stderr, "[wgrad] max active clusters of 4 = %d\n", mc);
        }
        // When the projections' backward runs beside this kernel (after_bptt: the caller forked it onto a side stream), leave it
        // 6 clusters = 24 more SMs: its GEMM CTAs (64 KB of shared memory) cannot share an SM with a 230 KB wgrad CTA.  Same-box
        // sweep on the pubmed step, clusters -> ms/step: 33 2.689, 31 2.698, 29 2.675, 27 2.647, 25 2.674 (this kernel alone slows
        // from 0.465 to 0.534 ms at 27, the step gains).  PN_WGRAD_CLUSTERS overrides.
        // (Split-precision mode: its three passes are the longer branch by far -- same-box sweep 33 / 30 / 27 / 24 clusters -> 5.17 / 5.20 /
        // 5.29 / 5.39 ms per step -- so it keeps every cluster.)
        int cap = (after_bptt && !split) ? std::max(1, max_clusters - 6) : max_clusters;
        if (const char *e = getenv("PN_WGRAD_CLUSTERS")) { const int v = atoi(e); if (v >= 1) cap = std::min(max_clusters, v); }
        const unsigned nsplit = (unsigned)std::max<int64_t>(1, std::min<int64_t>(cap, nblocks));
        lc.gridDim = dim3(nsplit, 4);
        PN_CUDA(cudaLaunchKernelEx(&lc, k_wgrad_tc, wa));
        if (split) {
            // split-precision mode: dW = [x|h]_hi^T da_hi (above) + [x|h]_lo^T da_hi + [x|h]_hi^T da_lo, all accumulated into dWp; the
            // bias gradient is the column sums of da_hi (above) + da_lo
            WgradTcArgs w2{xh_img + nblocks * XH_BLOCK_BYTES, da_img, nblocks, w.gscale, w.dWp, w.dbp, 0};
            PN_CUDA(cudaLaunchKernelEx(&lc, k_wgrad_tc, w2));
            WgradTcArgs w3{xh_img, da_img + nblocks * DA_BLOCK_BYTES, nblocks, w.gscale, w.dWp, w.dbp, 1};
            PN_CUDA(cudaLaunchKernelEx(&lc, k_wgrad_tc, w3));
        }
    }
    PN_LAUNCH_CHECK();
    return PN_OK;
}

// ---------------------------------------------------------------------------
// projections around the LSTM in tensor-core mode
// ---------------------------------------------------------------------------
// {s, 1/s} with s a power of two that maps max |v| (float bits in slot[0], from the column-sum sweeps) to (2^13, 2^14]:
// the fp16 copy of a raw gradient then keeps 11 significant bits down to 2^-28 of its largest entry.
__global__ void k_absmax_scale(float *slot)
{
    const float m = __uint_as_float(*reinterpret_cast<unsigned int *>(slot));
    float sc = 1.f;
    if (m > 0.f && isfinite(m)) sc = exp2f(14.f - ceilf(log2f(m)));
    slot[0] = sc;
    slot[1] = 1.f / sc;
}

// K0: Z = act(X W0^T + b0) (:175 / :242-243);  K1: T[n, d*H + j] = act(Z nets_w[d]^T + nets_b[d]) (:185-191 / :249-257, hoisted)
// Default: the TMA-fed GEMM of gemm_tma.cu on fp16 copies (X: the caller's cached cfg.x16 or converted here; Z: written
// by the fc0 epilogue; weights: converted here, 0.3 MB), which also writes the fp16 table the LSTM kernel gathers.
// PN_TC_GEMM=legacy: the round-1 kernel that converts fp32 operands while staging them (kept for A/B runs).
int pagg_tc_project(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const float *X, cudaStream_t st)
{
    const int64_t N = c.N, F = c.F, H = c.H, L = c.L;
    const bool homo = c.variant == PN_VARIANT_HOMO;
    const bool split3 = c.math_mode == PN_MATH_TC3;   // split-precision mode: three fp16 passes per product, fp32 operands in and out
    if (tc_legacy_gemm() || split3) {
        int rc = launch_gemm_tc(st, X, F, prm.fc0_w, F, w.Z, H, N, H, F, prm.fc0_b, homo, false, false, false, false, 1, nullptr, split3);
        if (rc) return rc;
        return launch_gemm_tc(st, w.Z, H, prm.nets_w, H, w.T, L * H, N, L * H, H, prm.nets_b, homo, false, false, false, false, 1, nullptr, split3);
    }
    const int64_t F16 = (F + 7) / 8 * 8;
    __half *x16 = reinterpret_cast<__half *>(w.x16), *z16 = reinterpret_cast<__half *>(w.z16);
    __half *w0 = reinterpret_cast<__half *>(w.w16), *wn = w0 + H * F16, *t16 = reinterpret_cast<__half *>(w.t16);
    int rc;
    if (c.x16) x16 = const_cast<__half *>(reinterpret_cast<const __half *>(c.x16));
    else if ((rc = launch_f32_to_f16_rows(st, X, N, F, F, x16, F16))) return rc;
    if ((rc = launch_f32_to_f16_rows(st, prm.fc0_w, H, F, F, w0, F16))) return rc;
    if ((rc = launch_f32_to_f16_rows(st, prm.nets_w, L * H, H, H, wn, H))) return rc;
    if ((rc = launch_gemm_tma(st, x16, w0, N, H, F16, prm.fc0_b, homo, w.Z, H, z16, H, 1.f))) return rc;
    float tscale = 1.f;   // the LSTM kernel's dropout keeps x * 1/(1-p): folded into the fp16 table it gathers
    if (c.training && c.dropout_p > 0.f) tscale = c.dropout_p >= 1.f ? 0.f : 1.f / (1.f - c.dropout_p);
    return launch_gemm_tma(st, z16, wn, N, L * H, H, prm.nets_b, homo, w.T, L * H, t16, L * H, tscale);
}

// backward of the two projections: nets_w/nets_b, fc0_w/fc0_b from dT, dZ (w.dT / w.dZ hold the scattered gradients).
// The raw gradients are staged as fp16 times a power-of-two scale taken from their largest entry (the column-sum
// sweeps that produce the bias gradients also take that maximum), fp32 accumulation.
int pagg_tc_project_backward(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const float *X,
                             const pn_pagg_params &g, cudaStream_t st)
{
    const int64_t N = c.N, F = c.F, H = c.H, L = c.L;
    const bool homo = c.variant == PN_VARIANT_HOMO;
    float *sc_T = w.gscale + 2, *sc_Z = w.gscale + 4;
    const bool split3 = c.math_mode == PN_MATH_TC3;   // split-precision mode: the scaled gradients travel as two fp16 terms
    PN_CUDA(cudaMemsetAsync(w.gscale + 2, 0, sizeof(float) * 4, st));
    int rc;
    // ---- table projection: T = act(Z Wn^T + bn) ----  (homo: ReLU backward fused with the bias gradient sweep)
    rc = homo ? launch_relu_mask_colsum(st, w.dT, w.T, N, L * H, g.nets_b, reinterpret_cast<unsigned int *>(sc_T))
              : launch_colsum(st, w.dT, N, L * H, L * H, g.nets_b, false, reinterpret_cast<unsigned int *>(sc_T));
    if (rc) return rc;
    k_absmax_scale<<<1, 1, 0, st>>>(sc_T);
    PN_LAUNCH_CHECK();
    {
        // split the reduction over N so the (L*H/128) x (H/128) output tiles fill one wave of the 148 SMs
        const int splitk = (int)std::max<int64_t>(1, std::min<int64_t>(148 / std::max<int64_t>(1, ceil_div(L * H, 128) * ceil_div(H, 128)), N / 256));
        if (splitk > 1) PN_CUDA(cudaMemsetAsync(g.nets_w, 0, sizeof(float) * (size_t)(L * H * H), st));
        // dWn[m][i] = sum_n dT[n][m] Z[n][i]: both operands have the reduction index as their row (MN-major)
        rc = launch_gemm_tc(st, w.dT, L * H, w.Z, H, g.nets_w, H, L * H, H, N, nullptr, false, true, true, false, false, splitk, sc_T, split3);
        if (rc) return rc;
    }
    rc = launch_gemm_tc(st, w.dT, L * H, prm.nets_w, H, w.dZ, H, N, H, L * H, nullptr, false, false, true, false, true, 1, sc_T, split3);  // dZ += dT Wn
    if (rc) return rc;
    // ---- fc0: Z = act(X W0^T + b0) ----
    rc = homo ? launch_relu_mask_colsum(st, w.dZ, w.Z, N, H, g.fc0_b, reinterpret_cast<unsigned int *>(sc_Z))
              : launch_colsum(st, w.dZ, N, H, H, g.fc0_b, false, reinterpret_cast<unsigned int *>(sc_Z));
    if (rc) return rc;
    k_absmax_scale<<<1, 1, 0, st>>>(sc_Z);
    PN_LAUNCH_CHECK();
    {
        const int splitk = (int)std::max<int64_t>(1, std::min<int64_t>(148 / std::max<int64_t>(1, ceil_div(H, 128) * ceil_div(F, 128)), N / 256));
        if (splitk > 1) PN_CUDA(cudaMemsetAsync(g.fc0_w, 0, sizeof(float) * (size_t)(H * F), st));
        rc = launch_gemm_tc(st, w.dZ, H, X, F, g.fc0_w, F, H, F, N, nullptr, false, true, true, false, false, splitk, sc_Z, split3);  // dW0 = dZ^T X
    }
    return rc;
}

}  // namespace pn

// stand-alone entry for the tcgen05 GEMM (unit-tested against fp64 math).  flags: bit0 A given as [K][M],
// bit1 B given as [K][N], bit2 bf16 operands, bit3 accumulate into C, bit4 split precision (three fp16 passes); splitk >= 1.
extern "C" int pn_debug_gemm_tc(const float *A, const float *B, float *C, const float *bias, int64_t M, int64_t N,
                                int64_t K, int relu, void *stream)
{
    return pn::launch_gemm_tc((cudaStream_t)stream, A, K, B, K, C, N, M, N, K, bias, relu != 0, false, false, false, false, 1, nullptr);
}
extern "C" int pn_debug_gemm_tc2(const float *A, const float *B, float *C, int64_t M, int64_t N, int64_t K, int flags,
                                 int splitk, void *stream)
{
    const bool a_mn = flags & 1, b_mn = flags & 2, bf16 = flags & 4, acc = flags & 8, split3 = flags & 16;
    return pn::launch_gemm_tc((cudaStream_t)stream, A, a_mn ? M : K, B, b_mn ? N : K, C, N, M, N, K, nullptr, false, a_mn,
                              b_mn, bf16, acc, splitk, nullptr, split3);
}
