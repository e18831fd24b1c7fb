// TMA-fed tcgen05 GEMM for the two forward projections of PAGG (fc0 and the per-distance table):
//     C[M,N] = act(A[M,K] . B[N,K]^T + bias)      A, B fp16 row-major (K contiguous), fp32 accumulate in TMEM
// Warp-specialised: warp 0 = TMA producer (cp.async.bulk.tensor.2d, 128-byte swizzle, zero fill past the edges),
// warp 1 = MMA issuer (one thread, tcgen05.mma kind::f16 M128 N128 K16) + TMEM owner, warps 2-5 = epilogue
// (tcgen05.ld -> bias / ReLU -> fp32 rows and, optionally, a scaled fp16 copy for the next consumer).
// 3-stage ring of 32 KB (A tile 128 x 64 + B tile 128 x 64) -> two CTAs per SM.
//
// Why fp16 operands in global memory: X never changes during training (PathNet_run.py:300), so its fp16 copy is
// made once (pn_pagg_pack_x16) instead of re-reading 4 bytes and converting on every step; Z's fp16 copy is written
// by the fc0 epilogue for the table projection; the table's fp16 copy (pre-multiplied by the dropout scale) is what
// the LSTM kernel gathers.
#include <cuda.h>
#include <cuda_fp16.h>

#include "common.h"
#include "tc_util.cuh"

namespace pn {
using namespace tc;

// ---------------------------------------------------------------- fp32 -> fp16 rows, K padded to a multiple of 8
__global__ void k_f32_to_f16_rows(const float *__restrict__ src, int64_t rows, int64_t K, int64_t ld_src,
                                  __half *__restrict__ dst, int64_t K16)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;   // one 8-element group per thread
    const int64_t g8 = K16 / 8;
    if (i >= rows * g8) return;
    const int64_t r = i / g8, k0 = (i % g8) * 8;
    float f[8];
#pragma unroll
    for (int j = 0; j < 8; j++) f[j] = (k0 + j < K) ? src[r * ld_src + k0 + j] : 0.f;
    *reinterpret_cast<uint4 *>(dst + r * K16 + k0) = pack8_f16(f);
}

int launch_f32_to_f16_rows(cudaStream_t st, const float *src, int64_t rows, int64_t K, int64_t ld_src, __half *dst, int64_t K16)
{
    if (rows <= 0) return PN_OK;
    const int64_t n = rows * (K16 / 8);
    k_f32_to_f16_rows<<<(unsigned)ceil_div(n, 256), 256, 0, st>>>(src, rows, K, ld_src, dst, K16);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

// ---------------------------------------------------------------- tensor maps
// 2-D map of a row-major fp16 matrix [rows][K16] with box {64 (K), 128 (rows)} and 128-byte swizzle
// (cuTensorMapEncodeTiled is resolved through the runtime at first use, so the library has no load-time dependency on
// libcuda.so.1 and still imports on a machine without a driver — the CPU-side tests load it to check its exports.)
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *, const cuuint64_t *,
                                  const cuuint32_t *, const cuuint32_t *, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeTiledFn encode_tiled_fn()
{
    static EncodeTiledFn fn = [] {
        void *p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) p = nullptr;
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

static int make_tmap_f16(CUtensorMap *tm, const void *ptr, int64_t rows, int64_t K16)
{
    EncodeTiledFn enc = encode_tiled_fn();
    if (!enc) { set_error("cuTensorMapEncodeTiled is not available from this CUDA driver"); return PN_ERR_CUDA; }
    const cuuint64_t dims[2] = {(cuuint64_t)K16, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)K16 * 2};   // bytes between rows; multiple of 16 because K16 % 8 == 0
    const cuuint32_t box[2] = {64, 128};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void *>(ptr), dims, strides, box, estr,
                           CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                           CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) for a [%lld][%lld] fp16 matrix at %p", (int)r, (long long)rows,
                  (long long)K16, ptr);
        return PN_ERR_CUDA;
    }
    return PN_OK;
}

__device__ __forceinline__ void tma_load_2d(void *dst_smem, const CUtensorMap *tm, int32_t c0, int32_t c1, uint64_t *bar)
{
    asm volatile("cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];"
                 ::"r"(smem_u32(dst_smem)), "l"(reinterpret_cast<uint64_t>(tm)), "r"(c0), "r"(c1), "r"(smem_u32(bar)) : "memory");
}

// K-major operand tile written by TMA with 128-byte swizzle: rows of 128 bytes (64 fp16), 8-row groups 1024 bytes apart
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);   // [0,14)  start address >> 4
    d |= (uint64_t)1 << 16;                    // [16,30) leading byte offset (unused for swizzled K-major; canonical value 1)
    d |= (uint64_t)(1024 >> 4) << 32;          // [32,46) stride byte offset: 8 rows x 128 B
    d |= (uint64_t)1 << 46;                    // [46,48) descriptor version = 1 (Blackwell)
    d |= (uint64_t)2 << 61;                    // [61,64) layout type SWIZZLE_128B
    return d;
}

struct GemmTmaArgs {
    int64_t M, N, K;       // K = the padded K16 of both operands
    const float *bias;     // [N] or null
    int relu;
    float *C; int64_t ldc;           // fp32 output or null
    __half *C16; int64_t ldc16;      // fp16 output (value * c16_scale) or null
    float c16_scale;
};

constexpr int GT_STAGES = 3, GT_STAGE_BYTES = 32768, GT_SMEM = GT_STAGES * GT_STAGE_BYTES + 1024;

__global__ void __launch_bounds__(192) k_gemm_tma(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
                                                  GemmTmaArgs g)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t full_bar[GT_STAGES], empty_bar[GT_STAGES], acc_full;
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t m0 = (int64_t)blockIdx.x * 128, n0 = (int64_t)blockIdx.y * 128;   // row tiles on grid.x: 10^7-row operands exceed grid.y's 65 535
    const int nk = (int)((g.K + 63) / 64);
    if (tid == 0) {
        for (int s = 0; s < GT_STAGES; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == 1) tmem_alloc<128>(&s_taddr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_taddr;

    if (warp == 0) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int kt = 0; kt < nk; kt++) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                unsigned char *As = base + stage * GT_STAGE_BYTES, *Bs = As + 16384;
                mbar_arrive_expect_tx(&full_bar[stage], GT_STAGE_BYTES);   // out-of-range parts of a box are zero-filled and still counted
                tma_load_2d(As, &tmA, kt * 64, (int32_t)m0, &full_bar[stage]);
                tma_load_2d(Bs, &tmB, kt * 64, (int32_t)n0, &full_bar[stage]);
                if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        if (lane == 0) {
            constexpr uint32_t idesc = make_idesc_f16(128, 128);
            uint32_t stage = 0, phase = 0;
            for (int kt = 0; kt < nk; kt++) {
                mbar_wait(&full_bar[stage], phase);
                tc_fence_after();
                const uint32_t abase = smem_u32(base + stage * GT_STAGE_BYTES), bbase = abase + 16384;
                const uint64_t ad = make_smem_desc_sw128(abase), bd = make_smem_desc_sw128(bbase);
#pragma unroll
                for (int kk = 0; kk < 4; kk++)   // K = 16 per MMA = 32 bytes inside the 128-byte swizzle atom
                    mma_f16(taddr, ad + (uint64_t)(kk * 2), bd + (uint64_t)(kk * 2), idesc, kt > 0 || kk > 0);
                mma_commit(&empty_bar[stage]);
                if (++stage == GT_STAGES) { stage = 0; phase ^= 1; }
            }
            mma_commit(&acc_full);
        }
    } else {
        // ===== epilogue: warp w may touch TMEM lanes 32 (w % 4) .. + 31 = rows m0 + 32 (w % 4) + lane =====
        mbar_wait(&acc_full, 0);
        tc_fence_after();
        const int q = warp & 3;
        // every MMA has completed, so the operand ring is idle: each warp transposes its 32 x 32 blocks through it
        float *stg = reinterpret_cast<float *>(base) + (warp - 2) * (32 * 33);
        const int64_t row0 = m0 + q * 32;
        const int64_t rmax = min((int64_t)32, g.M - row0);
        for (int c0 = 0; c0 < 128; c0 += 32) {
            uint32_t v[32];
            tmem_ld32(taddr + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; j++) stg[lane * 33 + j] = __uint_as_float(v[j]);
            __syncwarp();
            const int64_t n = n0 + c0 + lane;
            if (n0 + c0 < g.N && rmax > 0) {
                const float bn = (g.bias && n < g.N) ? g.bias[n] : 0.f;
                if (g.C && n < g.N) {
                    float *dst = g.C + row0 * g.ldc + n;
                    for (int rr = 0; rr < rmax; rr++) {
                        const float x = stg[rr * 33 + lane] + bn;
                        dst[(int64_t)rr * g.ldc] = g.relu ? fmaxf(x, 0.f) : x;
                    }
                }
                if (g.C16) {   // two rows per instruction: lanes 0-15 row rr, lanes 16-31 row rr + 1; a lane writes 2 columns
                    const int cp = (lane & 15) * 2, rsel = lane >> 4;
                    const int64_t nn = n0 + c0 + cp;
                    const float b0 = (g.bias && nn < g.N) ? g.bias[nn] : 0.f, b1 = (g.bias && nn + 1 < g.N) ? g.bias[nn + 1] : 0.f;
                    for (int rr = 0; rr < 32; rr += 2) {
                        const int r = rr + rsel;
                        if (r < rmax && nn < g.N) {
                            float x0 = stg[r * 33 + cp] + b0, x1 = stg[r * 33 + cp + 1] + b1;
                            if (g.relu) { x0 = fmaxf(x0, 0.f); x1 = fmaxf(x1, 0.f); }
                            __half *d16 = g.C16 + (row0 + r) * g.ldc16 + nn;
                            if (nn + 1 < g.N) *reinterpret_cast<__half2 *>(d16) = __floats2half2_rn(x0 * g.c16_scale, x1 * g.c16_scale);
                            else *d16 = __float2half_rn(x0 * g.c16_scale);
                        }
                    }
                }
            }
            __syncwarp();
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) { __syncwarp(); tmem_dealloc<128>(taddr); }
}

// C (fp32, may be null) / C16 (fp16 * c16_scale, may be null) = act(A16 [M][K16] . B16 [N][K16]^T + bias)
int launch_gemm_tma(cudaStream_t st, const __half *A16, const __half *B16, int64_t M, int64_t N, int64_t K16, const float *bias,
                    bool relu, float *C, int64_t ldc, __half *C16, int64_t ldc16, float c16_scale)
{
    if (M <= 0 || N <= 0) return PN_OK;
    if (K16 % 8 != 0 || (reinterpret_cast<uintptr_t>(A16) & 15) || (reinterpret_cast<uintptr_t>(B16) & 15) || (C16 && (ldc16 & 1))) {
        set_error("launch_gemm_tma: operands must be 16-byte aligned with K a multiple of 8 (K16 = %lld)", (long long)K16);
        return PN_ERR_INVALID;
    }
    CUtensorMap tmA, tmB;
    int rc = make_tmap_f16(&tmA, A16, M, K16);
    if (rc) return rc;
    rc = make_tmap_f16(&tmB, B16, N, K16);
    if (rc) return rc;
    GemmTmaArgs g{M, N, K16, bias, relu ? 1 : 0, C, ldc, C16, ldc16, c16_scale};
    PN_CUDA(cudaFuncSetAttribute(k_gemm_tma, cudaFuncAttributeMaxDynamicSharedMemorySize, GT_SMEM));
    dim3 grid((unsigned)ceil_div(M, 128), (unsigned)ceil_div(N, 128));
    k_gemm_tma<<<grid, 192, GT_SMEM, st>>>(tmA, tmB, g);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

}  // namespace pn

// X [N][F] fp32 -> fp16 [N][F16], F16 = F rounded up to 8, zero padded: the copy pn_pagg_cfg.x16 points at.
extern "C" size_t pn_pagg_x16_bytes(int32_t N, int32_t F) { return (size_t)N * (size_t)((F + 7) / 8 * 8) * 2; }
extern "C" int pn_pagg_pack_x16(int32_t N, int32_t F, const float *X, void *x16, void *stream)
{
    PN_REQUIRE(N >= 0 && F > 0, "pack_x16: bad sizes N=%d F=%d", N, F);
    if (N == 0) return PN_OK;
    PN_REQUIRE(X && x16 && (reinterpret_cast<uintptr_t>(x16) & 15) == 0, "pack_x16: NULL or misaligned pointer");
    return pn::launch_f32_to_f16_rows((cudaStream_t)stream, X, N, F, F, reinterpret_cast<__half *>(x16), (F + 7) / 8 * 8);
}

// unit-test entry: C = act(A B^T + bias) with fp32 inputs converted to fp16 scratch first (scratch: (M + N) * K16 halves)
extern "C" int pn_debug_gemm_tma(const float *A, const float *B, float *C, void *C16, const float *bias, int64_t M, int64_t N,
                                 int64_t K, int relu, float c16_scale, void *scratch, void *stream)
{
    using namespace pn;
    const int64_t K16 = (K + 7) / 8 * 8;
    __half *a16 = reinterpret_cast<__half *>(scratch), *b16 = a16 + M * K16;
    cudaStream_t st = (cudaStream_t)stream;
    int rc = launch_f32_to_f16_rows(st, A, M, K, K, a16, K16);
    if (rc) return rc;
    rc = launch_f32_to_f16_rows(st, B, N, K, K, b16, K16);
    if (rc) return rc;
    return launch_gemm_tma(st, a16, b16, M, N, K16, bias, relu != 0, C, N, reinterpret_cast<__half *>(C16), N, c16_scale);
}
