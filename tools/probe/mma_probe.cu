// Developer probe (not part of the library): how long does one tcgen05.mma (kind::f16, M = 128) take on this GPU as a
// function of N, of the shared-memory operand layout and of HOW the instruction is issued?
//   STYLE 0: `if (tid == 0)` around the whole issue loop (operands in ordinary registers -> R2UR + election loop per MMA)
//   STYLE 1: the whole warp runs the loop, elect.sync only around the MMA (operands stay in uniform registers)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pathnet_b200/csrc -o tools/probe/mma_probe tools/probe/mma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_util.cuh"
using namespace pn::tc;

__device__ __forceinline__ uint64_t desc_sw128(uint32_t saddr)
{
    uint64_t d = 0;
    d |= (uint64_t)((saddr & 0x3FFFF) >> 4);
    d |= (uint64_t)1 << 16;
    d |= (uint64_t)(1024 >> 4) << 32;
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)2 << 61;
    return d;
}
__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}

constexpr int R = 64;

// one measurement: R MMAs cycling over 8 K-chunks; returns {issue cycles, issue + completion cycles}
template <int STYLE, int N, int SWZ, int A_SBO, int B_SBO, int B_STEP, int DSPLIT>
__device__ __forceinline__ void measure(uint32_t taddr, unsigned char *base, uint64_t *bar, uint32_t &parity, long long *out, bool leader)
{
    constexpr uint32_t idesc = make_idesc_f16(128, N);
    const uint32_t abase = smem_u32(base), bbase = smem_u32(base + 64 * 1024);
    long long t0 = 0, t1 = 0, t2 = 0;
    for (int rep = 0; rep < 3; rep++) {
        t0 = clock64();
#pragma unroll 8
        for (int i = 0; i < R; i++) {
            const int k = i & 7;
            uint64_t ad, bd;
            if (SWZ) { ad = desc_sw128(abase + (k >> 2) * 16384) + (uint64_t)((k & 3) * 2); bd = desc_sw128(bbase + (k >> 2) * 32768) + (uint64_t)((k & 3) * 2); }
            else { ad = make_smem_desc(abase + k * 256, 128, A_SBO); bd = make_smem_desc(bbase + k * B_STEP, 128, B_SBO); }
            if (STYLE == 0 || elect_one()) mma_f16(taddr + (DSPLIT ? (uint32_t)((i & 1) * 256) : 0u), ad, bd, idesc, i >= (DSPLIT ? 2 : 1));
        }
        t1 = clock64();
        if (STYLE == 0 || elect_one()) mma_commit(bar);
        mbar_wait(bar, parity); parity ^= 1;
        tc_fence_after();
        t2 = clock64();
    }
    if (leader) { out[0] = t1 - t0; out[1] = t2 - t0; }
}

constexpr int NCFG = 9;
template <int STYLE>
__global__ void __launch_bounds__(128, 1) k_probe(long long *out, int hammer)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar;
    __shared__ uint32_t s_taddr;
    __shared__ volatile int stop;
    const int tid = threadIdx.x, warp = tid >> 5;
    for (int i = tid; i < 200 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;  // 1.0h
    if (tid == 0) { mbar_init(&bar, 1); fence_mbar_init(); stop = 0; }
    if (warp == 0) tmem_alloc<512>(&s_taddr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_taddr;
    if (STYLE == 0 ? tid == 0 : warp == 0) {
        uint32_t parity = 0;
        long long *o = out + (long long)blockIdx.x * NCFG * 2;
        const bool ld = tid == 0;
        measure<STYLE, 256, 0, 2048, 256, 8192, 0>(taddr, base, &bar, parity, o + 0, ld);    // round-1 forward
        measure<STYLE, 128, 0, 2048, 256, 4096, 0>(taddr, base, &bar, parity, o + 2, ld);    // quarter-pipelined forward
        measure<STYLE, 64, 0, 2048, 256, 2048, 0>(taddr, base, &bar, parity, o + 4, ld);
        measure<STYLE, 128, 0, 8192, 512, 256, 0>(taddr, base, &bar, parity, o + 6, ld);     // backward dgrad
        measure<STYLE, 256, 0, 8192, 512, 256, 0>(taddr, base, &bar, parity, o + 8, ld);
        measure<STYLE, 128, 1, 0, 0, 0, 0>(taddr, base, &bar, parity, o + 10, ld);           // 128-byte swizzle
        measure<STYLE, 256, 1, 0, 0, 0, 0>(taddr, base, &bar, parity, o + 12, ld);
        measure<STYLE, 128, 0, 2048, 256, 4096, 1>(taddr, base, &bar, parity, o + 14, ld);   // alternating accumulators
        measure<STYLE, 256, 0, 2048, 256, 8192, 1>(taddr, base, &bar, parity, o + 16, ld);
        if (tid == 0) stop = 1;
    } else if (hammer && warp >= 1) {
        // other warps stream 16-byte shared-memory loads + stores (a stand-in for the epilogue's traffic)
        uint4 acc = make_uint4(0, 0, 0, 0);
        unsigned char *hb = base + 160 * 1024;
        int it = 0;
        while (!stop) {
            uint4 v;
            const uint32_t la = smem_u32(hb + ((tid * 16 + it * 2048) & 0x7FFF)), sa = smem_u32(hb + ((tid * 16 + it * 2048 + 1024) & 0x7FFF));
            asm volatile("ld.shared.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "r"(la) : "memory");
            acc.x += v.x;
            asm volatile("st.shared.v4.u32 [%0], {%1, %2, %3, %4};" ::"r"(sa), "r"(acc.x), "r"(acc.y), "r"(acc.z), "r"(acc.w) : "memory");
            it++;
        }
        if (acc.x == 0x12345) out[0] = 0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

int main(int argc, char **argv)
{
    const char *names[NCFG] = {"N=256 nosw fwd-r1", "N=128 nosw fwd-q", "N= 64 nosw", "N=128 nosw bwd", "N=256 nosw bwd", "N=128 sw128", "N=256 sw128",
                               "N=128 nosw 2 accumulators", "N=256 nosw 2 accumulators"};
    const int floors[NCFG] = {128, 64, 32, 64, 128, 64, 128, 64, 128};
    long long *out;
    int grid = argc > 1 ? atoi(argv[1]) : 1;
    cudaMalloc(&out, sizeof(long long) * 2 * NCFG * grid);
    cudaFuncSetAttribute(k_probe<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
    cudaFuncSetAttribute(k_probe<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 201 * 1024 + 1024);
    for (int style = 0; style < 2; style++)
        for (int hammer = 0; hammer < 2; hammer++) {
            if (style == 0) k_probe<0><<<grid, 128, 201 * 1024 + 1024>>>(out, hammer);
            else k_probe<1><<<grid, 128, 201 * 1024 + 1024>>>(out, hammer);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); return 1; }
            long long *ho = (long long *)malloc(sizeof(long long) * 2 * NCFG * grid);
            cudaMemcpy(ho, out, sizeof(long long) * 2 * NCFG * grid, cudaMemcpyDeviceToHost);
            printf("== style %d grid %d, other warps: %s\n", style, grid, hammer == 0 ? "idle" : "smem loads + stores");
            for (int c = 0; c < NCFG; c++) {
                long long mx = 0;
                for (int b = 0; b < grid; b++) if (ho[(b * NCFG + c) * 2 + 1] > mx) mx = ho[(b * NCFG + c) * 2 + 1];
                printf("%-28s: issue %6lld cyc, done %6lld cyc (max over CTAs %6lld) = %.1f cyc / MMA (floor %d)\n", names[c], ho[c * 2], ho[c * 2 + 1], mx,
                       (double)ho[c * 2 + 1] / R, floors[c]);
            }
            free(ho);
        }
    return 0;
}
