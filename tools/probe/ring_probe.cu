// Developer probe (not part of the library): the MMA-issue loop of k_lstm_fwd_tc in isolation — a producer warp streams
// 256 KB of weights per step through a shared-memory ring, the MMA warp waits per stage, issues the stage's MMAs
// (M128 N128 K16) and commits the stage's "empty" barrier.  How many cycles per step (64 MMAs, floor 64 x 64 = 4096)?
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -I pathnet_b200/csrc -o tools/probe/ring_probe tools/probe/ring_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>
#include "tc_util.cuh"
using namespace pn::tc;

__device__ __forceinline__ bool elect_one()
{
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
constexpr int PIECE = 4096, NSTEP = 6;

template <int STYLE, int STAGE_KC, int NSTAGE, int FENCE>
__global__ void __launch_bounds__(128, 1) k_ring(const unsigned char *W, long long *out)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *Ws = base + 128 * 1024;
    __shared__ uint64_t full_bar[NSTAGE], empty_bar[NSTAGE], accq[4];
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int STAGE_BYTES = STAGE_KC * PIECE;
    for (int i = tid; i < 128 * 1024 / 4; i += blockDim.x) reinterpret_cast<uint32_t *>(base)[i] = 0x3c003c00u;
    if (tid == 0) {
        for (int s = 0; s < NSTAGE; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int q = 0; q < 4; q++) mbar_init(&accq[q], 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc<512>(&s_taddr);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_taddr;
    if (warp == 1) {
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = 0; t < NSTEP; t++)
                for (int q = 0; q < 4; q++)
                    for (int c = 0; c < 16; c += STAGE_KC) {
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], STAGE_BYTES);
                        bulk_g2s(Ws + stage * STAGE_BYTES, W + (size_t)(q * 16 + c) * PIECE, STAGE_BYTES, &full_bar[stage]);
                        if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                    }
        }
    } else if (warp == 0 && (STYLE == 1 || lane == 0)) {
        constexpr uint32_t idesc = make_idesc_f16(128, 128);
        uint32_t stage = 0, phase = 0;
        const uint32_t xbase = smem_u32(base), hbase = smem_u32(base + 32768);
        for (int t = 0; t < NSTEP; t++) {
            const long long t0 = clock64();
            long long tq0 = 0;
            for (int q = 0; q < 4; q++) {
                for (int c = 0; c < 16; c += STAGE_KC) {
                    mbar_wait(&full_bar[stage], phase);
                    if (FENCE) tc_fence_after();
                    const uint32_t wbase = smem_u32(Ws + stage * STAGE_BYTES);
                    if (STYLE == 0 || elect_one()) {
#pragma unroll
                        for (int i = 0; i < STAGE_KC; i++) {
                            const int kc = c + i;
                            const uint32_t abase = (kc < 8 ? xbase + (uint32_t)kc * 256 : hbase + (uint32_t)(kc - 8) * 256);
                            const uint64_t ad = make_smem_desc(abase, 128, 2048);
                            const uint64_t bd = make_smem_desc(wbase + (uint32_t)i * PIECE, 128, 256);
                            mma_f16(taddr + (uint32_t)q * 128, ad, bd, idesc, kc > 0);
                        }
                        mma_commit(&empty_bar[stage]);
                    }
                    if (STYLE == 1) __syncwarp();
                    if (++stage == NSTAGE) { stage = 0; phase ^= 1; }
                }
                if (STYLE == 0 || elect_one()) mma_commit(&accq[q]);
                if (q == 0) { mbar_wait(&accq[0], t & 1); tq0 = clock64(); }
            }
            mbar_wait(&accq[3], t & 1);
            const long long t1 = clock64();
            if (lane == 0) { out[((long long)blockIdx.x * NSTEP + t) * 2] = tq0 - t0; out[((long long)blockIdx.x * NSTEP + t) * 2 + 1] = t1 - t0; }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

template <int STYLE, int STAGE_KC, int NSTAGE, int FENCE>
void run(const unsigned char *W, long long *out, int grid)
{
    const int smem = 128 * 1024 + NSTAGE * STAGE_KC * PIECE + 1024;
    cudaFuncSetAttribute(k_ring<STYLE, STAGE_KC, NSTAGE, FENCE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    k_ring<STYLE, STAGE_KC, NSTAGE, FENCE><<<grid, 128, smem>>>(W, out);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("error %s\n", cudaGetErrorString(e)); exit(1); }
    long long *ho = (long long *)malloc(sizeof(long long) * 2 * NSTEP * grid);
    cudaMemcpy(ho, out, sizeof(long long) * 2 * NSTEP * grid, cudaMemcpyDeviceToHost);
    printf("style %d stage %2d KB x %2d fence %d grid %3d : per step (quarter 0 / all 64 MMAs) ", STYLE, STAGE_KC * 4, NSTAGE, FENCE, grid);
    for (int t = 0; t < NSTEP; t++) {
        long long a = 0, b = 0;
        for (int g = 0; g < grid; g++) { a += ho[(g * NSTEP + t) * 2]; b += ho[(g * NSTEP + t) * 2 + 1]; }
        printf(" %lld/%lld", a / grid, b / grid);
    }
    printf("\n");
    free(ho);
}

int main(int argc, char **argv)
{
    unsigned char *W; long long *out;
    cudaMalloc(&W, 256 * 1024); cudaMemset(W, 0, 256 * 1024);
    cudaMalloc(&out, sizeof(long long) * 2 * NSTEP * 148);
    for (int grid : {1, 148}) {
        run<0, 2, 11, 1>(W, out, grid);
        run<1, 2, 11, 1>(W, out, grid);
        run<0, 2, 11, 0>(W, out, grid);
        run<1, 2, 11, 0>(W, out, grid);
        run<0, 4, 5, 1>(W, out, grid);
        run<1, 4, 5, 1>(W, out, grid);
        run<1, 4, 5, 0>(W, out, grid);
        run<0, 8, 2, 1>(W, out, grid);
        run<1, 8, 2, 1>(W, out, grid);
        run<1, 1, 22, 1>(W, out, grid);
    }
    return 0;
}
