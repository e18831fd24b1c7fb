// Split-precision tensor-core math mode of PAGG ("tc3", PN_MATH_TC3): the LSTM recurrence on tcgen05 tensor cores at
// fp32-grade accuracy.  Every fp32 operand v is carried as two fp16 terms, hi = fp16(v) and lo = fp16(v - hi)
// (22 significant bits together), and a product is three tensor-core passes into the same fp32 TMEM accumulator,
//     A W  ~=  A_hi W_hi + A_lo W_hi + A_hi W_lo            (the dropped A_lo W_lo term is 2^-22 relative),
// the gates use ex2.approx / rcp.approx (1-2 ulp) instead of tanh.approx (2^-11).  The projections around the LSTM run
// on the strict fp32 kernels of pagg.cu in this mode; the backward is k_lstm_bwd_tc3 below plus three passes of k_wgrad_tc.
//
//   k_lstm_fwd_tc3  same tiling as k_lstm_fwd_tc (CTA = 128 sequences = 128 TMEM lanes, 16 compute warps + MMA warp +
//                   weight-producer warp), but four operand tiles (x_hi, x_lo, h_hi, h_lo) and a 5-stage weight ring
//                   that streams 32 chunks per step (hi and lo image of each 16-deep K-chunk): 96 MMAs (N = 256) per
//                   step instead of 32.  x is single-buffered: the next step's table rows are fetched into registers
//                   while the tensor core works and stored once its MMAs have completed.
//
// Reference semantics: nn.LSTM forward of PathNet_run.py:195 / :265 (fp32); parity bar of this mode = the fp32 mode's.
#include "pagg_tc.cuh"

#include <cstdlib>

namespace pn {
using namespace tc;

constexpr int ACT3_BLOCK_FLOATS = 5 * 32 * 128 * 4;   // saved activations of one (tile, step) as fp32: [5 (i,f,g,o,c)][32 unit4][128 rows][4]
constexpr int TC3_NSTAGE = 5;   // 80 KB ring beside 128 KB of operand tiles
constexpr int TC3_SMEM = 4 * TC_XBYTES + TC3_NSTAGE * TC_CHUNK_BYTES + 2048 /*bias*/ + 128 * TC_MAX_L * 4 /*row offsets*/ + 1024 /*align*/;
static_assert(TC3_SMEM <= 232448, "k_lstm_fwd_tc3: shared memory over the 227 KB limit");

__device__ __forceinline__ void split_f16(float v, __half &hi, __half &lo)
{
    hi = __float2half_rn(v);
    lo = __float2half_rn(v - __half2float(hi));
}

// forward weights: chunk 2c = hi image, chunk 2c + 1 = lo image of K-chunk c (same [512 cols][16 k] canonical layout as
// k_pack_lstm_tc; no pre-halving of the sigmoid gates: this mode evaluates the sigmoid itself)
__global__ void k_pack_lstm_tc3(const float *__restrict__ w_ih, const float *__restrict__ w_hh, const float *__restrict__ b_ih,
                                const float *__restrict__ b_hh, unsigned char *__restrict__ Wh, float *__restrict__ bp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < TC_G4 * TC_K2) {
        const int col = i / TC_K2, k = i % TC_K2, unit = col >> 2, gate = col & 3;
        const float v = k < TC_H ? w_ih[(gate * TC_H + unit) * TC_H + k] : w_hh[(gate * TC_H + unit) * TC_H + (k - TC_H)];
        const int chunk = k / TC_KC, kk = k % TC_KC;
        __half hi, lo;
        split_f16(v, hi, lo);
        *reinterpret_cast<__half *>(Wh + (size_t)(2 * chunk) * TC_CHUNK_BYTES + canon_off(col, kk, TC_KC)) = hi;
        *reinterpret_cast<__half *>(Wh + (size_t)(2 * chunk + 1) * TC_CHUNK_BYTES + canon_off(col, kk, TC_KC)) = lo;
    }
    if (i < TC_G4) {
        const int unit = i >> 2, gate = i & 3;
        bp[i] = b_ih[gate * TC_H + unit] + b_hh[gate * TC_H + unit];
    }
}

// backward weights W' = [W_ih | W_hh] as the dgrad B operand ([256 rows = input index][32 gate columns] per chunk, see
// k_pack_lstm_tc_bwd): chunk 2c = hi, 2c + 1 = lo
__global__ void k_pack_lstm_tc3_bwd(const float *__restrict__ w_ih, const float *__restrict__ w_hh, unsigned char *__restrict__ WhT)
{
    constexpr int KC = 32, CHUNK_BYTES = TC_K2 * KC * 2;
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= TC_G4 * TC_K2) return;
    const int col = i / TC_K2, k = i % TC_K2, unit = col >> 2, gate = col & 3;
    const float v = k < TC_H ? w_ih[(gate * TC_H + unit) * TC_H + k] : w_hh[(gate * TC_H + unit) * TC_H + (k - TC_H)];
    const int chunk = col / KC, kk = col % KC;
    __half hi, lo;
    split_f16(v, hi, lo);
    *reinterpret_cast<__half *>(WhT + (size_t)(2 * chunk) * CHUNK_BYTES + canon_off(k, kk, KC)) = hi;
    *reinterpret_cast<__half *>(WhT + (size_t)(2 * chunk + 1) * CHUNK_BYTES + canon_off(k, kk, KC)) = lo;
}

// the table the LSTM kernel gathers: T * 1/(1-p) as two fp16 terms (8 values per thread, 16-byte stores)
__global__ void k_split_table(const float *__restrict__ T, float scale, __half *__restrict__ hi, __half *__restrict__ lo, int64_t n8)
{
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n8) return;
    const float4 a = __ldg(reinterpret_cast<const float4 *>(T) + 2 * i), b = __ldg(reinterpret_cast<const float4 *>(T) + 2 * i + 1);
    const float v[8] = {a.x * scale, a.y * scale, a.z * scale, a.w * scale, b.x * scale, b.y * scale, b.z * scale, b.w * scale};
    float h[8], l[8];
#pragma unroll
    for (int j = 0; j < 8; j++) { h[j] = __half2float(__float2half_rn(v[j])); l[j] = v[j] - h[j]; }
    reinterpret_cast<uint4 *>(hi)[i] = pack8_f16(h);
    reinterpret_cast<uint4 *>(lo)[i] = pack8_f16(l);
}

// 20 warps like the backward kernels: 16 compute warps + a service warpgroup (MMA issuer, weight producer, two idle warps) so that
// setmaxnreg can move registers (96 per thread at launch -> service warps 32, compute warps 112: no spills around the 32 gather
// registers that stay live across the wait for the accumulators)
__global__ void __launch_bounds__(TCB_THREADS, 1) k_lstm_fwd_tc3(LstmTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 1023) & ~(uintptr_t)1023);
    unsigned char *Xhi = base, *Xlo = base + TC_XBYTES;                   // x_t   [128][128] fp16 canonical, two terms
    unsigned char *Hhi = base + 2 * TC_XBYTES, *Hlo = base + 3 * TC_XBYTES;   // h_{t-1}
    unsigned char *Ws = base + 4 * TC_XBYTES;                             // ring of weight chunks
    float *bias_s = reinterpret_cast<float *>(Ws + TC3_NSTAGE * TC_CHUNK_BYTES);
    int32_t *rowoff_s = reinterpret_cast<int32_t *>(bias_s + TC_G4);      // [128][TC_MAX_L]: table row of (sequence, step), -1 = padding
    __shared__ uint64_t full_bar[TC3_NSTAGE], empty_bar[TC3_NSTAGE], a_ready, acc_full;
    __shared__ uint32_t s_taddr;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int64_t p0 = (int64_t)blockIdx.x * 128;
    const int L = a.L;

    if (tid == 0) {
        for (int s = 0; s < TC3_NSTAGE; s++) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        mbar_init(&a_ready, TC_CWARPS * 32);
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == TC_CWARPS) tmem_alloc<512>(&s_taddr);
    for (int i = tid; i < TC_G4; i += TCB_THREADS) bias_s[i] = a.bp[i];
    for (int i = tid; i < 128 * L; i += TCB_THREADS) {
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

    if (warp > TC_CWARPS + 1) {
        reg_dec<32>();   // (idle warps of the service warpgroup)
    } else if (warp == TC_CWARPS + 1) {
        reg_dec<32>();
        // ===== weight producer: per step and K-chunk the hi image, then the lo image (step 0: x half only, h_0 = 0) =====
        if (lane == 0) {
            uint32_t stage = 0, phase = 0;
            for (int t = 0; t < L; t++) {
                const int nch = 2 * (t == 0 ? TC_NCHUNK / 2 : TC_NCHUNK);
                for (int c = 0; c < nch; c++) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    mbar_arrive_expect_tx(&full_bar[stage], TC_CHUNK_BYTES);
                    bulk_g2s(Ws + stage * TC_CHUNK_BYTES, a.Wh + (size_t)c * TC_CHUNK_BYTES, TC_CHUNK_BYTES, &full_bar[stage]);
                    if (++stage == TC3_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == TC_CWARPS) {
        reg_dec<32>();
        // ===== MMA issuer (whole warp converged, one elected lane issues; see k_lstm_fwd_tc) =====
        constexpr uint32_t idesc = make_idesc_f16(128, 256);
        uint32_t stage = 0, phase = 0;
        const bool trm = a.trace && blockIdx.x == a.trace_block && lane == 0;
        for (int t = 0; t < L; t++) {
            mbar_wait(&a_ready, t & 1);  // x_t (and h_{t-1}) staged, TMEM drained
            tc_fence_after();
            if (trm) a.trace[t * 8 + 4] = clock64();
            const int nch = t == 0 ? TC_NCHUNK / 2 : TC_NCHUNK;
            for (int c = 0; c < nch; c++) {
                if (trm && c == nch - 1) a.trace[t * 8 + 5] = clock64();
                // chunk c = 16 consecutive k: two core matrices (256 B) of the x half (c < 8) or the h half
                const uint32_t off = (uint32_t)(c < TC_NCHUNK / 2 ? c : c - TC_NCHUNK / 2) * 256;
                const uint32_t a_hi = smem_u32(c < TC_NCHUNK / 2 ? Xhi : Hhi) + off, a_lo = smem_u32(c < TC_NCHUNK / 2 ? Xlo : Hlo) + off;
#pragma unroll
                for (int part = 0; part < 2; part++) {   // part 0: W_hi against A_hi and A_lo; part 1: W_lo against A_hi
                    mbar_wait(&full_bar[stage], phase);
                    tc_fence_after();
                    const uint32_t wbase = smem_u32(Ws + stage * TC_CHUNK_BYTES);
                    if (elect_one()) {
                        const uint64_t ad_hi = make_smem_desc(a_hi, 128, (TC_H / 8) * 128), ad_lo = make_smem_desc(a_lo, 128, (TC_H / 8) * 128);
#pragma unroll
                        for (int nh = 0; nh < 2; nh++) {
                            const uint64_t bd = make_smem_desc(wbase + (uint32_t)nh * (256 / 8) * ((TC_KC / 8) * 128), 128, (TC_KC / 8) * 128);
                            if (part == 0) {
                                mma_f16(taddr + (uint32_t)nh * 256, ad_hi, bd, idesc, c > 0);
                                mma_f16(taddr + (uint32_t)nh * 256, ad_lo, bd, idesc, true);
                            } else {
                                mma_f16(taddr + (uint32_t)nh * 256, ad_hi, bd, idesc, true);
                            }
                        }
                        mma_commit(&empty_bar[stage]);
                        if (part == 1 && c == nch - 1) mma_commit(&acc_full);
                    }
                    __syncwarp();
                    if (++stage == TC3_NSTAGE) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== gather + gate math: warp = (TMEM lane quarter lq, column quarter cq); the thread owns sequence
        //       p0 + 32 lq + lane (= its TMEM lane) and hidden units [32 cq, 32 cq + 32) =====
        reg_inc<112>();
        const int lq = warp & 3, cq = warp >> 2;
        const int r = lq * 32 + lane;
        const int64_t p = p0 + r;
        const bool valid = p < a.P;
        float c[32];
#pragma unroll
        for (int j = 0; j < 32; j++) c[j] = 0.f;
        const int64_t ntiles = gridDim.x;
        const uint32_t key = dropout_key(a.seed_dev ? a.drop_seed + *a.seed_dev : a.drop_seed, 0);
        // this thread's quarter (32 values, two terms) of the table row of step t_: loads + dropout mask into registers ...
        uint4 gh[4], gl[4];
        uint32_t gmask = 0;
        auto gather_load = [&](int t_) {
            const int32_t row_off = rowoff_s[r * TC_MAX_L + t_];
            const bool ok = row_off >= 0;
            const int64_t eoff = (int64_t)(ok ? row_off : 0) * TC_H + cq * 32;
            const uint4 *sh = reinterpret_cast<const uint4 *>(a.T16 + eoff), *sl = reinterpret_cast<const uint4 *>(a.T16lo + eoff);
#pragma unroll
            for (int i = 0; i < 4; i++) {
                gh[i] = ok ? __ldg(sh + i) : make_uint4(0, 0, 0, 0);
                gl[i] = ok ? __ldg(sl + i) : make_uint4(0, 0, 0, 0);
            }
            gmask = 0;
            if (ok && a.drop_thr) {
                const uint64_t elem0 = (uint64_t)((p + a.p_base) * L + t_) * TC_H;
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    const uint64_t pair = (elem0 + (cq * 4 + kk) * 8) >> 1;
                    uint32_t *wh = reinterpret_cast<uint32_t *>(&gh[kk]), *wl = reinterpret_cast<uint32_t *>(&gl[kk]);
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const uint32_t h = dropout_hash_pair_k(key, pair + q);
                        const uint32_t k0 = (h & 0xFFFFu) < a.drop_thr ? 0u : 1u, k1 = (h >> 16) < a.drop_thr ? 0u : 1u;
                        const uint32_t m = (0u - k0) & 0x0000FFFFu | (0u - k1) & 0xFFFF0000u;
                        wh[q] &= m;
                        wl[q] &= m;
                        gmask |= (k0 | (k1 << 1)) << (kk * 8 + q * 2);
                    }
                }
            }
        };
        // ... and from there into the operand tiles (+ the x half of the wgrad operand image and the mask word)
        auto gather_store = [&](int t_) {
            unsigned char *xh_blk = a.xh_img + ((int64_t)t_ * ntiles + blockIdx.x) * XH_BLOCK_BYTES;
#pragma unroll
            for (int kk = 0; kk < 4; kk++) {
                const int k8 = cq * 4 + kk;
                *reinterpret_cast<uint4 *>(Xhi + canon_off(r, k8 * 8, TC_H)) = gh[kk];
                *reinterpret_cast<uint4 *>(Xlo + canon_off(r, k8 * 8, TC_H)) = gl[kk];
                if (a.save) {
                    *reinterpret_cast<uint4 *>(xh_blk + (size_t)k8 * 2048 + r * 16) = gh[kk];
                    *reinterpret_cast<uint4 *>(a.xh_img_lo + (xh_blk - a.xh_img) + (size_t)k8 * 2048 + r * 16) = gl[kk];
                }
            }
            if (a.dmask && a.save && a.drop_thr) a.dmask[(((int64_t)blockIdx.x * L + t_) * 4 + cq) * 128 + r] = gmask;
        };
        gather_load(0);
        gather_store(0);
        // h_{-1} = 0: h half of step 0's image
#pragma unroll
        for (int k8 = 16 + cq * 4; k8 < 16 + cq * 4 + 4; k8++)
            if (a.save) {
                *reinterpret_cast<uint4 *>(a.xh_img + (int64_t)blockIdx.x * XH_BLOCK_BYTES + (size_t)k8 * 2048 + r * 16) = make_uint4(0, 0, 0, 0);
                *reinterpret_cast<uint4 *>(a.xh_img_lo + (int64_t)blockIdx.x * XH_BLOCK_BYTES + (size_t)k8 * 2048 + r * 16) = make_uint4(0, 0, 0, 0);
            }
        const bool tr = a.trace && blockIdx.x == a.trace_block && tid == 0;
        for (int t = 0; t < L; t++) {
            if (tr) a.trace[t * 8 + 0] = clock64();
            fence_proxy_async();   // generic-proxy smem writes (x_t, h_{t-1}) -> async proxy (tcgen05.mma)
            tc_fence_before();     // order this thread's earlier tcgen05.ld before the MMA that overwrites TMEM
            mbar_arrive(&a_ready);
            if (t + 1 < L) gather_load(t + 1);   // in flight while the tensor core works on step t
            if (tr) a.trace[t * 8 + 1] = clock64();
            mbar_wait(&acc_full, t & 1);
            tc_fence_after();
            if (tr) a.trace[t * 8 + 2] = clock64();
            if (t + 1 < L) gather_store(t + 1);  // every MMA of step t has completed: the x tiles are free
            if (tr) a.trace[t * 8 + 6] = clock64();
            float *act = a.acts32 + ((int64_t)blockIdx.x * L + t) * ACT3_BLOCK_FLOATS + r * 4;   // fp32: [5][32 unit4][128 rows][4]
            const bool t_is_last = t == L - 1;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int c0 = cq * 128 + jj * 32;   // packed gate columns of units [c0/4, c0/4 + 8)
                uint32_t ph[4], pl[4];
                float hq[4], fi[4], ff[4], fg[4], fo[4], fc[4];
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
                        const int u = 2 * u2 + e, unit = jj * 8 + u;
                        const int vi = 4 * (u & 3);
                        const float4 b4 = *reinterpret_cast<const float4 *>(bias_s + c0 + 4 * u);
                        // sigmoid(x) = 1 / (1 + e^-x), tanh(x) = (1 - e^-2x) / (1 + e^-2x): five ex2 and two rcp per unit -- the
                        // reciprocals of i, f, g (and of o, tanh c) share one rcp of the product of their denominators
                        const float di = 1.f + exp_neg(__uint_as_float(v[vi + 0]) + b4.x);
                        const float df = 1.f + exp_neg(__uint_as_float(v[vi + 1]) + b4.y);
                        const float eg = exp_neg2(__uint_as_float(v[vi + 2]) + b4.z);
                        const float dg = 1.f + eg;
                        const float dov = 1.f + exp_neg(__uint_as_float(v[vi + 3]) + b4.w);
                        const float r3 = rcp_approx(di * df * dg);
                        gi2[e] = r3 * (df * dg);
                        gf2[e] = r3 * (di * dg);
                        gg2[e] = (1.f - eg) * (r3 * (di * df));
                        const float cn = fmaf(gf2[e], c[unit], gi2[e] * gg2[e]);
                        c[unit] = cn;
                        cn2[e] = cn;
                        const float ec = exp_neg2(cn), dc = 1.f + ec;
                        const float r2 = rcp_approx(dov * dc);
                        go2[e] = r2 * dc;
                        hh2[e] = go2[e] * ((1.f - ec) * (r2 * dov));
                    }
                    __half2 t2;
#pragma unroll
                    for (int e = 0; e < 2; e++) {
                        const int q4 = (u2 & 1) * 2 + e;
                        fi[q4] = gi2[e]; ff[q4] = gf2[e]; fg[q4] = gg2[e]; fo[q4] = go2[e]; fc[q4] = cn2[e];
                    }
                    t2 = __floats2half2_rn(hh2[0], hh2[1]); ph[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    const float2 back = __half22float2(t2);
                    t2 = __floats2half2_rn(hh2[0] - back.x, hh2[1] - back.y); pl[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    hq[(u2 & 1) * 2] = hh2[0]; hq[(u2 & 1) * 2 + 1] = hh2[1];
                    if ((u2 & 1) && valid && a.save) {   // four units complete: their activations leave as one float4 per tensor
                        float *ap = act + (c0 / 16 + (u2 >> 1)) * 512;
                        *reinterpret_cast<float4 *>(ap + 0 * 16384) = make_float4(fi[0], fi[1], fi[2], fi[3]);
                        *reinterpret_cast<float4 *>(ap + 1 * 16384) = make_float4(ff[0], ff[1], ff[2], ff[3]);
                        *reinterpret_cast<float4 *>(ap + 2 * 16384) = make_float4(fg[0], fg[1], fg[2], fg[3]);
                        *reinterpret_cast<float4 *>(ap + 3 * 16384) = make_float4(fo[0], fo[1], fo[2], fo[3]);
                        *reinterpret_cast<float4 *>(ap + 4 * 16384) = make_float4(fc[0], fc[1], fc[2], fc[3]);
                    }
                    if (t_is_last && (u2 & 1)) {
                        // final h (fp32) -> staging tile in the (now idle) x tiles, 16-byte slots XOR-swizzled by the row
                        const int j4 = c0 / 16 + (u2 >> 1);
                        *reinterpret_cast<float4 *>(Xhi + r * 512 + ((j4 ^ (r & 7)) << 4)) = make_float4(hq[0], hq[1], hq[2], hq[3]);
                    }
                }
                if (!t_is_last) {   // h_t -> operand tiles of the next step (two terms), and the h half of its wgrad image (hi term)
                    const uint4 hpk = make_uint4(ph[0], ph[1], ph[2], ph[3]);
                    *reinterpret_cast<uint4 *>(Hhi + canon_off(r, c0 / 4, TC_H)) = hpk;
                    *reinterpret_cast<uint4 *>(Hlo + canon_off(r, c0 / 4, TC_H)) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    if (a.save) {
                        const size_t ho = ((int64_t)(t + 1) * ntiles + blockIdx.x) * XH_BLOCK_BYTES + (size_t)(16 + c0 / 32) * 2048 + r * 16;
                        *reinterpret_cast<uint4 *>(a.xh_img + ho) = hpk;
                        *reinterpret_cast<uint4 *>(a.xh_img_lo + ho) = make_uint4(pl[0], pl[1], pl[2], pl[3]);
                    }
                }
            }
            if (tr) a.trace[t * 8 + 3] = clock64();
        }
        // ---- final h: staging tile -> hfin[P][H], whole 512-byte rows per warp ----
        asm volatile("bar.sync 1, %0;" ::"n"(TC_CWARPS * 32) : "memory");
        for (int rr = warp * 8; rr < warp * 8 + 8; rr++) {
            const int64_t pr = p0 + rr;
            if (pr < a.P) {
                const float4 hv = *reinterpret_cast<const float4 *>(Xhi + rr * 512 + ((lane ^ (rr & 7)) << 4));
                *reinterpret_cast<float4 *>(a.hfin + pr * TC_H + lane * 4) = hv;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CWARPS) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

// ---------------------------------------------------------------------------
// k_lstm_bwd_tc3: BPTT at fp32-grade accuracy.  Same tiling and TMEM map as k_lstm_bwd_tc ([0,128) dx, [128,384) dh ping-pong,
// [384,512) running dc; 16 compute warps, MMA warp, W' producer, setmaxnreg), but
//   * the saved activations are fp32 (written by k_lstm_fwd_tc3), tanh(c_t) goes through ex2 / rcp;
//   * the gate gradients da leave as TWO fp16 terms, into the dgrad operand and into the wgrad operand images;
//   * [dx | dh] = da W' is three passes per K-chunk: da_hi W'_hi + da_lo W'_hi + da_hi W'_lo (192 N=128 MMAs per step);
//   * the 128 KB operand region holds two K-QUARTER buffers (hi 32 KB + lo 32 KB each) instead of the whole tile: quarter
//     Q = 4 s + i goes to buffer Q & 1 once the MMAs of quarter Q - 2 have committed (q_free).
// ---------------------------------------------------------------------------
constexpr int TCB3_QBYTES = 128 * 128 * 2;   // one term of one K-quarter of da: [128 rows][128 gate columns] fp16 canonical

// PF bit 1: activations fetched into registers one half-iteration ahead; bit 0: prefetch.global.L2 of the next step's activations
// a whole step ahead.  Same-box A/B (pubmed step, PN_BWD3_PF): 0 -> 5.28 / 5.32 ms, 1 -> 5.67 / 5.65, 3 -> 5.60 / 5.59: the L2
// prefetch costs 0.35 ms here (this kernel runs at 0.7 of the HBM roof; the hints add traffic ahead of demand); second box,
// 0 vs 2: 5.26 / 5.17 vs 5.28 / 5.38 ms -- the register prefetch does not pay either (loads of 4 warps per scheduler already
// overlap), so the default is 0.
template <int PF>
__global__ void __launch_bounds__(TCB_THREADS, 1) k_lstm_bwd_tc3(LstmBwdTcArgs a)
{
    extern __shared__ unsigned char smem_dyn[];
    unsigned char *base = reinterpret_cast<unsigned char *>((reinterpret_cast<uintptr_t>(smem_dyn) + 127) & ~(uintptr_t)127);
    unsigned char *As = base;                 // buffer b at As + b * 2 * TCB3_QBYTES: hi term, then lo term
    unsigned char *Ws = base + TCB_A_BYTES;   // ring of W' chunks
    int32_t *rowoff_s = reinterpret_cast<int32_t *>(Ws + TCB_NSTAGE * TCB_CHUNK_BYTES);   // [128][L]: dT row of (sequence, step)
    __shared__ uint64_t full_bar[TCB_NSTAGE], empty_bar[TCB_NSTAGE], a_ready[4], q_free[2], acc_full;
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
        mbar_init(&q_free[0], 1); mbar_init(&q_free[1], 1);
        mbar_init(&acc_full, 1);
        fence_mbar_init();
    }
    if (warp == TC_CWARPS) tmem_alloc<512>(&s_taddr);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t taddr = s_taddr;

    if (warp >= TC_CWARPS) {
        reg_dec<32>();
        if (warp == TC_CWARPS + 1) {
            // ===== W' producer: quarter-major chunk order, hi image then lo image of each chunk =====
            if (lane == 0) {
                uint32_t stage = 0, phase = 0;
                for (int s = 0; s < L; s++)
                    for (int j = 0; j < 2 * TCB_NCHUNK; j++) {
                        const int jj = j >> 1, c = 4 * (jj & 3) + (jj >> 2);
                        mbar_wait(&empty_bar[stage], phase ^ 1);
                        mbar_arrive_expect_tx(&full_bar[stage], TCB_CHUNK_BYTES);
                        bulk_g2s(Ws + stage * TCB_CHUNK_BYTES, a.WhT + (size_t)(2 * c + (j & 1)) * TCB_CHUNK_BYTES, TCB_CHUNK_BYTES, &full_bar[stage]);
                        if (++stage == TCB_NSTAGE) { stage = 0; phase ^= 1; }
                    }
            }
        } else if (warp == TC_CWARPS) {
            // ===== MMA issuer (converged warp, one elected lane) =====
            uint32_t stage = 0, phase = 0;
            constexpr uint32_t idesc = make_idesc_f16(128, 128);
            for (int s = 0; s < L; s++) {
                const int t = L - 1 - s;
                const uint32_t dh_col = 128u + (uint32_t)(s & 1) * 128u;
                for (int j = 0; j < TCB_NCHUNK; j++) {
                    const int i = j >> 2, cqj = j & 3;   // K-quarter i, chunk 4 cqj + i of W' = 32 gate columns of column quarter cqj
                    if (cqj == 0) { mbar_wait(&a_ready[i], s & 1); tc_fence_after(); }
                    const int Q = 4 * s + i;
                    const uint32_t a_hi = smem_u32(As) + (uint32_t)(Q & 1) * (2 * TCB3_QBYTES) + (uint32_t)cqj * 4 * 128, a_lo = a_hi + TCB3_QBYTES;
#pragma unroll
                    for (int part = 0; part < 2; part++) {   // part 0: W'_hi against da_hi and da_lo; part 1: W'_lo against da_hi
                        mbar_wait(&full_bar[stage], phase);
                        tc_fence_after();
                        const uint32_t wbase = smem_u32(Ws + stage * TCB_CHUNK_BYTES);
                        if (elect_one()) {
#pragma unroll
                            for (int ks = 0; ks < TCB_KC / 16; ks++) {
                                const uint64_t ad_hi = make_smem_desc(a_hi + (uint32_t)ks * 256, 128, (128 / 8) * 128);
                                const uint64_t ad_lo = make_smem_desc(a_lo + (uint32_t)ks * 256, 128, (128 / 8) * 128);
                                // rows 0..127 of W' give dx, rows 128..255 (16 row groups further) give dh_{t-1} (unused at t = 0)
                                const uint64_t bd_x = make_smem_desc(wbase + (uint32_t)ks * 256, 128, (TCB_KC / 8) * 128);
                                const uint64_t bd_h = make_smem_desc(wbase + 16u * ((TCB_KC / 8) * 128) + (uint32_t)ks * 256, 128, (TCB_KC / 8) * 128);
                                const bool acc = (j > 0) || (ks > 0) || (part > 0);
                                mma_f16(taddr, ad_hi, bd_x, idesc, acc);
                                if (part == 0) mma_f16(taddr, ad_lo, bd_x, idesc, true);
                                if (t > 0) {
                                    mma_f16(taddr + dh_col, ad_hi, bd_h, idesc, acc);
                                    if (part == 0) mma_f16(taddr + dh_col, ad_lo, bd_h, idesc, true);
                                }
                            }
                            mma_commit(&empty_bar[stage]);
                            if (part == 1 && cqj == 3) mma_commit(&q_free[Q & 1]);
                            if (part == 1 && j == TCB_NCHUNK - 1) mma_commit(&acc_full);
                        }
                        __syncwarp();
                        if (++stage == TCB_NSTAGE) { stage = 0; phase ^= 1; }
                    }
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
        // dx of step `ss` (time tt = L-1-ss) -> dT with the forward's dropout mask (same scatter as k_lstm_bwd_tc)
        auto scatter_dx = [&](int ss) {
            const int tt = L - 1 - ss;
            uint32_t mbits = 0xFFFFFFFFu;
            const bool have_mask = a.dmask != nullptr && a.drop_thr != 0;
            if (have_mask) mbits = __ldg(a.dmask + (((int64_t)tile * L + tt) * 4 + cq) * 128 + r);
            uint32_t v[32];
            tmem_ld32(lane_base + cq * 32, v);   // warp-collective: every lane loads, only valid rows scatter
            tmem_ld_wait();
            if (valid) {
                const int64_t e = p * L + tt;
                float *dst = a.dT + (int64_t)rowoff_s[r * L + tt] * TC_H + cq * 32;
                const float mul = inv_scale * (a.drop_thr ? a.drop_scale : 1.f);
#pragma unroll
                for (int c8 = 0; c8 < 4; c8++) {
                    if (a.drop_thr && !have_mask) {
                        const uint64_t pair = ((uint64_t)(e + a.p_base * L) * TC_H + cq * 32 + c8 * 8) >> 1;
                        uint32_t b = 0;
#pragma unroll
                        for (int q = 0; q < 4; q++) {
                            const uint32_t h = dropout_hash_pair_k(dkey, pair + q);
                            b |= (((h & 0xFFFFu) < a.drop_thr ? 0u : 1u) | ((h >> 16) < a.drop_thr ? 0u : 2u)) << (2 * q);
                        }
                        mbits = (mbits & ~(0xFFu << (c8 * 8))) | (b << (c8 * 8));
                    }
#pragma unroll
                    for (int q = 0; q < 2; q++) {
                        const uint32_t b = a.drop_thr ? (mbits >> (c8 * 8 + 4 * q)) & 15u : 15u;
                        if (b) atomicAdd(reinterpret_cast<float4 *>(dst + c8 * 8 + 4 * q),
                                         make_float4((b & 1u) ? __uint_as_float(v[c8 * 8 + 4 * q]) * mul : 0.f,
                                                     (b & 2u) ? __uint_as_float(v[c8 * 8 + 4 * q + 1]) * mul : 0.f,
                                                     (b & 4u) ? __uint_as_float(v[c8 * 8 + 4 * q + 2]) * mul : 0.f,
                                                     (b & 8u) ? __uint_as_float(v[c8 * 8 + 4 * q + 3]) * mul : 0.f));
                    }
                }
            }
            __syncwarp();
        };
        // factored dh_final: this sequence's two scalars and its output row
        float dh_kk = 1.f, dh_ds = 0.f;
        int64_t grow = 0;
        if (!a.dhfin && valid) { dh_kk = a.dh_k[p]; dh_ds = a.dh_s[p]; grow = p / a.nw; }
        // the saved activations of a half-iteration (4 units x {i, f, g, o, c_t, c_{t-1}}: six 16-byte loads, plus dh_final at the
        // first step) are fetched one half-iteration ahead of the math that consumes them, across quarter and step boundaries
        struct Raw3 { float4 gi, gf, gg, go, cn, cp, d0; };
        auto load_raw = [&](int s_, int hi_, Raw3 &x) {
            const int t_ = L - 1 - s_, u4_ = cq * 8 + hi_;
            x.gi = x.gf = x.gg = x.go = x.cn = x.cp = x.d0 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (valid) {
                const float *ap = a.acts32 + ((int64_t)tile * L + t_) * ACT3_BLOCK_FLOATS + r * 4 + u4_ * 512;
                x.gi = __ldg(reinterpret_cast<const float4 *>(ap + 0 * 16384));
                x.gf = __ldg(reinterpret_cast<const float4 *>(ap + 1 * 16384));
                x.gg = __ldg(reinterpret_cast<const float4 *>(ap + 2 * 16384));
                x.go = __ldg(reinterpret_cast<const float4 *>(ap + 3 * 16384));
                x.cn = __ldg(reinterpret_cast<const float4 *>(ap + 4 * 16384));
                if (t_ > 0) x.cp = __ldg(reinterpret_cast<const float4 *>(ap - ACT3_BLOCK_FLOATS + 4 * 16384));
                if (s_ == 0) x.d0 = a.dhfin ? __ldg(reinterpret_cast<const float4 *>(a.dhfin + p * TC_H + u4_ * 4))
                                            : __ldg(reinterpret_cast<const float4 *>(a.gdrop + grow * a.gdrop_ld + u4_ * 4));
            }
        };
        Raw3 nxt;
        if (PF & 2) load_raw(0, 0, nxt);
        for (int s = 0; s < L; s++) {
            const int t = L - 1 - s;
            const int64_t blk = (int64_t)t * a.ntiles + tile;
            unsigned char *da_blk = a.da_img + blk * DA_BLOCK_BYTES, *da_blk_lo = a.da_img_lo + blk * DA_BLOCK_BYTES;
            // ---- pull the next step's activations towards L2 (a whole step ahead; one lane per 128-byte line) ----
            if ((PF & 1) && valid && t > 0 && (lane & 7) == 0) {
                const float *actp = a.acts32 + ((int64_t)tile * L + t - 1) * ACT3_BLOCK_FLOATS + r * 4;
#pragma unroll
                for (int u = cq * 8; u < cq * 8 + 8; u++) {
#pragma unroll
                    for (int q = 0; q < 5; q++) prefetch_l2(actp + q * 16384 + u * 512);
                    if (t > 1) prefetch_l2(actp - ACT3_BLOCK_FLOATS + 4 * 16384 + u * 512);
                }
            }
            // ---- the previous step's dx (complete since acc_full) -> dT, before this step's MMAs overwrite the region ----
            if (s > 0) scatter_dx(s - 1);
            const uint32_t dh_rd = 128u + (uint32_t)((s - 1) & 1) * 128u;
#pragma unroll 1
            for (int hi = 0; hi < 8; hi++) {   // half-iterations of 4 units; two of them fill one K-quarter of this warp's columns
                const int i = hi >> 1;
                const int u4 = cq * 8 + hi, u0 = u4 * 4;   // units [u0, u0 + 4)
                Raw3 cur;
                if (PF & 2) {
                    cur = nxt;
                    if (hi < 7) load_raw(s, hi + 1, nxt);
                    else if (s + 1 < L) load_raw(s + 1, 0, nxt);
                } else {
                    load_raw(s, hi, cur);
                }
                const float4 gi = cur.gi, gf = cur.gf, gg = cur.gg, go = cur.go, cn = cur.cn, cp = cur.cp;
                float4 d0 = cur.d0;
                float dh[4], dc[4];
                if (s == 0) {
                    if (!a.dhfin) {
                        const float4 wh = __ldg(reinterpret_cast<const float4 *>(a.attw_h + u0));
                        d0 = make_float4(fmaf(dh_kk, d0.x, dh_ds * wh.x), fmaf(dh_kk, d0.y, dh_ds * wh.y), fmaf(dh_kk, d0.z, dh_ds * wh.z),
                                         fmaf(dh_kk, d0.w, dh_ds * wh.w));
                        if (!valid) d0 = make_float4(0.f, 0.f, 0.f, 0.f);
                    }
                    dh[0] = d0.x * scale; dh[1] = d0.y * scale; dh[2] = d0.z * scale; dh[3] = d0.w * scale;
                    dc[0] = dc[1] = dc[2] = dc[3] = 0.f;
                } else {
                    uint32_t v0[4], v1[4];
                    tmem_ld4(lane_base + dh_rd + u0, v0);
                    tmem_ld4(lane_base + 384 + u0, v1);
                    tmem_ld_wait();
#pragma unroll
                    for (int j = 0; j < 4; j++) { dh[j] = __uint_as_float(v0[j]); dc[j] = __uint_as_float(v1[j]); }
                }
                const float gi_[4] = {gi.x, gi.y, gi.z, gi.w}, gf_[4] = {gf.x, gf.y, gf.z, gf.w}, gg_[4] = {gg.x, gg.y, gg.z, gg.w};
                const float go_[4] = {go.x, go.y, go.z, go.w}, cn_[4] = {cn.x, cn.y, cn.z, cn.w}, cp_[4] = {cp.x, cp.y, cp.z, cp.w};
                float da[2][8];  // [q][.]: packed columns (u0 + 2q + {0,1}) * 4 + gate
                uint32_t dcn_bits[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const float tcn = tanh_acc(cn_[j]);
                    const float dcn = fmaf(dh[j] * go_[j], 1.f - tcn * tcn, dc[j]);
                    da[j >> 1][(j & 1) * 4 + 0] = dcn * gg_[j] * gi_[j] * (1.f - gi_[j]);
                    da[j >> 1][(j & 1) * 4 + 1] = dcn * cp_[j] * gf_[j] * (1.f - gf_[j]);
                    da[j >> 1][(j & 1) * 4 + 2] = dcn * gi_[j] * (1.f - gg_[j] * gg_[j]);
                    da[j >> 1][(j & 1) * 4 + 3] = dh[j] * tcn * go_[j] * (1.f - go_[j]);
                    dcn_bits[j] = __float_as_uint(dcn * gf_[j]);
                }
                tmem_st4(lane_base + 384 + u0, dcn_bits);
                if ((hi & 1) == 0) {   // first write into quarter Q's buffer: the MMAs of quarter Q - 2 (same buffer) must have completed
                    const int Q = 4 * s + i;
                    if (Q >= 2) mbar_wait(&q_free[Q & 1], (uint32_t)(((Q >> 1) - 1) & 1));
                }
                unsigned char *buf = As + (size_t)((4 * s + i) & 1) * (2 * TCB3_QBYTES);
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    float lo[8];
                    const uint4 pk = pack8_f16(da[q]);
                    {
                        float back[8];
                        unpack8_f16(pk, back);
#pragma unroll
                        for (int e = 0; e < 8; e++) lo[e] = da[q][e] - back[e];
                    }
                    const uint4 pl = pack8_f16(lo);
                    const int kq = cq * 32 + (hi & 1) * 16 + q * 8;             // column inside the quarter buffer
                    const int chunk = (cq * 4 + i) * 4 + (hi & 1) * 2 + q;      // 8 packed gate columns of the whole tile = 16 bytes
                    *reinterpret_cast<uint4 *>(buf + canon_off(r, kq, 128)) = pk;
                    *reinterpret_cast<uint4 *>(buf + TCB3_QBYTES + canon_off(r, kq, 128)) = pl;
                    *reinterpret_cast<uint4 *>(da_blk + (size_t)chunk * 2048 + r * 16) = pk;
                    *reinterpret_cast<uint4 *>(da_blk_lo + (size_t)chunk * 2048 + r * 16) = pl;
                }
                if (hi & 1) {   // this quarter is written: let the tensor core start on it
                    fence_proxy_async();
                    tc_fence_before();
                    mbar_arrive(&a_ready[i]);
                }
            }
            tmem_st_wait();
            mbar_wait(&acc_full, s & 1);
            tc_fence_after();
        }
        scatter_dx(L - 1);   // the last step's dx
    }
    tc_fence_before();
    __syncthreads();
    if (warp == TC_CWARPS) { __syncwarp(); tmem_dealloc<512>(taddr); }
}

int launch_lstm_bwd_tc3(const LstmBwdTcArgs &a, int64_t ntiles, int smem, int smem_max, cudaStream_t st)
{
    static const int pf = [] { const char *e = getenv("PN_BWD3_PF"); return e ? atoi(e) : 0; }();
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc3<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc3<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc3<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_tc3<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_max));
    if (pf <= 0) k_lstm_bwd_tc3<0><<<(unsigned)ntiles, TCB_THREADS, smem, st>>>(a);
    else if (pf == 1) k_lstm_bwd_tc3<1><<<(unsigned)ntiles, TCB_THREADS, smem, st>>>(a);
    else if (pf == 2) k_lstm_bwd_tc3<2><<<(unsigned)ntiles, TCB_THREADS, smem, st>>>(a);
    else k_lstm_bwd_tc3<3><<<(unsigned)ntiles, TCB_THREADS, smem, st>>>(a);
    return PN_OK;
}

int launch_split_table(cudaStream_t st, const float *T, float scale, void *hi, void *lo, int64_t n)
{
    if (n % 8) { set_error("split_table: %lld values (not a multiple of 8)", (long long)n); return PN_ERR_INVALID; }
    if (n == 0) return PN_OK;
    k_split_table<<<(unsigned)ceil_div(n / 8, 256), 256, 0, st>>>(T, scale, reinterpret_cast<__half *>(hi), reinterpret_cast<__half *>(lo), n / 8);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

int launch_pack_lstm_tc3_bwd(cudaStream_t st, const float *w_ih, const float *w_hh, unsigned char *WhT)
{
    k_pack_lstm_tc3_bwd<<<(unsigned)ceil_div(TC_G4 * TC_K2, 256), 256, 0, st>>>(w_ih, w_hh, WhT);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

// Forward LSTM in split-precision mode (called from pn_pagg_forward).  w.Wp (512 KB) holds the 32 packed weight chunks,
// w.t16 the two fp16 terms of the scaled table (hi, then lo; written by pn_pagg_forward after the table projection).
int pagg_tc3_forward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const int32_t *seq_node,
                          const uint8_t *seq_type, const PaggChunk &ck, cudaStream_t st)
{
    if (c.H != TC_H || c.L > TC_MAX_L) {
        set_error("PN_MATH_TC3: the tensor-core LSTM is built for hidden size 128 and path_len <= %d (got H=%d L=%d)", TC_MAX_L, c.H, c.L);
        return PN_ERR_UNSUPPORTED;
    }
    unsigned char *Wh = reinterpret_cast<unsigned char *>(w.Wp);
    if (ck.first) {
        k_pack_lstm_tc3<<<(unsigned)ceil_div(TC_G4 * TC_K2, 256), 256, 0, st>>>(prm.w_ih, prm.w_hh, prm.b_ih, prm.b_hh, Wh, w.bp);
        PN_LAUNCH_CHECK();
    }
    uint32_t thr = 0; float scale = 1.f;
    if (c.training && c.dropout_p > 0.f) {
        thr = dropout_threshold(c.dropout_p);
        scale = c.dropout_p >= 1.f ? 0.f : 1.f / (1.f - c.dropout_p);
    }
    const int64_t P = ck.Sc * c.nw;
    const __half *t16hi = reinterpret_cast<const __half *>(w.t16), *t16lo = t16hi + (int64_t)c.N * c.L * c.H;
    LstmTcArgs a{t16hi, w.T, seq_node, seq_type, Wh, w.bp, reinterpret_cast<__half *>(w.acts), w.hfin, P, c.L, thr, scale,
                 c.dropout_seed, c.dropout_seed_dev, ck.s0 * c.nw, reinterpret_cast<unsigned char *>(w.XH),
                 reinterpret_cast<uint32_t *>(w.dmask), ck.save ? 1 : 0, nullptr, 0, t16lo,
                 reinterpret_cast<unsigned char *>(w.XH) + ceil_div(P, 128) * c.L * (int64_t)XH_BLOCK_BYTES, w.acts};
    long long *d_trace = nullptr;
    const char *trace_env = getenv("PN_TRACE_FWD");   // debug: phase timestamps (clock64) of one CTA
    if (trace_env) {
        PN_CUDA(cudaMalloc(&d_trace, sizeof(long long) * 8 * 64));
        PN_CUDA(cudaMemsetAsync(d_trace, 0, sizeof(long long) * 8 * 64, st));
        a.trace = d_trace;
        a.trace_block = atoi(trace_env);
    }
    PN_CUDA(cudaFuncSetAttribute(k_lstm_fwd_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM));
    {
        ProfScope prof("lstm_fwd", st);
        k_lstm_fwd_tc3<<<(unsigned)ceil_div(P, 128), TCB_THREADS, TC3_SMEM, st>>>(a);
    }
    PN_LAUNCH_CHECK();
    if (d_trace) {
        long long h[8 * 64];
        PN_CUDA(cudaStreamSynchronize(st));
        PN_CUDA(cudaMemcpy(h, d_trace, sizeof(h), cudaMemcpyDeviceToHost));
        cudaFree(d_trace);
        const long long t00 = h[0];
        for (int t = 0; t < c.L && t < 8; t++)
            fprintf(stderr, "[trace fwd3 cta %d] t=%d  begin %lld | gather loads issued %lld | mma saw a_ready %lld | last chunk reached %lld | acc_full seen %lld | x stored %lld | gate math done %lld\n",
                    a.trace_block, t, h[t * 8 + 0] - t00, h[t * 8 + 1] - t00, h[t * 8 + 4] - t00, h[t * 8 + 5] - t00, h[t * 8 + 2] - t00, h[t * 8 + 6] - t00, h[t * 8 + 3] - t00);
    }
    return PN_OK;
}

}  // namespace pn
