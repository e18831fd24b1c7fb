// Split-precision tensor-core math mode of PAGG ("tc3", PN_MATH_TC3): the LSTM recurrence on tcgen05 tensor cores at
// fp32-grade accuracy.  Every fp32 operand v is carried as two fp16 terms, hi = fp16(v) and lo = fp16(v - hi)
// (22 significant bits together), and a product is three tensor-core passes into the same fp32 TMEM accumulator,
//     A W  ~=  A_hi W_hi + A_lo W_hi + A_hi W_lo            (the dropped A_lo W_lo term is 2^-22 relative),
// the gates use ex2.approx / rcp.approx (1-2 ulp) instead of tanh.approx (2^-11).  The projections around the LSTM run
// on the strict fp32 kernels of pagg.cu in this mode; the backward is k_lstm_bwd_tc<.., SPLITW = true> (pagg_tc.cu).
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

__global__ void __launch_bounds__(TC_THREADS, 1) k_lstm_fwd_tc3(LstmTcArgs a)
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
        // ===== MMA issuer (whole warp converged, one elected lane issues; see k_lstm_fwd_tc) =====
        constexpr uint32_t idesc = make_idesc_f16(128, 256);
        uint32_t stage = 0, phase = 0;
        for (int t = 0; t < L; t++) {
            mbar_wait(&a_ready, t & 1);  // x_t (and h_{t-1}) staged, TMEM drained
            tc_fence_after();
            const int nch = t == 0 ? TC_NCHUNK / 2 : TC_NCHUNK;
            for (int c = 0; c < nch; c++) {
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
                if (a.save) *reinterpret_cast<uint4 *>(xh_blk + (size_t)k8 * 2048 + r * 16) = gh[kk];
            }
            if (a.dmask && a.save && a.drop_thr) a.dmask[(((int64_t)blockIdx.x * L + t_) * 4 + cq) * 128 + r] = gmask;
        };
        gather_load(0);
        gather_store(0);
        // h_{-1} = 0: h half of step 0's image
#pragma unroll
        for (int k8 = 16 + cq * 4; k8 < 16 + cq * 4 + 4; k8++)
            if (a.save) *reinterpret_cast<uint4 *>(a.xh_img + (int64_t)blockIdx.x * XH_BLOCK_BYTES + (size_t)k8 * 2048 + r * 16) = make_uint4(0, 0, 0, 0);
        for (int t = 0; t < L; t++) {
            fence_proxy_async();   // generic-proxy smem writes (x_t, h_{t-1}) -> async proxy (tcgen05.mma)
            tc_fence_before();     // order this thread's earlier tcgen05.ld before the MMA that overwrites TMEM
            mbar_arrive(&a_ready);
            if (t + 1 < L) gather_load(t + 1);   // in flight while the tensor core works on step t
            mbar_wait(&acc_full, t & 1);
            tc_fence_after();
            if (t + 1 < L) gather_store(t + 1);  // every MMA of step t has completed: the x tiles are free
            __half *act = a.acts + ((int64_t)blockIdx.x * L + t) * ACT_BLOCK_HALFS + r * 8;
            const bool t_is_last = t == L - 1;
#pragma unroll
            for (int jj = 0; jj < 4; jj++) {
                const int c0 = cq * 128 + jj * 32;   // packed gate columns of units [c0/4, c0/4 + 8)
                uint32_t pi[4], pf[4], pg[4], po[4], pc[4], ph[4], pl[4];
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
                    t2 = __floats2half2_rn(gi2[0], gi2[1]); pi[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(gf2[0], gf2[1]); pf[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(gg2[0], gg2[1]); pg[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(go2[0], go2[1]); po[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(cn2[0], cn2[1]); pc[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    t2 = __floats2half2_rn(hh2[0], hh2[1]); ph[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    const float2 back = __half22float2(t2);
                    t2 = __floats2half2_rn(hh2[0] - back.x, hh2[1] - back.y); pl[u2] = *reinterpret_cast<uint32_t *>(&t2);
                    hq[(u2 & 1) * 2] = hh2[0]; hq[(u2 & 1) * 2 + 1] = hh2[1];
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
                    if (a.save) *reinterpret_cast<uint4 *>(a.xh_img + ((int64_t)(t + 1) * ntiles + blockIdx.x) * XH_BLOCK_BYTES + (size_t)(16 + c0 / 32) * 2048 + r * 16) = hpk;
                }
                if (valid && a.save) {
                    const int ub = (c0 / 32) * 1024;
                    *reinterpret_cast<uint4 *>(act + 0 * 16384 + ub) = make_uint4(pi[0], pi[1], pi[2], pi[3]);
                    *reinterpret_cast<uint4 *>(act + 1 * 16384 + ub) = make_uint4(pf[0], pf[1], pf[2], pf[3]);
                    *reinterpret_cast<uint4 *>(act + 2 * 16384 + ub) = make_uint4(pg[0], pg[1], pg[2], pg[3]);
                    *reinterpret_cast<uint4 *>(act + 3 * 16384 + ub) = make_uint4(po[0], po[1], po[2], po[3]);
                    *reinterpret_cast<uint4 *>(act + 4 * 16384 + ub) = make_uint4(pc[0], pc[1], pc[2], pc[3]);
                }
            }
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
                 reinterpret_cast<uint32_t *>(w.dmask), ck.save ? 1 : 0, nullptr, 0, t16lo};
    PN_CUDA(cudaFuncSetAttribute(k_lstm_fwd_tc3, cudaFuncAttributeMaxDynamicSharedMemorySize, TC3_SMEM));
    {
        ProfScope prof("lstm_fwd", st);
        k_lstm_fwd_tc3<<<(unsigned)ceil_div(P, 128), TC_THREADS, TC3_SMEM, st>>>(a);
    }
    PN_LAUNCH_CHECK();
    return PN_OK;
}

}  // namespace pn
