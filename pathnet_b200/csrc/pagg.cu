// PAGG operator (forward + backward) — replaces PathNet.forward /
// PathNet_homo.forward (PathNet_run.py:172-211, :239-278) and their autograd
// backward (:351).  fp32 CUDA-core math mode; the tensor-core mode lives in
// pagg_tc.cu and shares this file's index handling, pooling and parameter-grad
// plumbing.
//
// Canonical form: after pn_pagg_canonicalize both variants are
//   Z = act0(X W0^T + b0)                       [N,H]
//   T[n,d] = act1(W_d Z[n] + b_d)               [N,L,H]  (hoisted: the reference
//            applies all L Linears to every gathered row and then selects one)
//   x_{s,w,t} = dropout(T[seq_node[s,w,t], seq_type[s,w,t]])
//   h_{s,w} = LSTM_final_h(x_{s,w,0..L-1})
//   score = attw . [h ; partner] + b
//   pooled[s] = 1/nw sum_w weight(score)_w h_{s,w}
//   logits = fc2(dropout([Z[rows[s]] ; pooled[s]]))
#include <algorithm>

#include <mutex>

#include "pagg_common.cuh"

namespace pn {

// effective dropout seed of a kernel-argument struct (see pn_pagg_cfg.dropout_seed_dev)
#define PN_SEED(a) ((a).seed_dev ? (a).drop_seed + *(a).seed_dev : (a).drop_seed)

// ---------------------------------------------------------------------------
// canonicalisation (integer only)
// ---------------------------------------------------------------------------
// NT / TT: element types of the caller's index tensors (int64 as the reference passes them, or the sampler's
// int32 node ids / uint8 labels).  Out-of-range entries set *status and become 0 (see the header).
template <typename NT, typename TT>
__global__ void k_canonicalize(int variant, int64_t N, int64_t S, int nw, int L, const NT *__restrict__ neis,
                               const TT *__restrict__ ltype, int32_t *__restrict__ seq_node,
                               uint8_t *__restrict__ seq_type, int32_t *__restrict__ partner, int32_t *status)
{
    const int64_t P = S * nw;
    const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;  // over S*nw*L
    if (i >= P * L) return;
    const int t = (int)(i % L);
    const int64_t sw = i / L;  // s*nw + w
    int64_t node, type, pnode = 0;
    if (variant != PN_VARIANT_HETERO) {
        // sequence (s,w) = path p = s*nw+w in natural order (:246-265; also baseline/GPRGNN/src/copy.py:335-349)
        node = (int64_t)neis[i];
        type = (int64_t)ltype[i];
        if (t == 0) pnode = (int64_t)neis[sw * L];  // step-0 entry of the same path (:259-260)
    } else {
        // output row s pools LSTM batch elements b = w*S + s (:196-197); batch
        // element b, step t is flat row r = b*L + t of the flipped/reshaped
        // gather (:182-183,:191): node idx[r mod P][L-1 - r div P], type ltype.flat[r]
        const int64_t s = sw / nw, w = sw % nw;
        const int64_t b = w * S + s;
        const int64_t r = b * L + t;
        node = (int64_t)neis[(r % P) * L + (L - 1 - r / P)];
        type = (int64_t)ltype[r];
        if (t == 0) pnode = (int64_t)neis[b * L];  // neis[0] row b (:199)
    }
    int bad = 0;
    if (node < 0 || node >= N) { bad |= 1; node = 0; }
    if (type < 0 || type >= L) { bad |= 2; type = 0; }
    if (t == 0 && (pnode < 0 || pnode >= N)) { bad |= 1; pnode = 0; }
    seq_node[i] = (int32_t)node;
    seq_type[i] = (uint8_t)type;
    if (t == 0) partner[sw] = (int32_t)pnode;
    if (bad && status) atomicOr(status, bad);
}

// ---------------------------------------------------------------------------
// LSTM weight packing: Wp[k][unit*4+gate] (k < H: W_ih^T, k >= H: W_hh^T)
// ---------------------------------------------------------------------------
// cell = PN_CELL_RNN (tanh RNN, baseline/GPRGNN/src/copy.py:307): w_ih / w_hh are [H,H]; they occupy the "g" slot
// (gate 2, the tanh gate) of the packed layout and the other three slots are zero, so the same kernels serve both cells.
// cell = PN_CELL_GRU (torch nn.GRU: w_ih / w_hh [3H,H] in r, z, n order): slot 0 = r, slot 1 = z (x and h rows summed by the
// GEMM, as the LSTM gates are), slot 2 = the x part of n (W_in, b_in; its h rows are zero), slot 3 = the h part of n
// (W_hn, b_hn; its x rows are zero) -- n = tanh(slot2 + r * slot3) needs the two parts separately.
__global__ void k_pack_lstm(int H, int cell, const float *__restrict__ w_ih, const float *__restrict__ w_hh,
                            const float *__restrict__ b_ih, const float *__restrict__ b_hh,
                            float *__restrict__ Wp, float *__restrict__ WpT, float *__restrict__ bp)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int G4 = 4 * H, K2 = 2 * H;
    if (i < K2 * G4) {
        int k = i / G4, col = i % G4, unit = col >> 2, gate = col & 3;
        float v;
        if (cell == PN_CELL_RNN) v = gate != 2 ? 0.f : (k < H ? w_ih[unit * H + k] : w_hh[unit * H + (k - H)]);
        else if (cell == PN_CELL_GRU) {
            if (gate < 2) v = k < H ? w_ih[(gate * H + unit) * H + k] : w_hh[(gate * H + unit) * H + (k - H)];
            else if (gate == 2) v = k < H ? w_ih[(2 * H + unit) * H + k] : 0.f;
            else v = k < H ? 0.f : w_hh[(2 * H + unit) * H + (k - H)];
        }
        else v = k < H ? w_ih[(gate * H + unit) * H + k] : w_hh[(gate * H + unit) * H + (k - H)];
        Wp[i] = v;
        WpT[col * K2 + k] = v;
    }
    if (i < G4) {
        int unit = i >> 2, gate = i & 3;
        if (cell == PN_CELL_RNN) bp[i] = gate != 2 ? 0.f : b_ih[unit] + b_hh[unit];
        else if (cell == PN_CELL_GRU) bp[i] = gate < 2 ? b_ih[gate * H + unit] + b_hh[gate * H + unit] : gate == 2 ? b_ih[2 * H + unit] : b_hh[2 * H + unit];
        else bp[i] = b_ih[gate * H + unit] + b_hh[gate * H + unit];
    }
}

__global__ void k_unpack_lstm_grads(int H, int cell, const float *__restrict__ dWp, const float *__restrict__ dbp,
                                    float *__restrict__ dw_ih, float *__restrict__ dw_hh,
                                    float *__restrict__ db_ih, float *__restrict__ db_hh)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int G4 = 4 * H;
    if (cell == PN_CELL_RNN) {
        if (i < H * H) {
            const int unit = i / H, k = i % H, col = unit * 4 + 2;
            dw_ih[i] = dWp[k * G4 + col];
            dw_hh[i] = dWp[(H + k) * G4 + col];
        }
        if (i < H) { const float v = dbp[i * 4 + 2]; db_ih[i] = v; db_hh[i] = v; }
        return;
    }
    if (cell == PN_CELL_GRU) {
        if (i < 3 * H * H) {
            const int row = i / H, k = i % H, gate = row / H, unit = row % H;
            dw_ih[i] = dWp[k * G4 + unit * 4 + (gate < 2 ? gate : 2)];
            dw_hh[i] = dWp[(H + k) * G4 + unit * 4 + (gate < 2 ? gate : 3)];
        }
        if (i < 3 * H) {
            const int gate = i / H, unit = i % H;
            db_ih[i] = dbp[unit * 4 + (gate < 2 ? gate : 2)];
            db_hh[i] = dbp[unit * 4 + (gate < 2 ? gate : 3)];
        }
        return;
    }
    if (i < G4 * H) {
        int row = i / H, k = i % H, gate = row / H, unit = row % H;
        int col = unit * 4 + gate;
        dw_ih[i] = dWp[k * G4 + col];
        dw_hh[i] = dWp[(H + k) * G4 + col];
    }
    if (i < G4) {
        int gate = i / H, unit = i % H;
        float v = dbp[unit * 4 + gate];
        db_ih[i] = v;
        db_hh[i] = v;
    }
}

// ---------------------------------------------------------------------------
// fused gather + dropout + LSTM forward, fp32 FMA
//   CTA = 64 sequences, all L steps; 256 threads; thread (ty, tx) owns
//   sequences ty*8..ty*8+7 and hidden units {tx + 32u}, all four gates.
// ---------------------------------------------------------------------------
struct LstmArgs {
    const float *T;          // [N,L,H]
    const int32_t *seq_node; // [P,L]
    const uint8_t *seq_type; // [P,L]
    const float *Wp, *WpT, *bp;
    float *acts;             // [L][P][5][H] (null: inference, nothing saved)
    float *hfin;             // [P,H]
    int64_t P;
    int L;
    // dropout of the LSTM input (stream 0, element index (p*L+t)*H + j)
    uint32_t drop_thr;       // drop iff hash < thr; 0 = no dropout
    float drop_scale;
    uint64_t drop_seed;
    const uint64_t *seed_dev; // optional device word added to drop_seed (CUDA-graph replay)
    int64_t p_base;          // global index of this chunk's first sequence (row-chunked calls): mask indices are global
    int cell;                // PN_CELL_LSTM | PN_CELL_RNN (h_t = tanh of the "g" slot; the saved "g" activation is h_t) |
                             // PN_CELL_GRU (saved: r, z, n, the h part of n's pre-activation, h_t)
    // backward only
    const float *dhfin;      // [P,H]
    float *dT;               // [N,L,H]
    float *dA;               // [L*P,4H]
    float *XH;               // [L*P,2H]
};

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + expf(-x)); }

template <int UPT>
__device__ __forceinline__ void gather_x_tile(const LstmArgs &a, float *As, int apitch, int64_t p0,
                                              int t, int ty, int tx)
{
    constexpr int H = 32 * UPT;
#pragma unroll
    for (int pp = 0; pp < 8; pp++) {
        const int row = ty * 8 + pp;
        const int64_t p = p0 + row;
        float *dst = As + row * apitch;
        if (p < a.P) {
            const int64_t e = p * a.L + t;
            const float *src = a.T + ((int64_t)a.seq_node[e] * a.L + a.seq_type[e]) * H;
            for (int j = tx * 4; j < H; j += 128) {
                float4 v = __ldg(reinterpret_cast<const float4 *>(src + j));
                if (a.drop_thr) {
                    const uint64_t pair = ((uint64_t)(e + a.p_base * a.L) * H + j) >> 1;  // (e*H + j) is even
                    const uint64_t seed = PN_SEED(a);
                    const uint32_t h0 = dropout_hash_pair(seed, 0, pair), h1 = dropout_hash_pair(seed, 0, pair + 1);
                    v.x = (h0 & 0xFFFFu) < a.drop_thr ? 0.f : v.x * a.drop_scale;
                    v.y = (h0 >> 16) < a.drop_thr ? 0.f : v.y * a.drop_scale;
                    v.z = (h1 & 0xFFFFu) < a.drop_thr ? 0.f : v.z * a.drop_scale;
                    v.w = (h1 >> 16) < a.drop_thr ? 0.f : v.w * a.drop_scale;
                }
                *reinterpret_cast<float4 *>(dst + j) = v;
            }
        } else {
            for (int j = tx * 4; j < H; j += 128) *reinterpret_cast<float4 *>(dst + j) = make_float4(0, 0, 0, 0);
        }
    }
}

template <int UPT>
__global__ void __launch_bounds__(256) k_lstm_fwd_f32(LstmArgs a)
{
    constexpr int H = 32 * UPT, K2 = 2 * H, G4 = 4 * H, AP = K2 + 4;
    extern __shared__ __align__(16) float smem[];
    float *As = smem;            // [64][AP]   [x_t | h_{t-1}] per sequence
    float *Bs = As + 64 * AP;    // [16][G4]   k-chunk of Wp
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * 64;

    float c[8][UPT];
#pragma unroll
    for (int pp = 0; pp < 8; pp++)
#pragma unroll
        for (int u = 0; u < UPT; u++) c[pp][u] = 0.f;
    // h_0 = 0
    for (int i = tid; i < 64 * H; i += 256) As[(i / H) * AP + H + (i % H)] = 0.f;

    for (int t = 0; t < a.L; t++) {
        gather_x_tile<UPT>(a, As, AP, p0, t, ty, tx);
        float acc[8][UPT][4];
#pragma unroll
        for (int u = 0; u < UPT; u++) {
            const float4 b = *reinterpret_cast<const float4 *>(a.bp + (u * 32 + tx) * 4);
#pragma unroll
            for (int pp = 0; pp < 8; pp++) { acc[pp][u][0] = b.x; acc[pp][u][1] = b.y; acc[pp][u][2] = b.z; acc[pp][u][3] = b.w; }
        }
        for (int k0 = 0; k0 < K2; k0 += 16) {
            __syncthreads();  // previous chunk consumed (first pass: gather + h visible)
            for (int i = tid; i < 16 * G4 / 4; i += 256)
                reinterpret_cast<float4 *>(Bs)[i] = __ldg(reinterpret_cast<const float4 *>(a.Wp + (int64_t)k0 * G4) + i);
            __syncthreads();
#pragma unroll
            for (int k4 = 0; k4 < 16; k4 += 4) {
                float4 av[8];
#pragma unroll
                for (int pp = 0; pp < 8; pp++) av[pp] = *reinterpret_cast<const float4 *>(As + (ty * 8 + pp) * AP + k0 + k4);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
#pragma unroll
                    for (int u = 0; u < UPT; u++) {
                        const float4 b = *reinterpret_cast<const float4 *>(Bs + (k4 + kk) * G4 + (u * 32 + tx) * 4);
#pragma unroll
                        for (int pp = 0; pp < 8; pp++) {
                            const float x = kk == 0 ? av[pp].x : kk == 1 ? av[pp].y : kk == 2 ? av[pp].z : av[pp].w;
                            acc[pp][u][0] = fmaf(x, b.x, acc[pp][u][0]);
                            acc[pp][u][1] = fmaf(x, b.y, acc[pp][u][1]);
                            acc[pp][u][2] = fmaf(x, b.z, acc[pp][u][2]);
                            acc[pp][u][3] = fmaf(x, b.w, acc[pp][u][3]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // everyone is done reading As before h_t / x_{t+1} overwrite it
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const int row = ty * 8 + pp;
            const int64_t p = p0 + row;
#pragma unroll
            for (int u = 0; u < UPT; u++) {
                const int unit = u * 32 + tx;
                float ig, fg, gg, og, cn, h;
                if (a.cell == PN_CELL_RNN) {   // h_t = tanh(W_ih x_t + b_ih + W_hh h_{t-1} + b_hh)  (torch nn.RNN)
                    gg = tanhf(acc[pp][u][2]);
                    ig = fg = og = cn = 0.f;
                    h = gg;
                } else if (a.cell == PN_CELL_GRU) {   // torch nn.GRU; this thread is the only reader / writer of its h element
                    const float hprev = As[row * AP + H + unit];
                    ig = sigmoidf_(acc[pp][u][0]);                      // r
                    fg = sigmoidf_(acc[pp][u][1]);                      // z
                    og = acc[pp][u][3];                                 // W_hn h + b_hn
                    gg = tanhf(acc[pp][u][2] + ig * og);                // n
                    h = (1.f - fg) * gg + fg * hprev;
                    cn = h;
                } else {
                    ig = sigmoidf_(acc[pp][u][0]); fg = sigmoidf_(acc[pp][u][1]);
                    gg = tanhf(acc[pp][u][2]); og = sigmoidf_(acc[pp][u][3]);
                    cn = fg * c[pp][u] + ig * gg;
                    c[pp][u] = cn;
                    h = og * tanhf(cn);
                }
                As[row * AP + H + unit] = h;
                if (p < a.P) {
                    if (a.acts) {
                        float *act = a.acts + (((int64_t)t * a.P + p) * 5) * H + unit;
                        act[0] = ig; act[H] = fg; act[2 * H] = gg; act[3 * H] = og; act[4 * H] = cn;
                    }
                    if (t == a.L - 1) a.hfin[p * H + unit] = h;
                }
            }
        }
        // the k0 = 0 barrier of the next step orders these writes before any read
    }
}

// ---------------------------------------------------------------------------
// LSTM backward (BPTT), fp32 FMA.  Same CTA/thread ownership as forward.
// Per step: da (pre-activation gate grads) -> smem + global dA; XH row
// [x_t | h_{t-1}] -> global (weight-grad GEMM operand); [dx | dh_{t-1}] =
// da . Wp^T; dx scattered into dT with red.global.add.
// ---------------------------------------------------------------------------
template <int UPT>
__global__ void __launch_bounds__(256) k_lstm_bwd_f32(LstmArgs a)
{
    constexpr int H = 32 * UPT, K2 = 2 * H, G4 = 4 * H, DP = G4 + 4;
    constexpr int NQ = K2 / 32;  // outputs per thread per sequence
    extern __shared__ __align__(16) float smem[];
    float *Ds = smem;            // [64][DP]  da per sequence, packed columns
    float *Bs = Ds + 64 * DP;    // [16][K2]  chunk of WpT
    const int tid = threadIdx.x, tx = tid & 31, ty = tid >> 5;
    const int64_t p0 = (int64_t)blockIdx.x * 64;

    float dh[8][UPT], dc[8][UPT];
#pragma unroll
    for (int pp = 0; pp < 8; pp++) {
        const int64_t p = p0 + ty * 8 + pp;
#pragma unroll
        for (int u = 0; u < UPT; u++) {
            dh[pp][u] = p < a.P ? a.dhfin[p * H + u * 32 + tx] : 0.f;
            dc[pp][u] = 0.f;
        }
    }
    for (int t = a.L - 1; t >= 0; t--) {
        // ---- gate grads from the saved activations ----
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const int row = ty * 8 + pp;
            const int64_t p = p0 + row;
#pragma unroll
            for (int u = 0; u < UPT; u++) {
                const int unit = u * 32 + tx;
                float4 da = make_float4(0, 0, 0, 0);
                if (p < a.P) {
                    const float *act = a.acts + (((int64_t)t * a.P + p) * 5) * H + unit;
                    const float ig = act[0], fg = act[H], gg = act[2 * H], og = act[3 * H], cn = act[4 * H];
                    float cprev = 0.f, hprev = 0.f;
                    if (t > 0) {
                        const float *actp = a.acts + (((int64_t)(t - 1) * a.P + p) * 5) * H + unit;
                        cprev = actp[4 * H];
                        hprev = a.cell == PN_CELL_RNN ? actp[2 * H] : a.cell == PN_CELL_GRU ? cprev : actp[3 * H] * tanhf(cprev);
                    }
                    if (a.cell == PN_CELL_RNN) {
                        da.z = dh[pp][u] * (1.f - gg * gg);     // h_t = gg = tanh(pre-activation)
                    } else if (a.cell == PN_CELL_GRU) {         // ig = r, fg = z, gg = n, og = W_hn h + b_hn
                        const float d = dh[pp][u];
                        const float dn = d * (1.f - fg), dz = d * (hprev - gg);
                        da.z = dn * (1.f - gg * gg);            // x part of n's pre-activation
                        da.w = da.z * ig;                       // h part of it
                        da.x = da.z * og * ig * (1.f - ig);     // r
                        da.y = dz * fg * (1.f - fg);            // z
                        dc[pp][u] = d * fg;                     // the direct path h_{t-1} -> h_t, added to dh_{t-1} below
                    } else {
                        const float tc = tanhf(cn);
                        const float dcn = dc[pp][u] + dh[pp][u] * og * (1.f - tc * tc);
                        da.x = dcn * gg * ig * (1.f - ig);          // i
                        da.y = dcn * cprev * fg * (1.f - fg);       // f
                        da.z = dcn * ig * (1.f - gg * gg);          // g
                        da.w = dh[pp][u] * tc * og * (1.f - og);    // o
                        dc[pp][u] = dcn * fg;
                    }
                    const int64_t r = (int64_t)t * a.P + p;
                    *reinterpret_cast<float4 *>(a.dA + r * G4 + unit * 4) = da;
                    a.XH[r * K2 + H + unit] = hprev;
                }
                *reinterpret_cast<float4 *>(Ds + row * DP + unit * 4) = da;
            }
        }
        // ---- x_t (regathered, same dropout mask) into XH ----
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const int64_t p = p0 + ty * 8 + pp;
            if (p < a.P) {
                const int64_t e = p * a.L + t;
                const float *src = a.T + ((int64_t)a.seq_node[e] * a.L + a.seq_type[e]) * H;
                float *dst = a.XH + ((int64_t)t * a.P + p) * K2;
                for (int j = tx; j < H; j += 32) {
                    float v = __ldg(src + j);
                    if (a.drop_thr) v = dropout_hash(PN_SEED(a), 0, (uint64_t)(e + a.p_base * a.L) * H + j) < a.drop_thr ? 0.f : v * a.drop_scale;
                    dst[j] = v;
                }
            }
        }
        // ---- [dx | dh_prev] = da . WpT ----
        float acc[8][NQ];
#pragma unroll
        for (int pp = 0; pp < 8; pp++)
#pragma unroll
            for (int q = 0; q < NQ; q++) acc[pp][q] = 0.f;
        for (int c0 = 0; c0 < G4; c0 += 16) {
            __syncthreads();
            for (int i = tid; i < 16 * K2 / 4; i += 256)
                reinterpret_cast<float4 *>(Bs)[i] = __ldg(reinterpret_cast<const float4 *>(a.WpT + (int64_t)c0 * K2) + i);
            __syncthreads();
#pragma unroll
            for (int c4 = 0; c4 < 16; c4 += 4) {
                float4 dv[8];
#pragma unroll
                for (int pp = 0; pp < 8; pp++) dv[pp] = *reinterpret_cast<const float4 *>(Ds + (ty * 8 + pp) * DP + c0 + c4);
#pragma unroll
                for (int cc = 0; cc < 4; cc++) {
#pragma unroll
                    for (int q = 0; q < NQ; q++) {
                        const float b = Bs[(c4 + cc) * K2 + q * 32 + tx];
#pragma unroll
                        for (int pp = 0; pp < 8; pp++) {
                            const float d = cc == 0 ? dv[pp].x : cc == 1 ? dv[pp].y : cc == 2 ? dv[pp].z : dv[pp].w;
                            acc[pp][q] = fmaf(d, b, acc[pp][q]);
                        }
                    }
                }
            }
        }
        __syncthreads();  // Ds free for the next step
        // ---- consume: dh_{t-1} stays in registers, dx goes to dT ----
#pragma unroll
        for (int pp = 0; pp < 8; pp++) {
            const int64_t p = p0 + ty * 8 + pp;
#pragma unroll
            for (int u = 0; u < UPT; u++) dh[pp][u] = acc[pp][UPT + u] + (a.cell == PN_CELL_GRU ? dc[pp][u] : 0.f);
            if (p < a.P) {
                const int64_t e = p * a.L + t;
                float *dst = a.dT + ((int64_t)a.seq_node[e] * a.L + a.seq_type[e]) * H;
#pragma unroll
                for (int q = 0; q < UPT; q++) {
                    const int j = q * 32 + tx;
                    float v = acc[pp][q];
                    if (a.drop_thr) v = dropout_hash(PN_SEED(a), 0, (uint64_t)(e + a.p_base * a.L) * H + j) < a.drop_thr ? 0.f : v * a.drop_scale;
                    atomicAdd(dst + j, v);
                }
            }
        }
    }
}

// ---------------------------------------------------------------------------
// attention + pooling + concat (+ dropout): one warp per output row
// ---------------------------------------------------------------------------
struct PoolArgs {
    int variant, H, nw, L;
    int64_t S;
    const float *Z, *T, *hfin;
    const int32_t *seq_node, *partner, *rows;
    const uint8_t *seq_type;
    const float *attw_w, *attw_b;
    float *score, *att, *cat;
    uint32_t drop_thr; float drop_scale; uint64_t drop_seed;
    const uint64_t *seed_dev;   // optional device word added to drop_seed
    int64_t s_base;             // global index of this chunk's first output row (mask indices are global)
    int pool;                   // PN_POOL_ATTENTION (the variant's own attention) | MEAN | SUM | MAX over the nw sequences
    // backward
    const float *dcat;   // [S,2H] grad wrt the dropped concat
    float *dhfin, *dZ, *dT, *dattw;   // dhfin may be null when the factored form below is written instead
    // factored dh_final for the tensor-core LSTM backward: dh_final[p] = dh_k[p] * gdrop[row(p)] + dh_s[p] * attw_w[0:H]
    // (dh_k = att/nw, dh_s = dscore; gdrop = the row's pooled-half gradient after dropout, leading dimension gdrop_ld),
    // 2 scalars per path + one vector per ROW instead of a [P,H] tensor written here and read back there
    float *dh_k, *dh_s, *gdrop;
    int64_t gdrop_ld;
    unsigned int *absmax_bits;  // max |dhfin| as float bits (tensor-core backward's gradient scale), may be null
    // classifier fused into the H = 128 kernels (few classes: each warp already holds its whole concat row)
    int C;
    const float *fc2_w, *fc2_b;   // [C][2H], [C]
    float *logits;                // forward: [S][C], null = not fused
    const float *dlogits;         // backward: [S][C], null = not fused (dcat is read instead)
    float *dfc2_w, *dfc2_b;       // backward: atomically accumulated
};
constexpr int POOL_FUSE_MAX_C = 16;

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

__device__ __forceinline__ const float *partner_vec(const PoolArgs &a, int64_t p)
{
    if (a.variant == PN_VARIANT_HOMO)  // step-0 table entry, before dropout (:259-260)
        return a.T + ((int64_t)a.seq_node[p * a.L] * a.L + a.seq_type[p * a.L]) * a.H;
    return a.Z + (int64_t)a.partner[p] * a.H;  // un-projected root embedding (:199)
}

__global__ void __launch_bounds__(256) k_pool_fwd(PoolArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= a.S) return;
    const int H = a.H, nw = a.nw;
    const float bias = a.attw_b[0];
    // raw scores
    float mx = -INFINITY;
#pragma unroll 4
    for (int w = 0; w < nw; w++) {
        const int64_t p = s * nw + w;
        const float *h = a.hfin + p * H, *pv = partner_vec(a, p);
        float d = 0.f;
        for (int j = lane; j < H; j += 32) d += a.attw_w[j] * h[j] + a.attw_w[H + j] * pv[j];
        d = warp_sum(d) + bias;
        if (lane == 0) a.score[p] = d;
        const float l = d > 0.f ? d : 0.01f * d;
        mx = fmaxf(mx, l);
    }
    __syncwarp();
    float denom = 0.f;
    if (a.variant == PN_VARIANT_HETERO) {
        for (int w = lane; w < nw; w += 32) {
            const float d = a.score[s * nw + w];
            denom += expf((d > 0.f ? d : 0.01f * d) - mx);
        }
        denom = warp_sum(denom);
    }
    for (int w = lane; w < nw; w += 32) {
        const int64_t p = s * nw + w;
        const float d = a.score[p];
        a.att[p] = a.variant == PN_VARIANT_HETERO ? expf((d > 0.f ? d : 0.01f * d) - mx) / denom : 1.f + d;
    }
    __syncwarp();
    const float inv_nw = 1.f / (float)nw;
    const int64_t node = a.rows[s];
    for (int j = lane; j < H; j += 32) {
        float pooled = 0.f;
#pragma unroll 8
        for (int w = 0; w < nw; w++) pooled += a.att[s * nw + w] * a.hfin[(s * nw + w) * H + j];
        pooled *= inv_nw;
        float ego = a.Z[node * H + j];
        if (a.drop_thr) {
            ego = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + j) < a.drop_thr ? 0.f : ego * a.drop_scale;
            pooled = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + H + j) < a.drop_thr ? 0.f : pooled * a.drop_scale;
        }
        a.cat[s * 2 * H + j] = ego;
        a.cat[s * 2 * H + H + j] = pooled;
    }
}

// backward of the above.  dattw is accumulated per CTA in shared memory.
__global__ void __launch_bounds__(256) k_pool_bwd(PoolArgs a)
{
    extern __shared__ float s_dattw[];  // [2H+1]
    const int H = a.H, nw = a.nw;
    for (int i = threadIdx.x; i < 2 * H + 1; i += blockDim.x) s_dattw[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s < a.S) {
        const float inv_nw = 1.f / (float)nw;
        const int64_t node = a.rows[s];
        // gradient wrt the pooled half of the concat, dropout applied once (H <= 128: j = lane + 32 q)
        float gq[4];
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = lane + 32 * q;
            float g = 0.f;
            if (j < H) {
                g = a.dcat[s * 2 * H + H + j];
                if (a.drop_thr) g = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + H + j) < a.drop_thr ? 0.f : g * a.drop_scale;
            }
            gq[q] = g;
        }
        // da_w = inv_nw * <dpooled, h_w>, kept by lane (w mod 32) for w < 64, else recomputed
        float da_keep0 = 0.f, da_keep1 = 0.f;
        float dot_sum = 0.f;  // sum_v att_v * da_v (softmax backward), hetero only
#pragma unroll 4
        for (int w = 0; w < nw; w++) {
            const int64_t p = s * nw + w;
            float d = 0.f;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = lane + 32 * q;
                if (j < H) d += gq[q] * a.hfin[p * H + j];
            }
            d = warp_sum(d) * inv_nw;
            if ((w & 31) == lane) { if (w < 32) da_keep0 = d; else if (w < 64) da_keep1 = d; }
            dot_sum += a.att[p] * d;
        }
        for (int j = lane; j < H; j += 32) {  // ego gradient
            float g = a.dcat[s * 2 * H + j];
            if (a.drop_thr) g = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + j) < a.drop_thr ? 0.f : g * a.drop_scale;
            atomicAdd(a.dZ + node * H + j, g);
        }
        float dbias = 0.f, amax = 0.f;
        float dw_h[4] = {0.f, 0.f, 0.f, 0.f}, dw_p[4] = {0.f, 0.f, 0.f, 0.f};
        // partner gradient: consecutive sequences often share the partner row (homo: always the row's own
        // step-0 entry), so accumulate in registers and flush with atomics only when the target changes
        float dp_acc[4] = {0.f, 0.f, 0.f, 0.f};
        float *dp_cur = nullptr;
#pragma unroll 2
        for (int w = 0; w < nw; w++) {
            const int64_t p = s * nw + w;
            const float *h = a.hfin + p * H, *pv = partner_vec(a, p);
            float hq[4];
#pragma unroll
            for (int q = 0; q < 4; q++) hq[q] = (lane + 32 * q < H) ? h[lane + 32 * q] : 0.f;
            float da;
            if (w < 64) da = __shfl_sync(0xffffffffu, w < 32 ? da_keep0 : da_keep1, w & 31);
            else {
                float d = 0.f;
#pragma unroll
                for (int q = 0; q < 4; q++) d += gq[q] * hq[q];
                da = warp_sum(d) * inv_nw;
            }
            const float raw = a.score[p], att = a.att[p];
            float dscore;
            if (a.variant == PN_VARIANT_HETERO) {
                const float dl = att * (da - dot_sum);
                dscore = dl * (raw > 0.f ? 1.f : 0.01f);
            } else {
                dscore = da;  // weight = 1 + score
            }
            dbias += dscore;
            float *dpv = a.variant == PN_VARIANT_HOMO
                             ? a.dT + ((int64_t)a.seq_node[p * a.L] * a.L + a.seq_type[p * a.L]) * H
                             : a.dZ + (int64_t)a.partner[p] * H;
            if (dpv != dp_cur) {
                if (dp_cur) {
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        const int j = lane + 32 * q;
                        if (j < H) atomicAdd(dp_cur + j, dp_acc[q]);
                        dp_acc[q] = 0.f;
                    }
                }
                dp_cur = dpv;
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = lane + 32 * q;
                if (j < H) {
                    const float dhv = inv_nw * att * gq[q] + dscore * a.attw_w[j];
                    a.dhfin[p * H + j] = dhv;
                    amax = fmaxf(amax, fabsf(dhv));
                    dw_h[q] += dscore * hq[q];
                    dw_p[q] += dscore * pv[j];
                    dp_acc[q] += dscore * a.attw_w[H + j];
                }
            }
        }
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = lane + 32 * q;
            if (j < H) {
                atomicAdd(&s_dattw[j], dw_h[q]); atomicAdd(&s_dattw[H + j], dw_p[q]);
                if (dp_cur) atomicAdd(dp_cur + j, dp_acc[q]);
            }
        }
        if (lane == 0) atomicAdd(&s_dattw[2 * H], dbias);
        if (a.absmax_bits) {
            amax = warp_max(amax);
            if (lane == 0 && amax > 0.f) atomicMax(a.absmax_bits, __float_as_uint(amax));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H + 1; i += blockDim.x) atomicAdd(a.dattw + i, s_dattw[i]);
}

// ---------------------------------------------------------------------------
// pooling without attention (SURVEY 8f-4 ablations; MEAN is what baseline/GPRGNN/src/copy.py:350 does):
// pooled[s] = mean / sum / max over the nw final hidden states of row s, then the same concat + dropout.
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pool_simple_fwd(PoolArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= a.S) return;
    const int H = a.H, nw = a.nw;
    const int64_t node = a.rows[s];
    for (int j = lane; j < H; j += 32) {
        float acc = a.pool == PN_POOL_MAX ? -INFINITY : 0.f;
        for (int w = 0; w < nw; w++) {
            const float v = a.hfin[(s * nw + w) * H + j];
            acc = a.pool == PN_POOL_MAX ? fmaxf(acc, v) : acc + v;
        }
        float pooled = a.pool == PN_POOL_MEAN ? acc / (float)nw : acc;
        float ego = a.Z[node * H + j];
        if (a.drop_thr) {
            ego = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + j) < a.drop_thr ? 0.f : ego * a.drop_scale;
            pooled = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + H + j) < a.drop_thr ? 0.f : pooled * a.drop_scale;
        }
        a.cat[s * 2 * H + j] = ego;
        a.cat[s * 2 * H + H + j] = pooled;
    }
}

__global__ void __launch_bounds__(256) k_pool_simple_bwd(PoolArgs a)
{
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    float amax = 0.f;
    if (s < a.S) {
        const int H = a.H, nw = a.nw;
        const int64_t node = a.rows[s];
        for (int j = lane; j < H; j += 32) {
            float g = a.dcat[s * 2 * H + H + j], e = a.dcat[s * 2 * H + j];
            if (a.drop_thr) {
                g = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + H + j) < a.drop_thr ? 0.f : g * a.drop_scale;
                e = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + j) < a.drop_thr ? 0.f : e * a.drop_scale;
            }
            atomicAdd(a.dZ + node * H + j, e);
            int arg = 0;
            if (a.pool == PN_POOL_MAX) {   // gradient goes to the first maximal entry
                float m = -INFINITY;
                for (int w = 0; w < nw; w++) {
                    const float v = a.hfin[(s * nw + w) * H + j];
                    if (v > m) { m = v; arg = w; }
                }
            }
            const float gm = a.pool == PN_POOL_MEAN ? g / (float)nw : g;
            for (int w = 0; w < nw; w++) {
                const float d = a.pool == PN_POOL_MAX ? (w == arg ? g : 0.f) : gm;
                a.dhfin[(s * nw + w) * H + j] = d;
                amax = fmaxf(amax, fabsf(d));
            }
        }
    }
    if (a.absmax_bits) {
        amax = warp_max(amax);
        if (lane == 0 && amax > 0.f) atomicMax(a.absmax_bits, __float_as_uint(amax));
    }
}

// ---------------------------------------------------------------------------
// H = 128 versions of the two pooling kernels: a lane owns 4 CONSECUTIVE hidden units, so every row of hfin / T / Z /
// dhfin moves as one 16-byte access per lane (512 B per warp instruction) instead of four 4-byte ones.  Same math.
// The partner vector is looked up once per run of equal partner rows (homo: the row's own step-0 entry for all its paths).
// ---------------------------------------------------------------------------
__device__ __forceinline__ float dot4(const float4 &a, const float4 &b) { return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w; }
__device__ __forceinline__ float4 ld4(const float *p) { return __ldg(reinterpret_cast<const float4 *>(p)); }

__global__ void __launch_bounds__(256) k_pool_fwd128(PoolArgs a)
{
    constexpr int H = 128;
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s >= a.S) return;
    const int nw = a.nw;
    const float bias = a.attw_b[0];
    const float4 wh = ld4(a.attw_w + 4 * lane), wp = ld4(a.attw_w + H + 4 * lane);
    float mx = -INFINITY;
    const float *pv_last = nullptr;
    float pd_last = 0.f;
    float4 pooled = make_float4(0.f, 0.f, 0.f, 0.f);
    const bool homo = a.variant != PN_VARIANT_HETERO;   // weight = 1 + score: no normaliser, so ONE pass over the h_w rows
#pragma unroll 4
    for (int w = 0; w < nw; w++) {
        const int64_t p = s * nw + w;
        const float4 h4 = ld4(a.hfin + p * H + 4 * lane);
        const float *pv = partner_vec(a, p);
        if (pv != pv_last) { pd_last = dot4(wp, ld4(pv + 4 * lane)); pv_last = pv; }
        const float d = warp_sum(dot4(wh, h4) + pd_last) + bias;
        if (lane == 0) { a.score[p] = d; if (homo) a.att[p] = 1.f + d; }
        mx = fmaxf(mx, d > 0.f ? d : 0.01f * d);
        if (homo) {
            const float at = 1.f + d;
            pooled.x += at * h4.x; pooled.y += at * h4.y; pooled.z += at * h4.z; pooled.w += at * h4.w;
        }
    }
    if (!homo) {
        __syncwarp();
        float denom = 0.f;
        for (int w = lane; w < nw; w += 32) {
            const float d = a.score[s * nw + w];
            denom += expf((d > 0.f ? d : 0.01f * d) - mx);
        }
        denom = warp_sum(denom);
        for (int w = lane; w < nw; w += 32) {
            const int64_t p = s * nw + w;
            const float d = a.score[p];
            a.att[p] = expf((d > 0.f ? d : 0.01f * d) - mx) / denom;
        }
        __syncwarp();
#pragma unroll 8
        for (int w = 0; w < nw; w++) {
            const float at = a.att[s * nw + w];
            const float4 h4 = ld4(a.hfin + (s * nw + w) * H + 4 * lane);
            pooled.x += at * h4.x; pooled.y += at * h4.y; pooled.z += at * h4.z; pooled.w += at * h4.w;
        }
    }
    const float inv_nw = 1.f / (float)nw;
    float po[4] = {pooled.x * inv_nw, pooled.y * inv_nw, pooled.z * inv_nw, pooled.w * inv_nw};
    const float4 e4 = ld4(a.Z + (int64_t)a.rows[s] * H + 4 * lane);
    float eg[4] = {e4.x, e4.y, e4.z, e4.w};
    if (a.drop_thr) {
#pragma unroll
        for (int q = 0; q < 4; q++) {
            const int j = 4 * lane + q;
            eg[q] = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + j) < a.drop_thr ? 0.f : eg[q] * a.drop_scale;
            po[q] = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + H + j) < a.drop_thr ? 0.f : po[q] * a.drop_scale;
        }
    }
    const float4 eg4 = make_float4(eg[0], eg[1], eg[2], eg[3]), po4 = make_float4(po[0], po[1], po[2], po[3]);
    *reinterpret_cast<float4 *>(a.cat + s * 2 * H + 4 * lane) = eg4;
    *reinterpret_cast<float4 *>(a.cat + s * 2 * H + H + 4 * lane) = po4;
    if (a.logits) {   // logits = cat W2^T + b2   (:210 / :277)
        for (int c = 0; c < a.C; c++) {
            const float *wc = a.fc2_w + (int64_t)c * 2 * H;
            const float d = warp_sum(dot4(eg4, ld4(wc + 4 * lane)) + dot4(po4, ld4(wc + H + 4 * lane)));
            if (lane == 0) a.logits[s * a.C + c] = d + a.fc2_b[c];
        }
    }
}

__global__ void __launch_bounds__(256) k_pool_bwd128(PoolArgs a)
{
    constexpr int H = 128;
    extern __shared__ float s_dattw[];  // [2H+1], then (fused classifier) dW2 [C][2H] and db2 [C]
    const int nw = a.nw;
    const int nfuse = a.dlogits ? a.C * (2 * H + 1) : 0;
    float *s_dw2 = s_dattw + 2 * H + 1, *s_db2 = s_dw2 + a.C * 2 * H;
    for (int i = threadIdx.x; i < 2 * H + 1 + nfuse; i += blockDim.x) s_dattw[i] = 0.f;
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    if (s < a.S) {
        const float inv_nw = 1.f / (float)nw;
        const int64_t node = a.rows[s];
        // gradient wrt the pooled half of the concat, dropout applied once
        float4 g4, e4;   // ... and wrt the ego half
        if (a.dlogits) {   // dcat = dY W2, dW2 += dY^T cat, db2 += dY  -- this row's share
            g4 = make_float4(0.f, 0.f, 0.f, 0.f);
            e4 = g4;
            const float4 ce = ld4(a.cat + s * 2 * H + 4 * lane), cp = ld4(a.cat + s * 2 * H + H + 4 * lane);
            for (int c = 0; c < a.C; c++) {
                const float dl = a.dlogits[s * a.C + c];
                const float *wc = a.fc2_w + (int64_t)c * 2 * H;
                const float4 we = ld4(wc + 4 * lane), wq = ld4(wc + H + 4 * lane);
                e4.x += dl * we.x; e4.y += dl * we.y; e4.z += dl * we.z; e4.w += dl * we.w;
                g4.x += dl * wq.x; g4.y += dl * wq.y; g4.z += dl * wq.z; g4.w += dl * wq.w;
                float *dw = s_dw2 + c * 2 * H;
                atomicAdd(dw + 4 * lane + 0, dl * ce.x); atomicAdd(dw + 4 * lane + 1, dl * ce.y);
                atomicAdd(dw + 4 * lane + 2, dl * ce.z); atomicAdd(dw + 4 * lane + 3, dl * ce.w);
                atomicAdd(dw + H + 4 * lane + 0, dl * cp.x); atomicAdd(dw + H + 4 * lane + 1, dl * cp.y);
                atomicAdd(dw + H + 4 * lane + 2, dl * cp.z); atomicAdd(dw + H + 4 * lane + 3, dl * cp.w);
                if (lane == 0) atomicAdd(s_db2 + c, dl);
            }
        } else {
            g4 = ld4(a.dcat + s * 2 * H + H + 4 * lane);
            e4 = ld4(a.dcat + s * 2 * H + 4 * lane);
        }
        if (a.drop_thr) {
            float *gp = &g4.x, *ep = &e4.x;
#pragma unroll
            for (int q = 0; q < 4; q++) {
                const int j = 4 * lane + q;
                gp[q] = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + H + j) < a.drop_thr ? 0.f : gp[q] * a.drop_scale;
                ep[q] = dropout_hash(PN_SEED(a), 1, (uint64_t)(s + a.s_base) * 2 * H + j) < a.drop_thr ? 0.f : ep[q] * a.drop_scale;
            }
        }
        atomicAdd(reinterpret_cast<float4 *>(a.dZ + node * H + 4 * lane), e4);
        if (a.gdrop) *reinterpret_cast<float4 *>(a.gdrop + s * a.gdrop_ld + 4 * lane) = g4;
        // da_w = inv_nw * <dpooled, h_w>, kept by lane (w mod 32) for w < 64, else recomputed.  Only the softmax backward of
        // the hetero variant needs every da_w before the main loop (its dot_sum); homo's weight is 1 + score, so da_w is
        // computed where it is used and h_w is read once.
        const bool two_pass = a.variant == PN_VARIANT_HETERO;
        float da_keep0 = 0.f, da_keep1 = 0.f;
        float dot_sum = 0.f;  // sum_v att_v * da_v (softmax backward), hetero only
        if (two_pass) {
#pragma unroll 4
            for (int w = 0; w < nw; w++) {
                const int64_t p = s * nw + w;
                const float d = warp_sum(dot4(g4, ld4(a.hfin + p * H + 4 * lane))) * inv_nw;
                if ((w & 31) == lane) { if (w < 32) da_keep0 = d; else if (w < 64) da_keep1 = d; }
                dot_sum += a.att[p] * d;
            }
        }
        const float4 wh = ld4(a.attw_w + 4 * lane), wp = ld4(a.attw_w + H + 4 * lane);
        float dbias = 0.f, amax = 0.f;
        float4 dw_h = make_float4(0.f, 0.f, 0.f, 0.f), dw_p = dw_h, dp_acc = dw_h;
        float *dp_cur = nullptr;
        const float *pv_last = nullptr;
        float4 pv4 = dw_h;
#pragma unroll 2
        for (int w = 0; w < nw; w++) {
            const int64_t p = s * nw + w;
            const float4 h4 = ld4(a.hfin + p * H + 4 * lane);
            const float *pv = partner_vec(a, p);
            if (pv != pv_last) { pv4 = ld4(pv + 4 * lane); pv_last = pv; }
            float da;
            if (two_pass && w < 64) da = __shfl_sync(0xffffffffu, w < 32 ? da_keep0 : da_keep1, w & 31);
            else da = warp_sum(dot4(g4, h4)) * inv_nw;
            const float att = a.att[p];
            float dscore;
            if (two_pass) dscore = att * (da - dot_sum) * (a.score[p] > 0.f ? 1.f : 0.01f);
            else dscore = da;  // weight = 1 + score
            dbias += dscore;
            float *dpv = a.variant == PN_VARIANT_HOMO
                             ? a.dT + ((int64_t)a.seq_node[p * a.L] * a.L + a.seq_type[p * a.L]) * H
                             : a.dZ + (int64_t)a.partner[p] * H;
            if (dpv != dp_cur) {
                if (dp_cur) atomicAdd(reinterpret_cast<float4 *>(dp_cur + 4 * lane), dp_acc);
                dp_acc = make_float4(0.f, 0.f, 0.f, 0.f);
                dp_cur = dpv;
            }
            const float k = inv_nw * att;
            const float4 dh = make_float4(k * g4.x + dscore * wh.x, k * g4.y + dscore * wh.y, k * g4.z + dscore * wh.z, k * g4.w + dscore * wh.w);
            if (a.dhfin) *reinterpret_cast<float4 *>(a.dhfin + p * H + 4 * lane) = dh;
            if (a.dh_k && lane == 0) { a.dh_k[p] = k; a.dh_s[p] = dscore; }   // (may alias att / score: both were read above)
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(dh.x), fabsf(dh.y)), fmaxf(fabsf(dh.z), fabsf(dh.w))));
            dw_h.x += dscore * h4.x; dw_h.y += dscore * h4.y; dw_h.z += dscore * h4.z; dw_h.w += dscore * h4.w;
            dw_p.x += dscore * pv4.x; dw_p.y += dscore * pv4.y; dw_p.z += dscore * pv4.z; dw_p.w += dscore * pv4.w;
            dp_acc.x += dscore * wp.x; dp_acc.y += dscore * wp.y; dp_acc.z += dscore * wp.z; dp_acc.w += dscore * wp.w;
        }
        if (dp_cur) atomicAdd(reinterpret_cast<float4 *>(dp_cur + 4 * lane), dp_acc);
        atomicAdd(&s_dattw[4 * lane + 0], dw_h.x); atomicAdd(&s_dattw[4 * lane + 1], dw_h.y);
        atomicAdd(&s_dattw[4 * lane + 2], dw_h.z); atomicAdd(&s_dattw[4 * lane + 3], dw_h.w);
        atomicAdd(&s_dattw[H + 4 * lane + 0], dw_p.x); atomicAdd(&s_dattw[H + 4 * lane + 1], dw_p.y);
        atomicAdd(&s_dattw[H + 4 * lane + 2], dw_p.z); atomicAdd(&s_dattw[H + 4 * lane + 3], dw_p.w);
        if (lane == 0) atomicAdd(&s_dattw[2 * H], dbias);
        if (a.absmax_bits) {
            amax = warp_max(amax);
            if (lane == 0 && amax > 0.f) atomicMax(a.absmax_bits, __float_as_uint(amax));
        }
    }
    __syncthreads();
    for (int i = threadIdx.x; i < 2 * H + 1; i += blockDim.x) atomicAdd(a.dattw + i, s_dattw[i]);
    if (a.dlogits) {
        for (int i = threadIdx.x; i < a.C * 2 * H; i += blockDim.x) atomicAdd(a.dfc2_w + i, s_dw2[i]);
        for (int i = threadIdx.x; i < a.C; i += blockDim.x) atomicAdd(a.dfc2_b + i, s_db2[i]);
    }
}

// ---------------------------------------------------------------------------
// host orchestration
// ---------------------------------------------------------------------------
static int check_cfg(const pn_pagg_cfg *c)
{
    PN_REQUIRE(c != nullptr, "pagg: cfg is NULL");
    PN_REQUIRE(c->N > 0 && c->F > 0 && c->C > 0 && c->S >= 0 && c->nw > 0 && c->L > 0 && c->L <= 255,
               "pagg: bad sizes N=%d F=%d C=%d S=%d nw=%d L=%d", c->N, c->F, c->C, c->S, c->nw, c->L);
    PN_REQUIRE(c->H % 32 == 0 && c->H >= 32 && c->H <= 128,
               "pagg: hidden size %d unsupported (multiples of 32 up to 128)", c->H);
    PN_REQUIRE(c->variant == PN_VARIANT_HETERO || c->variant == PN_VARIANT_HOMO || c->variant == PN_VARIANT_PLAIN, "pagg: bad variant %d", c->variant);
    PN_REQUIRE(c->cell == PN_CELL_LSTM || c->cell == PN_CELL_RNN || c->cell == PN_CELL_GRU, "pagg: bad cell %d", c->cell);
    PN_REQUIRE(c->pool >= PN_POOL_ATTENTION && c->pool <= PN_POOL_MAX, "pagg: bad pool %d", c->pool);
    PN_REQUIRE(!(c->variant == PN_VARIANT_PLAIN && c->pool == PN_POOL_ATTENTION),
               "pagg: the plain variant (no ReLU, natural path order) has no attention of its own: choose pool mean / sum / max");
    if ((c->math_mode == PN_MATH_TC || c->math_mode == PN_MATH_TC3) && c->cell != PN_CELL_LSTM) {
        set_error("pagg: the tensor-core kernels implement the LSTM cell only (cell %d needs math_mode fp32)", c->cell);
        return PN_ERR_UNSUPPORTED;
    }
    PN_REQUIRE(c->math_mode == PN_MATH_FP32 || c->math_mode == PN_MATH_TC || c->math_mode == PN_MATH_TC3, "pagg: bad math_mode %d", c->math_mode);
    PN_REQUIRE(c->dropout_p >= 0.f && c->dropout_p <= 1.f, "pagg: dropout p=%f outside [0,1]", (double)c->dropout_p);
    PN_REQUIRE(c->row_chunk >= 0, "pagg: row_chunk %d < 0", c->row_chunk);
    PN_REQUIRE((int64_t)c->N * c->L < ((int64_t)1 << 31), "pagg: N*L = %lld table rows exceed int32 row offsets", (long long)c->N * c->L);
    return PN_OK;
}

static void drop_params(const pn_pagg_cfg &c, uint32_t &thr, float &scale)
{
    thr = 0; scale = 1.f;
    if (c.training && c.dropout_p > 0.f) {
        thr = dropout_threshold(c.dropout_p);
        scale = c.dropout_p >= 1.f ? 0.f : 1.f / (1.f - c.dropout_p);
    }
}

template <int UPT>
static int launch_lstm_fwd(const LstmArgs &a, cudaStream_t st)
{
    constexpr int H = 32 * UPT;
    const size_t smem = sizeof(float) * (64 * (2 * H + 4) + 16 * 4 * H);
    PN_CUDA(cudaFuncSetAttribute(k_lstm_fwd_f32<UPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof("lstm_fwd", st);
    k_lstm_fwd_f32<UPT><<<(unsigned)ceil_div(a.P, 64), 256, smem, st>>>(a);
    PN_LAUNCH_CHECK();
    return PN_OK;
}
template <int UPT>
static int launch_lstm_bwd(const LstmArgs &a, cudaStream_t st)
{
    constexpr int H = 32 * UPT;
    const size_t smem = sizeof(float) * (64 * (4 * H + 4) + 16 * 2 * H);
    PN_CUDA(cudaFuncSetAttribute(k_lstm_bwd_f32<UPT>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    ProfScope prof("lstm_bwd", st);
    k_lstm_bwd_f32<UPT><<<(unsigned)ceil_div(a.P, 64), 256, smem, st>>>(a);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

int pagg_tc_forward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const int32_t *seq_node,
                         const uint8_t *seq_type, const PaggChunk &ck, cudaStream_t st);   // pagg_tc.cu
int pagg_tc_backward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const int32_t *seq_node,
                          const uint8_t *seq_type, const PaggChunk &ck, bool factored_dh, cudaStream_t st, cudaEvent_t after_bptt = nullptr);  // pagg_tc.cu
int pagg_tc_project(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const float *X, cudaStream_t st);          // pagg_tc.cu
int pagg_tc3_forward_lstm(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const int32_t *seq_node,
                          const uint8_t *seq_type, const PaggChunk &ck, cudaStream_t st);   // pagg_tc3.cu
int launch_split_table(cudaStream_t st, const float *T, float scale, void *hi, void *lo, int64_t n);   // pagg_tc3.cu
int pagg_tc_project_backward(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const float *X,
                             const pn_pagg_params &g, cudaStream_t st);                                                          // pagg_tc.cu

}  // namespace pn

using namespace pn;

extern "C" int pn_pagg_canonicalize(int32_t variant, int32_t N, int32_t S, int32_t nw, int32_t L,
                                    const void *neis, int32_t neis_dtype, const void *layer_type,
                                    int32_t type_dtype, int32_t *seq_node, uint8_t *seq_type, int32_t *partner,
                                    int32_t *status, void *stream)
{
    PN_REQUIRE(variant == PN_VARIANT_HETERO || variant == PN_VARIANT_HOMO || variant == PN_VARIANT_PLAIN, "canonicalize: bad variant %d", variant);
    PN_REQUIRE(N > 0 && S >= 0 && nw > 0 && L > 0 && L <= 255, "canonicalize: bad sizes N=%d S=%d nw=%d L=%d", N, S, nw, L);
    PN_REQUIRE(neis_dtype == PN_IDX_I64 || neis_dtype == PN_IDX_I32, "canonicalize: node ids must be int64 or int32 (dtype code %d)", neis_dtype);
    PN_REQUIRE(type_dtype == PN_IDX_I64 || type_dtype == PN_IDX_I32 || type_dtype == PN_IDX_U8,
               "canonicalize: labels must be int64, int32 or uint8 (dtype code %d)", type_dtype);
    if (S == 0) return PN_OK;
    PN_REQUIRE(neis && layer_type && seq_node && seq_type && partner, "canonicalize: NULL pointer");
    const int64_t total = (int64_t)S * nw * L;
    const unsigned grid = (unsigned)ceil_div(total, 256);
    cudaStream_t st = (cudaStream_t)stream;
#define PN_CANON(NT, TT) \
    k_canonicalize<NT, TT><<<grid, 256, 0, st>>>(variant, N, S, nw, L, (const NT *)neis, (const TT *)layer_type, seq_node, seq_type, partner, status)
    if (neis_dtype == PN_IDX_I64) {
        if (type_dtype == PN_IDX_I64) PN_CANON(int64_t, int64_t);
        else if (type_dtype == PN_IDX_I32) PN_CANON(int64_t, int32_t);
        else PN_CANON(int64_t, uint8_t);
    } else {
        if (type_dtype == PN_IDX_I64) PN_CANON(int32_t, int64_t);
        else if (type_dtype == PN_IDX_I32) PN_CANON(int32_t, int32_t);
        else PN_CANON(int32_t, uint8_t);
    }
#undef PN_CANON
    PN_LAUNCH_CHECK();
    return PN_OK;
}

extern "C" size_t pn_pagg_workspace_bytes(const pn_pagg_cfg *cfg)
{
    if (check_cfg(cfg) != PN_OK) return 0;
    return carve_workspace(*cfg, nullptr).bytes + 256;
}

namespace {

// rows [s0, s0 + Sc) of the call: the chunk's slice of the index tensors
struct ChunkPtrs {
    const int32_t *seq_node, *partner, *rows;
    const uint8_t *seq_type;
};
ChunkPtrs chunk_ptrs(const pn_pagg_cfg &c, const PaggChunk &ck, const int32_t *seq_node, const uint8_t *seq_type,
                     const int32_t *partner, const int32_t *rows)
{
    const int64_t p0 = ck.s0 * c.nw;
    return {seq_node + p0 * c.L, partner + p0, rows + ck.s0, seq_type + p0 * c.L};
}

// K2 of one chunk: gather + dropout + LSTM                         (:194-195 / :264-265)
int lstm_forward_chunk(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const ChunkPtrs &ix,
                       const PaggChunk &ck, cudaStream_t st)
{
    if (c.math_mode == PN_MATH_TC) return pagg_tc_forward_lstm(c, w, prm, ix.seq_node, ix.seq_type, ck, st);
    if (c.math_mode == PN_MATH_TC3) return pagg_tc3_forward_lstm(c, w, prm, ix.seq_node, ix.seq_type, ck, st);
    uint32_t thr; float scale;
    drop_params(c, thr, scale);
    LstmArgs la{};
    la.T = w.T; la.seq_node = ix.seq_node; la.seq_type = ix.seq_type; la.Wp = w.Wp; la.WpT = w.WpT; la.bp = w.bp;
    la.acts = ck.save ? w.acts : nullptr; la.hfin = w.hfin; la.P = ck.Sc * c.nw; la.L = c.L;
    la.drop_thr = thr; la.drop_scale = scale; la.drop_seed = c.dropout_seed; la.seed_dev = c.dropout_seed_dev;
    la.p_base = ck.s0 * c.nw; la.cell = c.cell;
    switch (c.H / 32) {
        case 1: return launch_lstm_fwd<1>(la, st);
        case 2: return launch_lstm_fwd<2>(la, st);
        case 3: return launch_lstm_fwd<3>(la, st);
        default: return launch_lstm_fwd<4>(la, st);
    }
}

PoolArgs pool_args(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const ChunkPtrs &ix, const PaggChunk &ck)
{
    uint32_t thr; float scale;
    drop_params(c, thr, scale);
    PoolArgs pa{};
    pa.variant = c.variant; pa.H = c.H; pa.nw = c.nw; pa.L = c.L; pa.S = ck.Sc;
    pa.Z = w.Z; pa.T = w.T; pa.hfin = w.hfin; pa.seq_node = ix.seq_node; pa.partner = ix.partner; pa.rows = ix.rows;
    pa.seq_type = ix.seq_type; pa.attw_w = prm.attw_w; pa.attw_b = prm.attw_b;
    pa.score = w.score; pa.att = w.att; pa.cat = w.cat;
    pa.drop_thr = thr; pa.drop_scale = scale; pa.drop_seed = c.dropout_seed; pa.seed_dev = c.dropout_seed_dev;
    pa.s_base = ck.s0; pa.pool = c.pool;
    return pa;
}

// the H = 128 pooling kernels use 16-byte accesses: parameter pointers handed in through the C ABI must be aligned for that
bool pool128_ok(const pn_pagg_cfg &c, const pn_pagg_params &prm)
{
    return c.H == 128 && ((reinterpret_cast<uintptr_t>(prm.attw_w) | reinterpret_cast<uintptr_t>(prm.fc2_w)) & 15) == 0;
}

// K3 (+K4) of one chunk: attention + pool + concat + dropout (+ classifier) (:196-210 / :266-277).
// logits == nullptr: recompute pass of the chunked backward (only score / att / cat are needed).
int pool_forward_chunk(const pn_pagg_cfg &c, const PaggWs &w, const pn_pagg_params &prm, const ChunkPtrs &ix,
                       const PaggChunk &ck, float *logits, cudaStream_t st)
{
    PoolArgs pa = pool_args(c, w, prm, ix, ck);
    if (c.pool != PN_POOL_ATTENTION) {
        k_pool_simple_fwd<<<(unsigned)ceil_div(ck.Sc, 8), 256, 0, st>>>(pa);
        PN_LAUNCH_CHECK();
        if (logits == nullptr) return PN_OK;
        return launch_gemm(st, w.cat, 2 * c.H, 1, prm.fc2_w, 1, 2 * c.H, logits, c.C, ck.Sc, c.C, 2 * c.H, prm.fc2_b, false, false);
    }
    const bool pool128 = pool128_ok(c, prm);
    const bool fuse_fc2 = pool128 && c.C <= POOL_FUSE_MAX_C && logits != nullptr;
    if (fuse_fc2) { pa.C = c.C; pa.fc2_w = prm.fc2_w; pa.fc2_b = prm.fc2_b; pa.logits = logits; }
    if (pool128) k_pool_fwd128<<<(unsigned)ceil_div(ck.Sc, 8), 256, 0, st>>>(pa);
    else k_pool_fwd<<<(unsigned)ceil_div(ck.Sc, 8), 256, 0, st>>>(pa);
    PN_LAUNCH_CHECK();
    if (fuse_fc2 || logits == nullptr) return PN_OK;
    return launch_gemm(st, w.cat, 2 * c.H, 1, prm.fc2_w, 1, 2 * c.H, logits, c.C, ck.Sc, c.C, 2 * c.H, prm.fc2_b, false, false);
}

}  // namespace

extern "C" int pn_pagg_forward(const pn_pagg_cfg *cfg, const pn_pagg_params *prm, const float *X,
                               const int32_t *seq_node, const uint8_t *seq_type,
                               const int32_t *partner, const int32_t *rows, void *workspace,
                               float *logits, void *stream)
{
    NvtxScope nvtx("pn_pagg_forward");
    int rc = check_cfg(cfg);
    if (rc != PN_OK) return rc;
    const pn_pagg_cfg &c = *cfg;
    PN_REQUIRE(prm && X && workspace, "pagg_forward: NULL pointer");
    if (c.S == 0) return PN_OK;
    PN_REQUIRE(seq_node && seq_type && partner && rows && logits, "pagg_forward: NULL pointer");
    cudaStream_t st = (cudaStream_t)stream;
    const PaggWs w = carve_workspace(c, (void *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255));
    const int64_t N = c.N, F = c.F, H = c.H, L = c.L;
    const bool homo = c.variant == PN_VARIANT_HOMO;
    const bool tcm = c.math_mode == PN_MATH_TC;

    // K0: Z = act(X W0^T + b0)                                     (:175 / :242-243)
    // K1: T[n, d*H + j] = act(Z nets_w[d]^T + nets_b[d])           (:185-191 / :249-257, hoisted)
    if (tcm || c.math_mode == PN_MATH_TC3) {   // (TC3: the same projections as three fp16 passes per product, fp32 in and out)
        rc = pagg_tc_project(c, w, *prm, X, st);
        if (rc) return rc;
    } else {
        rc = launch_gemm(st, X, F, 1, prm->fc0_w, 1, F, w.Z, H, N, H, F, prm->fc0_b, homo, false);
        if (rc) return rc;
        rc = launch_gemm(st, w.Z, H, 1, prm->nets_w, 1, H, w.T, L * H, N, L * H, H, prm->nets_b, homo, false);
        if (rc) return rc;
    }
    if (c.math_mode == PN_MATH_TC3) {   // split-precision LSTM: it gathers the table as two fp16 terms, times the dropout scale
        uint32_t thr; float tscale;
        drop_params(c, thr, tscale);
        rc = launch_split_table(st, w.T, tscale, w.t16, reinterpret_cast<unsigned char *>(w.t16) + 2 * (N * L * H), N * L * H);
        if (rc) return rc;
    } else if (!tcm) {   // (the tensor-core modes pack their own fp16 images of the weights into the same regions)
        k_pack_lstm<<<(unsigned)ceil_div(2 * H * 4 * H, 256), 256, 0, st>>>((int)H, c.cell, prm->w_ih, prm->w_hh, prm->b_ih,
                                                                          prm->b_hh, w.Wp, w.WpT, w.bp);
        PN_LAUNCH_CHECK();
    }
    const int64_t rows_per = chunk_rows(c);
    const bool single = rows_per >= c.S;
    for (int64_t s0 = 0; s0 < c.S; s0 += rows_per) {
        // a chunked call saves nothing in the forward: its backward recomputes chunk by chunk
        const PaggChunk ck{s0, std::min<int64_t>(rows_per, c.S - s0), single && !c.no_backward, s0 == 0};
        const ChunkPtrs ix = chunk_ptrs(c, ck, seq_node, seq_type, partner, rows);
        rc = lstm_forward_chunk(c, w, *prm, ix, ck, st);
        if (rc) return rc;
        rc = pool_forward_chunk(c, w, *prm, ix, ck, logits + s0 * c.C, st);
        if (rc) return rc;
    }
    return PN_OK;
}

// side stream + fork / join events of the tensor-core backward (one set per device, created on first use)
namespace pn {
struct BwdFork { cudaStream_t side; cudaEvent_t ev_fork, ev_join; };
static int bwd_fork_get(BwdFork &f)
{
    static BwdFork slots[64] = {};
    static bool made[64] = {};
    static std::mutex mu;
    f = BwdFork{};
    const char *e = getenv("PN_BWD_FORK");
    if (e && e[0] == '0') return PN_OK;
    int dev = 0;
    PN_CUDA(cudaGetDevice(&dev));
    if (dev < 0 || dev >= 64) return PN_OK;
    std::lock_guard<std::mutex> lk(mu);
    if (!made[dev]) {
        PN_CUDA(cudaStreamCreateWithFlags(&slots[dev].side, cudaStreamNonBlocking));
        PN_CUDA(cudaEventCreateWithFlags(&slots[dev].ev_fork, cudaEventDisableTiming));
        PN_CUDA(cudaEventCreateWithFlags(&slots[dev].ev_join, cudaEventDisableTiming));
        made[dev] = true;
    }
    f = slots[dev];
    return PN_OK;
}
}  // namespace pn

extern "C" int pn_pagg_backward(const pn_pagg_cfg *cfg, const pn_pagg_params *prm, const float *X,
                                const int32_t *seq_node, const uint8_t *seq_type,
                                const int32_t *partner, const int32_t *rows, void *workspace,
                                const float *dlogits, const pn_pagg_params *grads, void *stream)
{
    NvtxScope nvtx("pn_pagg_backward");
    int rc = check_cfg(cfg);
    if (rc != PN_OK) return rc;
    const pn_pagg_cfg &c = *cfg;
    PN_REQUIRE(prm && X && workspace && grads, "pagg_backward: NULL pointer");
    PN_REQUIRE(!cfg->no_backward, "pagg_backward: the forward ran with no_backward = 1 (nothing was saved)");
    cudaStream_t st = (cudaStream_t)stream;
    const PaggWs w = carve_workspace(c, (void *)(((uintptr_t)workspace + 255) & ~(uintptr_t)255));
    const int64_t N = c.N, F = c.F, H = c.H, L = c.L, S = c.S, C = c.C;
    const bool homo = c.variant == PN_VARIANT_HOMO;
    const bool tcm = c.math_mode == PN_MATH_TC;                    // projections on tensor cores
    const bool tcl = tcm || c.math_mode == PN_MATH_TC3;            // LSTM on tensor cores
    const pn_pagg_params &g = *grads;
#define ZERO(ptr, n) PN_CUDA(cudaMemsetAsync((ptr), 0, sizeof(float) * (size_t)(n), st))
    if (S == 0) {  // no rows: every gradient is zero
        ZERO(g.fc0_w, H * F); ZERO(g.fc0_b, H); ZERO(g.nets_w, L * H * H); ZERO(g.nets_b, L * H);
        const int64_t ng = c.cell == PN_CELL_RNN ? 1 : c.cell == PN_CELL_GRU ? 3 : 4;   // gate blocks of the cell's parameters
        ZERO(g.w_ih, ng * H * H); ZERO(g.w_hh, ng * H * H); ZERO(g.b_ih, ng * H); ZERO(g.b_hh, ng * H);
        ZERO(g.attw_w, 2 * H); ZERO(g.attw_b, 1); ZERO(g.fc2_w, C * 2 * H); ZERO(g.fc2_b, C);
        return PN_OK;
    }
    PN_REQUIRE(seq_node && seq_type && partner && rows && dlogits, "pagg_backward: NULL pointer");
    const bool attention = c.pool == PN_POOL_ATTENTION;
    const bool pool128 = attention && pool128_ok(c, *prm);
    const bool fuse_fc2 = pool128 && C <= POOL_FUSE_MAX_C;   // fc2 backward inside the pooling kernel (it holds every concat row anyway)
    // accumulators shared by all chunks (the chunk kernels add into them atomically)
    ZERO(w.dZ, N * H); ZERO(w.dT, N * L * H); ZERO(w.dattw, 2 * H + 1);
    ZERO(w.dWp, 2 * H * 4 * H); ZERO(w.dbp, 4 * H);
    ZERO(g.fc2_w, C * 2 * H); ZERO(g.fc2_b, C);
    uint32_t thr; float scale;
    drop_params(c, thr, scale);
    const int64_t rows_per = chunk_rows(c);
    const bool single = rows_per >= S;
    // Tensor-core mode: after the last chunk's BPTT kernel the step forks.  The weight-gradient GEMM of the LSTM (0.45 ms on 132
    // SMs, one 230 KB CTA each) and the backward of the two projections (five small kernels, 0.2 ms, that need only dT / dZ) are
    // independent, so the projections run on a side stream -- the column sweeps beside the wgrad CTAs, the GEMMs on the 16 SMs
    // the wgrad clusters leave free -- and join before the call returns.  Works the same under stream capture (the graph gets
    // two branches).  PN_BWD_FORK=0 keeps everything on the caller's stream.
    BwdFork fork{};
    if (tcl) { rc = bwd_fork_get(fork); if (rc) return rc; }
    for (int64_t s0 = 0; s0 < S; s0 += rows_per) {
        const PaggChunk ck{s0, std::min<int64_t>(rows_per, S - s0), true, s0 == 0};
        const ChunkPtrs ix = chunk_ptrs(c, ck, seq_node, seq_type, partner, rows);
        const int64_t Sc = ck.Sc, Pc = Sc * c.nw;
        const float *dl = dlogits + s0 * C;
        if (!single) {   // recompute this chunk's forward (same masks: mask indices are global)
            rc = lstm_forward_chunk(c, w, *prm, ix, ck, st);
            if (rc) return rc;
            rc = pool_forward_chunk(c, w, *prm, ix, ck, nullptr, st);
            if (rc) return rc;
        }
        // ---- fc2 ----
        if (!fuse_fc2) {
            const int splitk = (int)std::max<int64_t>(1, std::min<int64_t>(128, Sc / 128));  // K = Sc on very few output tiles
            rc = launch_gemm(st, dl, 1, C, w.cat, 2 * H, 1, g.fc2_w, 2 * H, C, 2 * H, Sc, nullptr, false, splitk == 1, splitk);  // dW2 += dY^T cat
            if (rc) return rc;
            rc = launch_colsum(st, dl, Sc, C, C, g.fc2_b, /*accumulate=*/true);
            if (rc) return rc;
            rc = launch_gemm(st, dl, C, 1, prm->fc2_w, 2 * H, 1, w.dcat, 2 * H, Sc, 2 * H, C, nullptr, false, false);  // dcat = dY W2
            if (rc) return rc;
        }
        // ---- attention / pooling ----
        PoolArgs pa = pool_args(c, w, *prm, ix, ck);
        pa.dcat = w.dcat; pa.dhfin = w.dhfin; pa.dZ = w.dZ; pa.dT = w.dT; pa.dattw = w.dattw;
        pa.absmax_bits = nullptr;
        if (tcl) {
            ZERO(w.gscale, 2);
            pa.absmax_bits = reinterpret_cast<unsigned int *>(w.gscale);
        }
        if (fuse_fc2) { pa.C = (int)C; pa.fc2_w = prm->fc2_w; pa.dlogits = dl; pa.dfc2_w = g.fc2_w; pa.dfc2_b = g.fc2_b; }
        const bool factored = tcl && pool128;   // k_pool_bwd128 -> k_lstm_bwd_tc: dh_final travels as (k, dscore) per path + one vector per row
        if (factored) {
            pa.dhfin = nullptr;
            pa.dh_k = w.att; pa.dh_s = w.score;          // in place: a path's att / score are read before they are overwritten
            pa.gdrop = w.dcat + H; pa.gdrop_ld = 2 * H;  // in place of the row's pooled-half gradient (unused when fc2 is fused)
        }
        if (!attention) k_pool_simple_bwd<<<(unsigned)ceil_div(Sc, 8), 256, 0, st>>>(pa);
        else if (pool128) k_pool_bwd128<<<(unsigned)ceil_div(Sc, 8), 256, sizeof(float) * (2 * H + 1 + (fuse_fc2 ? C * (2 * H + 1) : 0)), st>>>(pa);
        else k_pool_bwd<<<(unsigned)ceil_div(Sc, 8), 256, sizeof(float) * (2 * H + 1), st>>>(pa);
        PN_LAUNCH_CHECK();
        // ---- LSTM BPTT ----
        if (tcl) {
            rc = pagg_tc_backward_lstm(c, w, *prm, ix.seq_node, ix.seq_type, ck, factored, st, (s0 + rows_per >= S) ? fork.ev_fork : nullptr);
            if (rc) return rc;
        } else {
            LstmArgs la{};
            la.T = w.T; la.seq_node = ix.seq_node; la.seq_type = ix.seq_type; la.Wp = w.Wp; la.WpT = w.WpT; la.bp = w.bp;
            la.acts = w.acts; la.hfin = w.hfin; la.P = Pc; la.L = (int)L;
            la.drop_thr = thr; la.drop_scale = scale; la.drop_seed = c.dropout_seed; la.seed_dev = c.dropout_seed_dev;
            la.p_base = s0 * c.nw; la.cell = c.cell;
            la.dhfin = w.dhfin; la.dT = w.dT; la.dA = w.dA; la.XH = w.XH;
            switch (H / 32) {
                case 1: rc = launch_lstm_bwd<1>(la, st); break;
                case 2: rc = launch_lstm_bwd<2>(la, st); break;
                case 3: rc = launch_lstm_bwd<3>(la, st); break;
                default: rc = launch_lstm_bwd<4>(la, st); break;
            }
            if (rc) return rc;
            // dWp[k][col] += sum_r XH[r][k] dA[r][col];  dbp += colsum(dA)
            const int64_t R = L * Pc;
            const int splitk = (int)std::max<int64_t>(1, std::min<int64_t>(256, R / 2048));
            {
                ProfScope prof("wgrad", st);
                rc = launch_gemm(st, w.XH, 1, 2 * H, w.dA, 4 * H, 1, w.dWp, 4 * H, 2 * H, 4 * H, R, nullptr, false, splitk == 1, splitk);
            }
            if (rc) return rc;
            rc = launch_colsum(st, w.dA, R, 4 * H, 4 * H, w.dbp, /*accumulate=*/true);
            if (rc) return rc;
        }
    }
    if (tcl && fork.ev_fork) {   // the projections' backward on the side stream, from the point where dT / dZ were complete
        PN_CUDA(cudaStreamWaitEvent(fork.side, fork.ev_fork, 0));
        rc = pagg_tc_project_backward(c, w, *prm, X, g, fork.side);
        if (rc) return rc;
        PN_CUDA(cudaEventRecord(fork.ev_join, fork.side));
    }
    PN_CUDA(cudaMemcpyAsync(g.attw_w, w.dattw, sizeof(float) * 2 * H, cudaMemcpyDeviceToDevice, st));
    PN_CUDA(cudaMemcpyAsync(g.attw_b, w.dattw + 2 * H, sizeof(float), cudaMemcpyDeviceToDevice, st));
    k_unpack_lstm_grads<<<(unsigned)ceil_div(4 * H * H, 256), 256, 0, st>>>((int)H, c.cell, w.dWp, w.dbp, g.w_ih, g.w_hh, g.b_ih, g.b_hh);
    PN_LAUNCH_CHECK();
    if (tcl && fork.ev_fork) { PN_CUDA(cudaStreamWaitEvent(st, fork.ev_join, 0)); return PN_OK; }
    if (tcl) return pagg_tc_project_backward(c, w, *prm, X, g, st);
    // ---- table projection: T = act(Z Wn^T + bn) ----
    // homo: ReLU backward fused with the bias gradient (column sums) of the table projection
    if (homo) {
        rc = launch_relu_mask_colsum(st, w.dT, w.T, N, L * H, g.nets_b);
        if (rc) return rc;
    }
    {
        const int splitk = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 1024));
        if (splitk > 1) ZERO(g.nets_w, L * H * H);
        // dWn[m][i] = sum_n dT[n][m] Z[n][i]
        rc = launch_gemm(st, w.dT, 1, L * H, w.Z, H, 1, g.nets_w, H, L * H, H, N, nullptr, false, false, splitk);
        if (rc) return rc;
    }
    if (!homo) {
        rc = launch_colsum(st, w.dT, N, L * H, L * H, g.nets_b);
        if (rc) return rc;
    }
    rc = launch_gemm(st, w.dT, L * H, 1, prm->nets_w, H, 1, w.dZ, H, N, H, L * H, nullptr, false, true);  // dZ += dT Wn
    if (rc) return rc;
    // ---- fc0: Z = act(X W0^T + b0) ----
    if (homo) {
        rc = launch_relu_mask_colsum(st, w.dZ, w.Z, N, H, g.fc0_b);
        if (rc) return rc;
    }
    {
        const int splitk = (int)std::max<int64_t>(1, std::min<int64_t>(64, N / 1024));
        if (splitk > 1) ZERO(g.fc0_w, H * F);
        rc = launch_gemm(st, w.dZ, 1, H, X, F, 1, g.fc0_w, F, H, F, N, nullptr, false, false, splitk);  // dW0 = dZ^T X
        if (rc) return rc;
    }
    if (!homo) rc = launch_colsum(st, w.dZ, N, H, H, g.fc0_b);
#undef ZERO
    return rc;
}
