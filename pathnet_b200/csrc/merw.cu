// MERW path sampler for sm_100a — replaces preprocess/gen_merw.cpp and
// gen_epoch_merw.cpp of the reference (citations are to gen_merw.cpp).
//
//   host  : CSR in file-row order + alias cells, exactly the reference's
//           two-queue fp64 construction (:23-79); each cell's fp64 split S is
//           reduced to an int32 threshold so the device loop is integer-only.
//   K_d1/K_relax : dense hop-distance table D[n][n] (u8) — the reference's
//           `char dis[N][N]` (:10) filled by all-sources bfs() (:101-123),
//           here by in-place min-plus relaxation sweeps over the CSR.
//   K_walk: one thread per chunk of WPT consecutive walks.  The thread's
//           31-word rand() window lives in shared memory; its start offset in
//           the single sequential srand(seed) stream is reached by polynomial
//           jump-ahead done cooperatively by the warp (lfg.h).  Each
//           transition is ONE dependent 32-byte sector: the alias cell carries
//           the (offset, degree) of both candidate next nodes.
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <thread>
#include <vector>

#include "common.h"
#include "lfg.h"

namespace pn {

// one alias cell = one 32 B sector
struct __align__(32) Cell {
    int32_t A, B;      // next node if r2 > thr / otherwise        (:90)
    int32_t thr;       // (1.0*r2/RAND_MAX > S) <=> (r2 > thr)
    uint32_t offA;     // first cell of A's row
    uint32_t degA;     // cells(A)
    uint32_t offB, degB;
    uint32_t pad;
};
static_assert(sizeof(Cell) == 32, "cell must be one sector");

}  // namespace pn

struct pn_graph {
    int device = 0;
    int32_t n = 0;
    int64_t m = 0;
    int32_t max_deg = 0;
    int32_t dead_end = -1;
    bool uniform = false;         // uniform-random-walk variant (preprocess/gen.cpp): one rand() per roll
    // host copies (parity tests, text tools)
    std::vector<int64_t> off;
    std::vector<int32_t> A, B, thr;
    std::vector<double> S;
    // device
    uint2 *d_noderec = nullptr;   // [n] (off, deg)
    pn::Cell *d_cells = nullptr;  // [m]
    uint32_t *d_rowptr = nullptr; // [n+1] for the distance-table build
    int32_t *d_adj = nullptr;     // [m] E[u] in file-row order
    uint8_t *d_dist = nullptr;    // [n][n] hop distance, saturated
    int *d_overflow = nullptr;    // sparse-label kernel's ball-capacity flag
    int32_t dist_L = 0;           // walk length the table is valid for
};

namespace pn {

// largest r in [-1, RAND_MAX] with (double)r / RAND_MAX <= S  (-1: none)
static int32_t split_to_threshold(double S)
{
    const double RM = 2147483647.0;
    if (std::isnan(S)) return 2147483647;  // p > NaN is false -> always B
    if (S < 0.0) return -1;
    if (S >= 1.0) return 2147483647;
    int64_t g = (int64_t)std::floor(S * RM);
    if (g < 0) g = 0;
    if (g > 2147483647) g = 2147483647;
    while (g < 2147483647 && (double)(g + 1) / RM <= S) g++;
    while (g >= 0 && (double)g / RM > S) g--;
    return (int32_t)g;
}

// AliasTable::init (:23-79) for one node; writes exactly deg cells.
static void alias_build_node(int deg, const int32_t *ids, const double *p_in, int32_t *A,
                             int32_t *B, double *S, std::vector<int32_t> &qa_id,
                             std::vector<double> &qa_p, std::vector<int32_t> &qb_id,
                             std::vector<double> &qb_p)
{
    qa_id.clear(); qa_p.clear(); qb_id.clear(); qb_p.clear();
    size_t ah = 0, bh = 0;
    int64_t cells = 0;
    for (int i = 0; i < deg; i++) {
        double p = p_in[i] * deg;
        if (p > 1.0) { qa_id.push_back(ids[i]); qa_p.push_back(p); }
        else         { qb_id.push_back(ids[i]); qb_p.push_back(p); }
    }
    while (ah < qa_id.size() && bh < qb_id.size()) {
        int32_t idA = qa_id[ah]; double probA = qa_p[ah]; ah++;
        int32_t idB = qb_id[bh]; double probB = qb_p[bh]; bh++;
        A[cells] = idA; B[cells] = idB; S[cells] = probB; cells++;
        double res = probA - (1.0 - probB);
        if (std::fabs(res - 1.0) < 1e-5) {
            A[cells] = idA; B[cells] = idA; S[cells] = res; cells++;
            continue;
        }
        if (res > 1.0) { qa_id.push_back(idA); qa_p.push_back(res); }
        else           { qb_id.push_back(idA); qb_p.push_back(res); }
    }
    for (; ah < qa_id.size(); ah++) { A[cells] = B[cells] = qa_id[ah]; S[cells] = 1.0; cells++; }
    for (; bh < qb_id.size(); bh++) { A[cells] = B[cells] = qb_id[bh]; S[cells] = 1.0; cells++; }
}

// ----------------------------------------------------------------------------
// distance table
// ----------------------------------------------------------------------------

// D[u][u] = 0, D[u][v] = 1 for every row (u,v), v != u; everything else 255.
__global__ void k_dist_level1(uint8_t *__restrict__ D, const uint32_t *__restrict__ rowptr,
                              const int32_t *__restrict__ adj, int32_t n)
{
    int64_t u = blockIdx.x;
    uint8_t *row = D + u * (int64_t)n;
    for (uint32_t e = rowptr[u] + threadIdx.x; e < rowptr[u + 1]; e += blockDim.x) {
        int32_t v = adj[e];
        if (v != u) row[v] = 1;
    }
    if (threadIdx.x == 0) row[u] = 0;
}

// One in-place sweep  D[u][x] = min(D[u][x], 1 + min_{v in E[u]} D[v][x]).
// After level-1 init and k-1 sweeps every entry with true distance <= k is
// exact and every other entry is >= its true distance (a real path length or
// 255), which is all the label needs (see pn_graph_prepare_labels).
// grid: (n, chunks of 16*blockDim bytes)
__global__ void __launch_bounds__(256)
k_dist_relax(uint8_t *D, const uint32_t *__restrict__ rowptr, const int32_t *__restrict__ adj,
             int32_t n, int64_t pitch16 /* n/16 rounded up, in uint4 */)
{
    const int64_t u = blockIdx.x;
    const int64_t c = (int64_t)blockIdx.y * blockDim.x + threadIdx.x;  // uint4 index in row
    if (c >= pitch16) return;
    const uint4 *Dv = reinterpret_cast<const uint4 *>(D);
    uint4 best = make_uint4(~0u, ~0u, ~0u, ~0u);
    const uint32_t e0 = rowptr[u], e1 = rowptr[u + 1];
    for (uint32_t e = e0; e < e1; e++) {
        const int64_t v = adj[e];
        uint4 d = Dv[v * pitch16 + c];
        best.x = __vminu4(best.x, d.x); best.y = __vminu4(best.y, d.y);
        best.z = __vminu4(best.z, d.z); best.w = __vminu4(best.w, d.w);
    }
    const uint32_t one = 0x01010101u;
    uint4 own = Dv[u * pitch16 + c];
    own.x = __vminu4(own.x, __vaddus4(best.x, one)); own.y = __vminu4(own.y, __vaddus4(best.y, one));
    own.z = __vminu4(own.z, __vaddus4(best.z, one)); own.w = __vminu4(own.w, __vaddus4(best.w, one));
    reinterpret_cast<uint4 *>(D)[u * pitch16 + c] = own;
}

// ----------------------------------------------------------------------------
// walk kernel
// ----------------------------------------------------------------------------

constexpr int WALK_THREADS = 256;
constexpr int NPOLY_BITS = 32;
constexpr int NDIGITS = 8;  // radix-16 digits of the chunk index

struct WalkParams {
    const uint2 *noderec;
    const Cell *cells;
    const uint8_t *dist;     // may be null (no labels)
    int64_t dist_pitch;      // bytes per row of dist
    int32_t *nodes;
    uint8_t *types;
    int32_t n, nw, L, wpt;
    int32_t node_begin;
    int32_t dpr;             // rand() draws per roll: 2 (MERW alias roll, gen_merw.cpp:88-89) or 1 (uniform, gen.cpp:112-113)
    int64_t wpe;             // walks per epoch in this call = n' * nw
    int64_t cpe;             // chunks per epoch = ceil(wpe / wpt)
    int64_t total_chunks;    // E * cpe
    // root-centric (sparse label) kernel only
    const uint32_t *rowptr;  // [n+1]
    const int32_t *adj;      // [m]
    int64_t epochs;          // E
    int *overflow;           // set to 1 when a start node's radius-(L-2) ball exceeds the shared-memory table
    // rng_mode = PN_RNG_PHILOX: walk g = ((epoch_begin + e) * n + st) * nw + i draws from its own counter stream
    int32_t rng;             // a pn_rng_mode value
    uint32_t key0, key1;     // lo / hi word of the 64-bit seed
    int64_t epoch_begin;
    // jump-ahead tables (host-computed per call); chunk stride = 2L*wpt draws
    uint32_t base[lfg::DEG];               // window whose first generated word is the
                                           // first rand() of (epoch_begin, node_begin, walk 0)
    uint32_t jstep[lfg::DEG];              // x^(chunk stride)      : next chunk, same epoch
    uint32_t jwrap[lfg::DEG];              // x^(epoch stride - (cpe-1)*chunk stride) : first chunk of next epoch
    uint32_t jlane[32][lfg::DEG];          // x^(chunk stride * l), l = 0..31 : lane l's chunk from the warp's base
    uint32_t jchunk16[NDIGITS][15][lfg::DEG];  // x^(chunk stride * d * 16^i), d = 1..15
    uint32_t jepoch[NPOLY_BITS][lfg::DEG]; // x^(2L*n*nw * 2^b)
};

// warp-cooperative window_apply: w (lane j holds s_j, j < 31) <- window at +k,
// poly in global/param space.  scratch: 64 words of shared memory per warp.
__device__ __forceinline__ uint32_t warp_window_apply(const uint32_t *poly, uint32_t w,
                                                      uint32_t *scratch, int lane)
{
    if (lane < lfg::DEG) scratch[lane] = w;
    __syncwarp();
    // e_{31+d} = e_d + e_{28+d}, d = 0..29: three independent words per round
    for (int d0 = 0; d0 < 30; d0 += 3) {
        if (lane < 3) scratch[31 + d0 + lane] = scratch[d0 + lane] + scratch[28 + d0 + lane];
        __syncwarp();
    }
    uint32_t acc = 0;
    if (lane < lfg::DEG) {
#pragma unroll
        for (int i = 0; i < lfg::DEG; i++) acc += poly[i] * scratch[i + lane];
    }
    __syncwarp();
    return acc;
}

__global__ void __launch_bounds__(WALK_THREADS, 4)
k_merw_walk(const __grid_constant__ WalkParams P)
{
    __shared__ uint32_t ring[lfg::DEG][WALK_THREADS];  // thread-private windows, bank = tid
    __shared__ uint32_t scratch[WALK_THREADS / 32][64];
    __shared__ uint32_t lane_poly[32][lfg::DEG];       // stride 31 words: conflict-free across lanes
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int i = tid; i < 32 * lfg::DEG; i += WALK_THREADS) lane_poly[i / lfg::DEG][i % lfg::DEG] = P.jlane[i / lfg::DEG][i % lfg::DEG];
    __syncthreads();
    const int64_t q_warp = ((int64_t)blockIdx.x * WALK_THREADS + warp * 32);
    if (q_warp >= P.total_chunks) return;  // whole warp idle

    // ---- warp: window of lane 0's chunk (radix-16 digits of c, bits of e) ----
    int64_t e = q_warp / P.cpe, c = q_warp % P.cpe;
    uint32_t w = lane < lfg::DEG ? P.base[lane] : 0u;
    {
        uint64_t cb = (uint64_t)c, eb = (uint64_t)e;
        for (int i = 0; cb; i++, cb >>= 4)
            if (cb & 15) w = warp_window_apply(P.jchunk16[i][(cb & 15) - 1], w, scratch[warp], lane);
        for (int b = 0; eb; b++, eb >>= 1)
            if (eb & 1) w = warp_window_apply(P.jepoch[b], w, scratch[warp], lane);
    }
    int64_t my_e = e, my_c = c + lane;
    if (c + 31 < P.cpe) {
        // ---- fast path: the warp's 32 chunks lie in one epoch; every lane applies its
        //      own polynomial x^(lane * chunk stride) to the warp's window, in parallel ----
        uint32_t *ext = scratch[warp];
        if (lane < lfg::DEG) ext[lane] = w;
        __syncwarp();
        for (int d0 = 0; d0 < 30; d0 += 3) {
            if (lane < 3) ext[31 + d0 + lane] = ext[d0 + lane] + ext[28 + d0 + lane];
            __syncwarp();
        }
        // four passes of 8 coefficients: the 38-word slice of the extended window a pass needs is read
        // once (warp-broadcast) into registers, so each multiply-add runs on registers only
#pragma unroll 1
        for (int pass = 0; pass < 4; pass++) {
            const int i0 = pass * 8;
            uint32_t cf[8], win[38];
#pragma unroll
            for (int i = 0; i < 8; i++) cf[i] = (i0 + i < lfg::DEG) ? lane_poly[lane][i0 + i] : 0u;
#pragma unroll
            for (int k = 0; k < 38; k++) win[k] = ext[(i0 + k) & 63];   // ext[61..63] only meet zero coefficients
#pragma unroll
            for (int j = 0; j < lfg::DEG; j++) {
                uint32_t acc = pass ? ring[j][tid] : 0u;
#pragma unroll
                for (int i = 0; i < 8; i++) acc += cf[i] * win[i + j];
                ring[j][tid] = acc;
            }
        }
    } else {
        // ---- slow path (warp straddles an epoch boundary or the end): step lane by lane ----
        for (int l = 0; l < 32; l++) {
            if (lane < lfg::DEG) ring[lane][warp * 32 + l] = w;
            if (lane == l) { my_e = e; my_c = c; }
            if (l == 31) break;
            if (c + 1 < P.cpe) { w = warp_window_apply(P.jstep, w, scratch[warp], lane); c++; }
            else               { w = warp_window_apply(P.jwrap, w, scratch[warp], lane); c = 0; e++; }
        }
    }
    __syncwarp();
    if (q_warp + lane >= P.total_chunks) return;

    // ---- thread: its chunk of walks ----------------------------------------
    // The 31-word window moves to registers.  All lanes consume draws in lockstep (2L per walk,
    // whatever the walk does), so a refill of 31 draws is fully unrolled with static register
    // indices and every draw is handed to a warp-uniform stepper: even draws are r1 of a roll
    // (the cell index), odd draws are r2 (the alias decision) and perform the roll.
    uint32_t s[lfg::DEG];
#pragma unroll
    for (int j = 0; j < lfg::DEG; j++) s[j] = ring[j][tid];
    const int L = P.L;
    int64_t wi = my_c * P.wpt;
    const int64_t w_end = min(wi + (int64_t)P.wpt, P.wpe);
    const int nwalks = (int)(w_end - wi);                  // this lane's walks (the last chunk of an epoch may be short)
    int32_t *out_n = P.nodes + (my_e * P.wpe + wi) * L;
    uint8_t *out_t = P.types ? P.types + (my_e * P.wpe + wi) * L : nullptr;
    int32_t st = P.node_begin + (int32_t)(wi / P.nw);
    int32_t k_in_st = (int32_t)(wi % P.nw);
    // walk state
    int walk = 0, t = 0;          // walk index within the chunk, position within the walk
    uint32_t r1 = 0;
    int32_t u = st;
    uint2 rec = __ldg(&P.noderec[st]);                     // (off, deg) of the start node, kept while st is unchanged
    uint32_t off = rec.x, deg = rec.y;
    const bool vec4 = (L & 3) == 0;                        // walk outputs are 16-byte (ids) / 4-byte (labels) aligned groups
    int32_t pk_n[4] = {0, 0, 0, 0};
    uint32_t pk_t = 0;
    const int total_draws = P.wpt * P.dpr * L;             // uniform across the warp
    const bool one_draw = P.dpr == 1;
    int drawn = 0;
    bool second = false;                                   // false: next draw is r1, true: r2
    while (drawn < total_draws) {
#pragma unroll
        for (int j = 0; j < lfg::DEG; j++) {
            s[j] += s[(j + 28) % lfg::DEG];                // s_{k+31} = s_k + s_{k+28}
            if (drawn < total_draws) {                     // uniform
                const uint32_t o = s[j] >> 1;
                drawn++;
                if (!second && !one_draw) { r1 = o; second = true; }
                else {
                    if (one_draw) r1 = o;                  // gen.cpp: the single draw picks the neighbour (cells have A == B)
                    second = false;
                    const bool active = walk < nwalks;
                    if (active) {
                        uint32_t lab = 0;
                        if (out_t) {
                            if (t == 1) lab = (u != st);
                            else if (t >= 2) {
                                const uint32_t d = __ldg(P.dist + (int64_t)st * P.dist_pitch + u);
                                lab = d < (uint32_t)t ? d : (uint32_t)t;
                            }
                        }
                        if (vec4) {                        // four positions per 16-byte / 4-byte store
                            const int q = t & 3;
                            pk_n[0] = q == 0 ? u : pk_n[0]; pk_n[1] = q == 1 ? u : pk_n[1];
                            pk_n[2] = q == 2 ? u : pk_n[2]; pk_n[3] = q == 3 ? u : pk_n[3];
                            pk_t = q == 0 ? lab : (pk_t | (lab << (8 * q)));
                            if (q == 3) {
                                *reinterpret_cast<int4 *>(out_n + t - 3) = make_int4(pk_n[0], pk_n[1], pk_n[2], pk_n[3]);
                                if (out_t) *reinterpret_cast<uint32_t *>(out_t + t - 3) = pk_t;
                            }
                        } else {
                            out_n[t] = u;
                            if (out_t) out_t[t] = (uint8_t)lab;
                        }
                        if (t + 1 < L) {                   // the L-th roll's result is unused (:190-197)
                            const uint32_t x = r1 % deg;
                            uint32_t c0, c1, c2, c3, c4, c5, c6, c7;   // one 32-byte sector = one 256-bit load
                            asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                         : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7)
                                         : "l"(P.cells + (off + x)));
                            const bool takeA = (int32_t)o > (int32_t)c2;       // (1.0*r2/RAND_MAX > S) (:88-90)
                            u = takeA ? (int32_t)c0 : (int32_t)c1;
                            off = takeA ? c3 : c5;
                            deg = takeA ? c4 : c6;
                        }
                    }
                    if (++t == L) {                        // next walk (uniform)
                        t = 0;
                        walk++;
                        out_n += L;
                        if (out_t) out_t += L;
                        if (++k_in_st == P.nw) {           // next start node
                            k_in_st = 0; st++;
                            if (walk < nwalks) rec = __ldg(&P.noderec[st]);
                        }
                        off = rec.x; deg = rec.y;
                        u = st;
                    }
                }
            }
        }
    }
}


// ----------------------------------------------------------------------------
// root-centric walk kernel with sparse labels: for graphs whose dense n x n
// distance table does not fit.  CTA = one start node:
//   1. truncated BFS of radius L-2 from the start node into a shared-memory hash
//      set (key = node, value = hop distance) — the ball the reference reads out
//      of dis[st][*] (gen_merw.cpp:101-123, :193);
//   2. the node's nw walks for every epoch: lane = walk, its 2L draws are the
//      first 2L words of the window at the walk's stream offset, each word one
//      31-term dot product of the lane's jump polynomial with the warp's window;
//   3. labels by hash lookup: found -> distance, not found -> position (= L-1).
// ----------------------------------------------------------------------------
// ----------------------------------------------------------------------------
// rng_mode = PN_RNG_PHILOX (SURVEY 8b lists it beside glibc_compat; the reference itself has only rand()):
// Philox4x32-10 (Salmon et al., SC'11), one counter stream per walk -- counter {lo32(g), hi32(g), block, 0}, key = seed;
// roll t takes words 2(t & 1), 2(t & 1) + 1 of block t / 2, each >> 1 (31 bits like rand()); the reference's wasted
// L-th roll is not drawn.  No jump-ahead, no ordering between walks: any split of epochs / start nodes / walks over
// calls, streams or GPUs reproduces the same walks.  Restated on the CPU in oracle/merw_oracle.c (pinned by the Random123
// known-answer vectors).
// ----------------------------------------------------------------------------
__device__ __forceinline__ void philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0, uint32_t k1, uint32_t (&out)[4])
{
#pragma unroll
    for (int r = 0; r < 10; r++) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c0), lo0 = 0xD2511F53u * c0;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c2), lo1 = 0xCD9E8D57u * c2;
        const uint32_t n0 = hi1 ^ c1 ^ k0, n2 = hi0 ^ c3 ^ k1;
        c0 = n0; c1 = lo1; c2 = n2; c3 = lo0;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
    out[0] = c0; out[1] = c1; out[2] = c2; out[3] = c3;
}

// one thread per walk; labels from the dense distance table (P.dist) or none
__global__ void __launch_bounds__(256)
k_merw_walk_philox(const __grid_constant__ WalkParams P)
{
    const int64_t w = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;     // walk index within this call
    if (w >= P.epochs * P.wpe) return;
    const int L = P.L;
    const int64_t e = w / P.wpe, r = w % P.wpe;
    const int32_t st = P.node_begin + (int32_t)(r / P.nw);
    const uint64_t g = ((uint64_t)(P.epoch_begin + e) * (uint64_t)P.n + (uint64_t)st) * (uint64_t)P.nw + (uint64_t)(r % P.nw);
    int32_t *out_n = P.nodes + w * L;
    uint8_t *out_t = P.types ? P.types + w * L : nullptr;
    const uint2 rec = __ldg(&P.noderec[st]);
    uint32_t off = rec.x, deg = rec.y;
    int32_t u = st;
    uint32_t blk[4] = {0u, 0u, 0u, 0u};
    for (int t = 0; t < L; t++) {
        out_n[t] = u;
        if (out_t) {
            uint32_t lab = 0;
            if (t == 1) lab = (u != st);
            else if (t >= 2) {
                const uint32_t d = __ldg(P.dist + (int64_t)st * P.dist_pitch + u);
                lab = d < (uint32_t)t ? d : (uint32_t)t;
            }
            out_t[t] = (uint8_t)lab;
        }
        if (t + 1 == L) break;
        if ((t & 1) == 0) philox4x32_10((uint32_t)g, (uint32_t)(g >> 32), (uint32_t)(t >> 1), 0u, P.key0, P.key1, blk);
        const uint32_t r1 = ((t & 1) ? blk[2] : blk[0]) >> 1, r2 = ((t & 1) ? blk[3] : blk[1]) >> 1;
        const uint32_t x = r1 % deg;
        uint32_t c0, c1, c2, c3, c4, c5, c6, c7;   // one 32-byte sector = one 256-bit load
        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                     : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7)
                     : "l"(P.cells + (off + x)));
        const bool takeA = (int32_t)r2 > (int32_t)c2;
        u = takeA ? (int32_t)c0 : (int32_t)c1;
        off = takeA ? c3 : c5;
        deg = takeA ? c4 : c6;
    }
}

constexpr int ROOT_THREADS = 128;
constexpr int ROOT_HASH = 8192;    // slots (power of two); ball capacity = ROOT_QCAP
constexpr int ROOT_QCAP = 4096;
constexpr int ROOT_SMEM = (ROOT_HASH + ROOT_QCAP + (ROOT_THREADS / 32) * 64 + 32 * lfg::DEG) * 4;

__device__ __forceinline__ uint32_t hash_node(uint32_t v) { return (v * 0x9e3779b1u) >> 19; }  // 13 bits

__global__ void __launch_bounds__(ROOT_THREADS)
k_merw_walk_root(const __grid_constant__ WalkParams P)
{
    extern __shared__ __align__(16) uint32_t root_smem[];
    uint32_t *htab = root_smem;                                            // [ROOT_HASH] 0 = empty, else (node + 1) | dist << 28
    int32_t *queue = reinterpret_cast<int32_t *>(htab + ROOT_HASH);        // [ROOT_QCAP]
    uint32_t(*scratch)[64] = reinterpret_cast<uint32_t(*)[64]>(queue + ROOT_QCAP);                    // [warps][64]
    uint32_t(*lane_poly)[lfg::DEG] = reinterpret_cast<uint32_t(*)[lfg::DEG]>(scratch + ROOT_THREADS / 32);   // [32][31]
    __shared__ int q_tail, q_overflow;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int32_t st = P.node_begin + blockIdx.x;
    const int L = P.L, R = L - 2;
    for (int i = tid; i < ROOT_HASH; i += ROOT_THREADS) htab[i] = 0;
    for (int i = tid; i < 32 * lfg::DEG; i += ROOT_THREADS) lane_poly[i / lfg::DEG][i % lfg::DEG] = P.jlane[i / lfg::DEG][i % lfg::DEG];
    if (tid == 0) { q_tail = 1; q_overflow = 0; queue[0] = st; }
    __syncthreads();
    if (tid == 0) htab[hash_node((uint32_t)st)] = ((uint32_t)st + 1u);   // distance 0
    __syncthreads();
    // ---- 1. BFS, one level per iteration, one warp per frontier node ----
    if (P.types) {
        int lvl_begin = 0;
        for (int level = 0; level < R; level++) {
            const int lvl_end = q_tail;
            __syncthreads();
            for (int qi = lvl_begin + warp; qi < lvl_end && qi < ROOT_QCAP; qi += ROOT_THREADS / 32) {
                const int32_t v = queue[qi];
                const uint32_t e1 = P.rowptr[v + 1];
                for (uint32_t e = P.rowptr[v] + lane; e < e1; e += 32) {
                    const uint32_t w = (uint32_t)P.adj[e];
                    const uint32_t val = (w + 1u) | ((uint32_t)(level + 1) << 28);
                    uint32_t slot = hash_node(w);
                    for (int probe = 0; probe < ROOT_HASH; probe++) {
                        const uint32_t cur = htab[slot];
                        if ((cur & 0x0FFFFFFFu) == w + 1u) break;             // already known (at a smaller or equal distance)
                        if (cur == 0) {
                            const uint32_t old = atomicCAS(&htab[slot], 0u, val);
                            if (old == 0) {                                    // newly discovered
                                const int pos = atomicAdd(&q_tail, 1);
                                if (pos < ROOT_QCAP) queue[pos] = (int32_t)w; else q_overflow = 1;
                                break;
                            }
                            if ((old & 0x0FFFFFFFu) == w + 1u) break;
                        }
                        slot = (slot + 1) & (ROOT_HASH - 1);
                    }
                }
            }
            __syncthreads();
            lvl_begin = lvl_end;
        }
        if (q_overflow) { if (tid == 0) *P.overflow = 1; return; }
    }
    // ---- 2./3. walks: group g = 32 consecutive walks of this start node, lane = walk ----
    const uint2 rec0 = __ldg(&P.noderec[st]);
    const int64_t c_root = (int64_t)(st - P.node_begin) * P.nw;          // chunk (= walk) index of the node's first walk
    const bool philox = P.rng == PN_RNG_PHILOX;
    for (int g = warp; g * 32 < P.nw; g += ROOT_THREADS / 32) {
        // warp window at (epoch 0, walk c_root + 32 g)
        uint32_t w = lane < lfg::DEG ? P.base[lane] : 0u;
        if (!philox) {
            uint64_t cb = (uint64_t)(c_root + g * 32);
            for (int i = 0; cb; i++, cb >>= 4)
                if (cb & 15) w = warp_window_apply(P.jchunk16[i][(cb & 15) - 1], w, scratch[warp], lane);
        }
        const int wi = g * 32 + lane;                                      // walk index within the start node
        const bool active = wi < P.nw;
        uint32_t cf[lfg::DEG];
#pragma unroll
        for (int i = 0; i < lfg::DEG; i++) cf[i] = lane_poly[lane][i];
        for (int64_t ep = 0; ep < P.epochs; ep++) {
            if (ep > 0 && !philox) w = warp_window_apply(P.jepoch[0], w, scratch[warp], lane);   // + one epoch
            // extended window e_0..e_60 in shared memory
            uint32_t *ext = scratch[warp];
            __syncwarp();
            if (lane < lfg::DEG) ext[lane] = w;
            __syncwarp();
            for (int d0 = 0; d0 < 30 && !philox; d0 += 3) {
                if (lane < 3) ext[31 + d0 + lane] = ext[d0 + lane] + ext[28 + d0 + lane];
                __syncwarp();
            }
            const uint64_t gw = ((uint64_t)(P.epoch_begin + ep) * (uint64_t)P.n + (uint64_t)st) * (uint64_t)P.nw + (uint64_t)wi;
            uint32_t blk[4] = {0u, 0u, 0u, 0u};
            int32_t *out_n = P.nodes + ((ep * P.wpe) + c_root + wi) * L;
            uint8_t *out_t = P.types ? P.types + ((ep * P.wpe) + c_root + wi) * L : nullptr;
            int32_t u = st;
            uint32_t off = rec0.x, deg = rec0.y;
            for (int t = 0; t < L; t++) {
                // the walk's draws 2t and 2t+1: words 2t, 2t+1 of the lane's window = dot(cf, ext[i .. i+30])
                uint32_t r1 = 0, r2 = 0;
                if (philox) {
                    if ((t & 1) == 0) philox4x32_10((uint32_t)gw, (uint32_t)(gw >> 32), (uint32_t)(t >> 1), 0u, P.key0, P.key1, blk);
                    r1 = (t & 1) ? blk[2] : blk[0]; r2 = (t & 1) ? blk[3] : blk[1];
                } else {
                    const int i1 = P.dpr * t, i2 = P.dpr == 2 ? i1 + 1 : i1;
#pragma unroll
                    for (int j = 0; j < lfg::DEG; j++) { r1 += cf[j] * ext[i1 + j]; r2 += cf[j] * ext[i2 + j]; }
                }
                r1 >>= 1; r2 >>= 1;
                if (active) {
                    out_n[t] = u;
                    if (out_t) {
                        uint32_t lab = 0;
                        if (t >= 1 && u != st) {
                            lab = (uint32_t)t;                             // not in the ball: distance = position
                            uint32_t slot = hash_node((uint32_t)u);
                            for (int probe = 0; probe < ROOT_HASH; probe++) {
                                const uint32_t cur = htab[slot];
                                if (cur == 0) break;
                                if ((cur & 0x0FFFFFFFu) == (uint32_t)u + 1u) { lab = min(lab, cur >> 28); break; }
                                slot = (slot + 1) & (ROOT_HASH - 1);
                            }
                        }
                        out_t[t] = (uint8_t)lab;
                    }
                    if (t + 1 < L) {
                        const uint32_t x = r1 % deg;
                        uint32_t c0, c1, c2, c3, c4, c5, c6, c7;
                        asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                                     : "=r"(c0), "=r"(c1), "=r"(c2), "=r"(c3), "=r"(c4), "=r"(c5), "=r"(c6), "=r"(c7)
                                     : "l"(P.cells + (off + x)));
                        const bool takeA = (int32_t)r2 > (int32_t)c2;
                        u = takeA ? (int32_t)c0 : (int32_t)c1;
                        off = takeA ? c3 : c5;
                        deg = takeA ? c4 : c6;
                    }
                }
            }
            __syncwarp();
        }
    }
}

}  // namespace pn

// ============================================================================
// C ABI
// ============================================================================
using namespace pn;

extern "C" int pn_graph_from_edges(int32_t n, int64_t m, const int32_t *u, const int32_t *v,
                                   const double *p, int device, pn_graph **out)
{
    PN_REQUIRE(out != nullptr, "pn_graph_from_edges: out is NULL");
    *out = nullptr;
    PN_REQUIRE(n > 0 && m >= 0, "pn_graph_from_edges: need n > 0, m >= 0 (n=%d m=%lld)", n, (long long)m);
    PN_REQUIRE(m < (int64_t)0xFFFFFFF0u, "pn_graph_from_edges: more than 2^32 rows");
    PN_REQUIRE(m == 0 || (u && v && p), "pn_graph_from_edges: NULL edge arrays");
    for (int64_t e = 0; e < m; e++)
        PN_REQUIRE(u[e] >= 0 && u[e] < n && v[e] >= 0 && v[e] < n,
                   "pn_graph_from_edges: row %lld (%d,%d) outside [0,%d)", (long long)e, u[e], v[e], n);
    pn_graph *g = new pn_graph();
    g->device = device; g->n = n; g->m = m;
    g->off.assign((size_t)n + 1, 0);
    for (int64_t e = 0; e < m; e++) g->off[u[e] + 1]++;
    for (int32_t i = 0; i < n; i++) {
        int64_t d = g->off[i + 1];
        if (d > g->max_deg) g->max_deg = (int32_t)d;
        if (d == 0 && g->dead_end < 0) g->dead_end = i;
        g->off[i + 1] += g->off[i];
    }
    std::vector<int32_t> adj((size_t)m);
    std::vector<double> pr((size_t)m);
    {
        std::vector<int64_t> cur(g->off.begin(), g->off.end() - 1);
        for (int64_t e = 0; e < m; e++) {  // stable: file row order per node (:95-99)
            int64_t k = cur[u[e]]++;
            adj[k] = v[e]; pr[k] = p[e];
        }
    }
    g->A.resize(m); g->B.resize(m); g->thr.resize(m); g->S.resize(m);
    // alias build: sequential per node (queue order matters), nodes independent
    unsigned nthreads = std::max(1u, std::min(16u, std::thread::hardware_concurrency()));
    if (m < 100000) nthreads = 1;
    auto work = [&](int32_t lo, int32_t hi) {
        std::vector<int32_t> qa, qb; std::vector<double> pa, pb;
        for (int32_t i = lo; i < hi; i++) {
            int64_t o = g->off[i]; int deg = (int)(g->off[i + 1] - o);
            alias_build_node(deg, adj.data() + o, pr.data() + o, g->A.data() + o, g->B.data() + o,
                             g->S.data() + o, qa, pa, qb, pb);
            for (int k = 0; k < deg; k++) g->thr[o + k] = split_to_threshold(g->S[o + k]);
        }
    };
    if (nthreads == 1) work(0, n);
    else {
        std::vector<std::thread> th;
        for (unsigned t = 0; t < nthreads; t++)
            th.emplace_back(work, (int32_t)((int64_t)n * t / nthreads), (int32_t)((int64_t)n * (t + 1) / nthreads));
        for (auto &x : th) x.join();
    }
    // device upload
    std::vector<Cell> cells((size_t)m);
    for (int64_t k = 0; k < m; k++) {
        Cell c;
        c.A = g->A[k]; c.B = g->B[k]; c.thr = g->thr[k];
        c.offA = (uint32_t)g->off[c.A]; c.degA = (uint32_t)(g->off[c.A + 1] - g->off[c.A]);
        c.offB = (uint32_t)g->off[c.B]; c.degB = (uint32_t)(g->off[c.B + 1] - g->off[c.B]);
        c.pad = 0;
        cells[k] = c;
    }
    std::vector<uint2> rec((size_t)n);
    std::vector<uint32_t> rowptr((size_t)n + 1);
    for (int32_t i = 0; i < n; i++) rec[i] = make_uint2((uint32_t)g->off[i], (uint32_t)(g->off[i + 1] - g->off[i]));
    for (int32_t i = 0; i <= n; i++) rowptr[i] = (uint32_t)g->off[i];
    auto fail = [&](int rc) { pn_graph_free(g); return rc; };
#define PN_G(expr)                                                                    \
    do {                                                                              \
        cudaError_t _e = (expr);                                                      \
        if (_e != cudaSuccess) {                                                      \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                \
            return fail(PN_ERR_CUDA);                                                 \
        }                                                                             \
    } while (0)
    if (device < 0) { *out = g; return PN_OK; }  // host-only graph (tests, text tools)
    PN_G(cudaSetDevice(device));
    PN_G(cudaMalloc(&g->d_noderec, sizeof(uint2) * (size_t)n));
    PN_G(cudaMalloc(&g->d_cells, sizeof(Cell) * (size_t)std::max<int64_t>(m, 1)));
    PN_G(cudaMalloc(&g->d_rowptr, sizeof(uint32_t) * ((size_t)n + 1)));
    PN_G(cudaMalloc(&g->d_adj, sizeof(int32_t) * (size_t)std::max<int64_t>(m, 1)));
    PN_G(cudaMemcpy(g->d_noderec, rec.data(), sizeof(uint2) * (size_t)n, cudaMemcpyHostToDevice));
    PN_G(cudaMemcpy(g->d_cells, cells.data(), sizeof(Cell) * (size_t)m, cudaMemcpyHostToDevice));
    PN_G(cudaMemcpy(g->d_rowptr, rowptr.data(), sizeof(uint32_t) * ((size_t)n + 1), cudaMemcpyHostToDevice));
    PN_G(cudaMemcpy(g->d_adj, adj.data(), sizeof(int32_t) * (size_t)m, cudaMemcpyHostToDevice));
#undef PN_G
    *out = g;
    return PN_OK;
}

// Uniform-random-walk variant (preprocess/gen.cpp:83-94): adjacency = a self loop per node first, then
// (u,v) and (v,u) for every input row with u != v; a roll is `E[u][rand() % E[u].size()]` (:112-114).
// Implemented as degenerate alias cells (A = B = neighbour) plus one draw per roll.
extern "C" int pn_graph_from_edges_uniform(int32_t n, int64_t m, const int32_t *u, const int32_t *v, int device,
                                           pn_graph **out)
{
    PN_REQUIRE(out != nullptr, "pn_graph_from_edges_uniform: out is NULL");
    *out = nullptr;
    PN_REQUIRE(n > 0 && m >= 0 && (m == 0 || (u && v)), "pn_graph_from_edges_uniform: bad arguments");
    std::vector<std::vector<int32_t>> E((size_t)n);
    for (int32_t i = 0; i < n; i++) E[i].push_back(i);
    for (int64_t e = 0; e < m; e++) {
        PN_REQUIRE(u[e] >= 0 && u[e] < n && v[e] >= 0 && v[e] < n, "pn_graph_from_edges_uniform: row %lld outside [0,%d)", (long long)e, n);
        if (u[e] == v[e]) continue;
        E[u[e]].push_back(v[e]);
        E[v[e]].push_back(u[e]);
    }
    std::vector<int32_t> ru, rv;
    std::vector<double> rp;
    for (int32_t i = 0; i < n; i++)
        for (int32_t w : E[i]) { ru.push_back(i); rv.push_back(w); rp.push_back(0.0); }
    // p = 0 for every row: the two-queue construction then emits each row as the cell (id, id, 1.0)
    // in row order, i.e. A = B = E[u][x] — exactly the table a uniform pick indexes
    int rc = pn_graph_from_edges(n, (int64_t)ru.size(), ru.data(), rv.data(), rp.data(), device, out);
    if (rc == PN_OK) (*out)->uniform = true;
    return rc;
}

extern "C" int pn_graph_free(pn_graph *g)
{
    if (!g) return PN_OK;
    if (g->device < 0) { delete g; return PN_OK; }
    cudaSetDevice(g->device);
    cudaFree(g->d_noderec); cudaFree(g->d_cells); cudaFree(g->d_rowptr); cudaFree(g->d_adj);
    cudaFree(g->d_dist); cudaFree(g->d_overflow);
    delete g;
    return PN_OK;
}

extern "C" int pn_graph_info(const pn_graph *g, int32_t *n, int64_t *m, int32_t *max_deg,
                             int32_t *dead_end)
{
    PN_REQUIRE(g, "pn_graph_info: NULL graph");
    if (n) *n = g->n;
    if (m) *m = g->m;
    if (max_deg) *max_deg = g->max_deg;
    if (dead_end) *dead_end = g->dead_end;
    return PN_OK;
}

extern "C" int pn_graph_cells_host(const pn_graph *g, int64_t *off, int32_t *A, int32_t *B,
                                   double *S, int32_t *thr)
{
    PN_REQUIRE(g, "pn_graph_cells_host: NULL graph");
    if (off) memcpy(off, g->off.data(), sizeof(int64_t) * g->off.size());
    if (A) memcpy(A, g->A.data(), sizeof(int32_t) * g->A.size());
    if (B) memcpy(B, g->B.data(), sizeof(int32_t) * g->B.size());
    if (S) memcpy(S, g->S.data(), sizeof(double) * g->S.size());
    if (thr) memcpy(thr, g->thr.data(), sizeof(int32_t) * g->thr.size());
    return PN_OK;
}

extern "C" int pn_graph_prepare_labels(pn_graph *g, int32_t L, size_t max_bytes, void *stream_)
{
    PN_REQUIRE(g, "pn_graph_prepare_labels: NULL graph");
    PN_REQUIRE(L >= 1 && L <= 64, "pn_graph_prepare_labels: L=%d outside [1,64]", L);
    cudaStream_t stream = (cudaStream_t)stream_;
    // A table built for L' >= L is also valid for L (more sweeps only make more
    // entries exact; the rest stay upper bounds).
    if (g->d_dist && g->dist_L >= L) return PN_OK;
    if (L <= 2) { g->dist_L = std::max(g->dist_L, L); return PN_OK; }  // labels need no table
    PN_CUDA(cudaSetDevice(g->device));
    const int64_t n = g->n;
    const int64_t pitch16 = ceil_div(n, 16);
    const size_t bytes = (size_t)pitch16 * 16 * (size_t)n;
    if (max_bytes == 0) max_bytes = (size_t)24 << 30;
    if (bytes > max_bytes) {
        set_error("pn_graph_prepare_labels: dense distance table needs %zu bytes (> %zu); "
                  "sparse ball tables are not implemented yet", bytes, max_bytes);
        return PN_ERR_UNSUPPORTED;
    }
    if (!g->d_dist) PN_CUDA(cudaMalloc(&g->d_dist, bytes));
    PN_CUDA(cudaMemsetAsync(g->d_dist, 0xFF, bytes, stream));
    // level 1 writes bytes; the row pitch is padded to 16
    k_dist_level1<<<(unsigned)n, 128, 0, stream>>>(g->d_dist, g->d_rowptr, g->d_adj, (int32_t)(pitch16 * 16));
    PN_LAUNCH_CHECK();
    // distances up to L-2 must be exact: level 1 + (L-3) sweeps
    dim3 grid((unsigned)n, (unsigned)ceil_div(pitch16, 256));
    for (int k = 2; k <= L - 2; k++) {
        k_dist_relax<<<grid, 256, 0, stream>>>(g->d_dist, g->d_rowptr, g->d_adj, (int32_t)n, pitch16);
        PN_LAUNCH_CHECK();
    }
    g->dist_L = L;
    return PN_OK;
}

static void copy_poly(uint32_t dst[lfg::DEG], const lfg::Poly &p) { memcpy(dst, p.c, sizeof(p.c)); }

// Chunking + jump-ahead tables of one pn_merw_sample call (host).
static int plan_walk(WalkParams &P, int32_t n, int32_t nw, int32_t L, int64_t epoch_begin,
                     int64_t E, int32_t node_begin, int64_t np, uint64_t seed, bool root_mode = false, int dpr = 2)
{
    memset(&P, 0, sizeof(P));
    const int64_t wpe = np * nw;
    P.n = n; P.nw = nw; P.L = L; P.node_begin = node_begin; P.dpr = dpr;
    P.wpe = wpe;
    // chunk size: enough chunks to fill the machine, at most 64 walks per thread
    // (amortises the jump-ahead, ~45 % of a warp's instructions at 16), at least 4
    int wpt = 32;
    const int64_t want_threads = (int64_t)148 * 2048;
    while (wpt > 4 && E * ceil_div(wpe, wpt) < want_threads) wpt >>= 1;
    if (root_mode) wpt = 1;  // one walk per lane; the window is taken AT the walk's first draw (see below)
    P.wpt = wpt;
    P.cpe = ceil_div(wpe, wpt);
    P.total_chunks = E * P.cpe;
    PN_REQUIRE(P.cpe < ((int64_t)1 << (4 * NDIGITS)) && E < ((int64_t)1 << NPOLY_BITS),
               "pn_merw_sample: too many chunks/epochs in one call");
    const uint64_t draws_per_walk = (uint64_t)dpr * (uint64_t)L;  // every roll draws dpr times, L rolls
    const uint64_t chunk_stride = draws_per_walk * (uint64_t)wpt;
    const uint64_t epoch_stride = draws_per_walk * (uint64_t)n * (uint64_t)nw;
    uint32_t s0[lfg::DEG];
    lfg::seed_window((uint32_t)seed, s0);
    // the thread's first generated word must be s_{341 + first_draw}: its
    // window is s_{310 + first_draw .. 340 + first_draw}
    __uint128_t first = (__uint128_t)((uint64_t)epoch_begin) * epoch_stride +
                        (__uint128_t)((uint64_t)node_begin * (uint64_t)nw) * draws_per_walk;
    // root mode: window words 0..2L-1 ARE the walk's draws (no generation step), so no -31
    first += (uint64_t)(root_mode ? lfg::FIRST_OUTPUT : lfg::FIRST_OUTPUT - lfg::DEG);
    PN_REQUIRE((first >> 64) == 0, "pn_merw_sample: stream offset exceeds 2^64");
    lfg::window_apply(lfg::poly_xpow((uint64_t)first), s0, P.base);
    lfg::Poly pc = lfg::poly_xpow(chunk_stride), pe = lfg::poly_xpow(epoch_stride);
    copy_poly(P.jstep, pc);
    copy_poly(P.jwrap, lfg::poly_xpow(epoch_stride - (uint64_t)(P.cpe - 1) * chunk_stride));
    {
        lfg::Poly acc = lfg::poly_one();
        for (int l = 0; l < 32; l++) { copy_poly(P.jlane[l], acc); acc = lfg::poly_mul(acc, pc); }
    }
    {
        lfg::Poly digit = pc;  // x^(chunk stride * 16^i)
        for (int i = 0; i < NDIGITS; i++) {
            lfg::Poly acc = digit;
            for (int d = 1; d <= 15; d++) { copy_poly(P.jchunk16[i][d - 1], acc); acc = lfg::poly_mul(acc, digit); }
            digit = acc;  // digit^16
        }
    }
    for (int b = 0; b < NPOLY_BITS; b++) { copy_poly(P.jepoch[b], pe); pe = lfg::poly_mul(pe, pe); }
    return PN_OK;
}

// Host-only debug view of the plan (tests/test_sampler_host.py): the tables the kernel
// receives: base[31] jstep[31] jwrap[31] jlane[32][31] jchunk16[8][15][31] jepoch[32][31].
extern "C" int pn_merw_plan_host(int32_t n, int32_t nw, int32_t L, int64_t epoch_begin,
                                 int64_t E, int32_t node_begin, int64_t np, uint64_t seed,
                                 int32_t *wpt, int64_t *cpe, uint32_t *tables)
{
    WalkParams P;
    int rc = plan_walk(P, n, nw, L, epoch_begin, E, node_begin, np, seed);
    if (rc != PN_OK) return rc;
    *wpt = P.wpt; *cpe = P.cpe;
    memcpy(tables, P.base, sizeof(uint32_t) * lfg::DEG * (3 + 32 + NDIGITS * 15 + NPOLY_BITS));
    return PN_OK;
}

extern "C" int pn_merw_sample(pn_graph *g, int32_t nw, int32_t L, int64_t epoch_begin,
                              int64_t epoch_end, int32_t node_begin, int32_t node_end,
                              int32_t rng_mode, uint64_t seed, int32_t *nodes, uint8_t *types,
                              void *stream_)
{
    NvtxScope nvtx("pn_merw_sample");
    PN_REQUIRE(g, "pn_merw_sample: NULL graph");
    PN_REQUIRE(nw >= 0 && L >= 1 && L <= 64, "pn_merw_sample: bad nw=%d / L=%d", nw, L);
    PN_REQUIRE(epoch_begin >= 0 && epoch_end >= epoch_begin, "pn_merw_sample: bad epoch range");
    PN_REQUIRE(node_begin >= 0 && node_end >= node_begin && node_end <= g->n,
               "pn_merw_sample: bad node range [%d,%d) for n=%d", node_begin, node_end, g->n);
    PN_REQUIRE(rng_mode == PN_RNG_GLIBC || rng_mode == PN_RNG_PHILOX, "pn_merw_sample: unknown rng_mode %d", rng_mode);
    if (g->dead_end >= 0) {
        set_error("pn_merw_sample: node %d has no out-rows (the reference prints "
                  "'A.size() == 0 in Alias Table' and exits)", g->dead_end);
        return PN_ERR_DEAD_END;
    }
    cudaStream_t stream = (cudaStream_t)stream_;
    const int64_t E = epoch_end - epoch_begin;
    const int64_t np = node_end - node_begin;
    const int64_t wpe = np * nw;
    if (E == 0 || wpe == 0) return PN_OK;
    PN_REQUIRE(nodes != nullptr, "pn_merw_sample: nodes is NULL");
    PN_CUDA(cudaSetDevice(g->device));
    // label source: the dense table when it fits (24 GiB budget), else per-start-node balls in shared
    // memory (root-centric kernel); PN_MERW_SPARSE=1 forces the latter (tests)
    bool sparse = false;
    if (types && L > 2) {
        const char *force = getenv("PN_MERW_SPARSE");
        if (force && force[0] == '1') sparse = true;
        else {
            int rc = pn_graph_prepare_labels(g, L, 0, stream_);
            if (rc == PN_ERR_UNSUPPORTED) sparse = true;
            else if (rc != PN_OK) return rc;
        }
    }
    WalkParams P;
    {
        int rc = plan_walk(P, g->n, nw, L, epoch_begin, E, node_begin, np, seed, sparse, g->uniform ? 1 : 2);
        if (rc != PN_OK) return rc;
    }
    P.noderec = g->d_noderec; P.cells = g->d_cells;
    P.rng = rng_mode; P.key0 = (uint32_t)seed; P.key1 = (uint32_t)(seed >> 32); P.epoch_begin = epoch_begin; P.epochs = E;
    P.dist = (types && L > 2 && !sparse) ? g->d_dist : nullptr;
    P.dist_pitch = ceil_div(g->n, 16) * 16;
    P.nodes = nodes; P.types = types;
    if (sparse) {
        PN_REQUIRE(2 * L <= 30, "pn_merw_sample: sparse-label mode supports path_len <= 15 (got %d)", L);
        PN_REQUIRE(np < ((int64_t)1 << 31), "pn_merw_sample: too many start nodes in one call");
        if (!g->d_overflow) PN_CUDA(cudaMalloc(&g->d_overflow, sizeof(int)));
        PN_CUDA(cudaMemsetAsync(g->d_overflow, 0, sizeof(int), stream));
        P.rowptr = g->d_rowptr; P.adj = g->d_adj; P.epochs = E; P.overflow = g->d_overflow;
        {
            ProfScope prof("merw_walk", stream);
            PN_CUDA(cudaFuncSetAttribute(k_merw_walk_root, cudaFuncAttributeMaxDynamicSharedMemorySize, ROOT_SMEM));
            k_merw_walk_root<<<(unsigned)np, ROOT_THREADS, ROOT_SMEM, stream>>>(P);
        }
        PN_LAUNCH_CHECK();
        int h = 0;   // the ball-capacity check needs the result: sparse mode synchronises the stream
        PN_CUDA(cudaMemcpyAsync(&h, g->d_overflow, sizeof(int), cudaMemcpyDeviceToHost, stream));
        PN_CUDA(cudaStreamSynchronize(stream));
        if (h) {
            set_error("pn_merw_sample: a start node's radius-%d ball has more than %d nodes; the sparse label "
                      "path keeps it in shared memory (dense table needs n*n bytes <= 24 GiB)", L - 2, ROOT_QCAP);
            return PN_ERR_UNSUPPORTED;
        }
        return PN_OK;
    }
    if (rng_mode == PN_RNG_PHILOX) {
        const int64_t pblocks = ceil_div(E * wpe, 256);
        PN_REQUIRE(pblocks < ((int64_t)1 << 31), "pn_merw_sample: grid too large; split the epoch range");
        ProfScope prof("merw_walk", stream);
        k_merw_walk_philox<<<(unsigned)pblocks, 256, 0, stream>>>(P);
        PN_LAUNCH_CHECK();
        return PN_OK;
    }
    const int64_t blocks = ceil_div(P.total_chunks, WALK_THREADS);
    PN_REQUIRE(blocks < ((int64_t)1 << 31), "pn_merw_sample: grid too large; split the epoch range");
    ProfScope prof("merw_walk", stream);
    k_merw_walk<<<(unsigned)blocks, WALK_THREADS, 0, stream>>>(P);
    PN_LAUNCH_CHECK();
    return PN_OK;
}
