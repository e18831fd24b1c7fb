// The rest of the reference's training step on the device (PathNet_run.py:345-352): mean cross-entropy loss with its
// gradient, and torch.optim.Adam's update over flat buffers, plus the per-step counter that lets a captured CUDA graph
// draw fresh dropout masks and apply the right bias correction on every replay.  Together with pn_pagg_forward /
// pn_pagg_backward the whole step is device work enqueued on one stream: no host round trip, capturable as one graph.
#include <cmath>

#include "common.h"

namespace pn {

// state->step += 1; state->seed_add advances by an odd 64-bit constant (so cfg.dropout_seed + seed_add never repeats)
__global__ void k_step_advance(pn_step_state *s)
{
    s->step += 1;
    s->seed_add += 0x9E3779B97F4A7C15ull;
}

// one warp per row: loss_i = logsumexp(x_i) - x_i[y_i];  dlogits = (softmax(x_i) - onehot(y_i)) * inv_count.
// Labels outside [0, C) contribute nothing (the reference's CrossEntropyLoss would raise; ignore_index semantics
// for -100 are kept: such rows are skipped).
template <typename LT>
__global__ void __launch_bounds__(256) k_ce_loss_grad(int64_t S, int C, const float *__restrict__ logits,
                                                      const LT *__restrict__ labels, float inv_count,
                                                      float *__restrict__ loss, float *__restrict__ dlogits)
{
    const int lane = threadIdx.x & 31;
    const int64_t s = (int64_t)blockIdx.x * 8 + (threadIdx.x >> 5);
    float li = 0.f;
    if (s < S) {
        const float *x = logits + s * C;
        float mx = -INFINITY;
        for (int c = lane; c < C; c += 32) mx = fmaxf(mx, x[c]);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
        float den = 0.f;
        for (int c = lane; c < C; c += 32) den += expf(x[c] - mx);
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) den += __shfl_xor_sync(0xffffffffu, den, o);
        const int64_t y = (int64_t)labels[s];
        const bool ok = y >= 0 && y < C;
        const float lse = mx + logf(den);
        for (int c = lane; c < C; c += 32) {
            const float pr = expf(x[c] - lse);
            if (dlogits) dlogits[s * C + c] = ok ? (pr - (c == y ? 1.f : 0.f)) * inv_count : 0.f;
        }
        if (ok && lane == 0) li = (lse - x[y]) * inv_count;
    }
    // block reduction of the 8 rows' losses, one atomic per block
    __shared__ float part[8];
    if (lane == 0) part[threadIdx.x >> 5] = li;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float t = 0.f;
#pragma unroll
        for (int i = 0; i < 8; i++) t += part[i];
        if (t != 0.f) atomicAdd(loss, t);
    }
}

// torch.optim.Adam (amsgrad=False, maximize=False), L2 weight decay folded into the gradient as torch does
__global__ void __launch_bounds__(256) k_adam_flat(int64_t n, float *__restrict__ p, const float *__restrict__ g,
                                                   float *__restrict__ m, float *__restrict__ v, pn_adam_cfg a,
                                                   const pn_step_state *__restrict__ st)
{
    const float t = (float)st->step;
    const float bc1 = 1.f - powf(a.beta1, t), bc2 = 1.f - powf(a.beta2, t);
    const float step_size = a.lr / bc1, bc2_sqrt = sqrtf(bc2);
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
        const float pi = p[i];
        const float gi = fmaf(a.weight_decay, pi, g[i]);
        const float mi = fmaf(a.beta1, m[i], (1.f - a.beta1) * gi);
        const float vi = fmaf(a.beta2, v[i], (1.f - a.beta2) * gi * gi);
        m[i] = mi;
        v[i] = vi;
        const float denom = sqrtf(vi) / bc2_sqrt + a.eps;
        p[i] = pi - step_size * (mi / denom);
    }
}

}  // namespace pn

using namespace pn;

extern "C" int pn_step_advance(pn_step_state *state, void *stream)
{
    PN_REQUIRE(state, "step_advance: NULL state");
    k_step_advance<<<1, 1, 0, (cudaStream_t)stream>>>(state);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

extern "C" int pn_ce_loss_grad(int64_t S, int32_t C, const float *logits, const void *labels, int32_t label_dtype,
                               float inv_count, float *loss, float *dlogits, void *stream)
{
    PN_REQUIRE(S >= 0 && C > 0, "ce_loss_grad: bad sizes S=%lld C=%d", (long long)S, C);
    PN_REQUIRE(label_dtype == PN_IDX_I64 || label_dtype == PN_IDX_I32, "ce_loss_grad: labels must be int64 or int32");
    if (S == 0) return PN_OK;
    PN_REQUIRE(logits && labels, "ce_loss_grad: NULL pointer");
    const unsigned grid = (unsigned)ceil_div(S, 8);
    cudaStream_t st = (cudaStream_t)stream;
    if (label_dtype == PN_IDX_I64) k_ce_loss_grad<int64_t><<<grid, 256, 0, st>>>(S, C, logits, (const int64_t *)labels, inv_count, loss, dlogits);
    else k_ce_loss_grad<int32_t><<<grid, 256, 0, st>>>(S, C, logits, (const int32_t *)labels, inv_count, loss, dlogits);
    PN_LAUNCH_CHECK();
    return PN_OK;
}

extern "C" int pn_adam_step(int64_t n, float *params, const float *grads, float *exp_avg, float *exp_avg_sq,
                            const pn_adam_cfg *cfg, const pn_step_state *state, void *stream)
{
    PN_REQUIRE(n >= 0 && cfg && state, "adam_step: bad arguments");
    if (n == 0) return PN_OK;
    PN_REQUIRE(params && grads && exp_avg && exp_avg_sq, "adam_step: NULL pointer");
    const unsigned grid = (unsigned)std::min<int64_t>(ceil_div(n, 256), 148 * 8);
    k_adam_flat<<<grid, 256, 0, (cudaStream_t)stream>>>(n, params, grads, exp_avg, exp_avg_sq, *cfg, state);
    PN_LAUNCH_CHECK();
    return PN_OK;
}
