// MERW transition probabilities on the GPU — replaces preprocess/compute_merw.py:109-121
// (`compute_merw`: dominant eigenpair of A by scipy eigsh, then an O(n^2) Python loop
// P[u,v] = A[u,v] psi[v] / (lambda psi[u])) and feeds init_rw.py:81-86's row emission.
//
//   * connected components by min-label propagation (the reference ignores components, which
//     is where the negative / zero p of cora, citeseer, bgp ... come from: eigsh's vector lives
//     on one component only);
//   * per component: power iteration on A + I in fp64 (the shift removes the -lambda twin of
//     bipartite components), x <- (A x + x) / ||.||_component, lambda = ||A x + x|| - 1;
//   * p[e] = w[e] x[col[e]] / (lambda[comp] x[row]).
// On a connected graph this is exactly the reference's P (up to eigsh's convergence).
#include "common.h"

namespace pn {

__global__ void k_cc_init(int32_t *label, int32_t n)
{
    int32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) label[i] = i;
}

// one sweep of min-label propagation over CSR rows, with pointer jumping
__global__ void k_cc_sweep(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, int32_t *label,
                           int32_t n, int *changed)
{
    int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    int32_t best = label[u];
    for (int64_t e = rowptr[u]; e < rowptr[u + 1]; e++) best = min(best, label[col[e]]);
    best = min(best, label[best]);  // jump
    if (best < label[u]) { atomicMin(&label[u], best); *changed = 1; }
}

// y = A x + x ; norm2[label[u]] += y_u^2       (one warp per row)
__global__ void k_spmv_shift(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col,
                             const double *__restrict__ w, const double *__restrict__ x, double *__restrict__ y,
                             const int32_t *__restrict__ label, double *norm2, int32_t n)
{
    const int32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    double s = 0.0;
    for (int64_t e = rowptr[u] + lane; e < rowptr[u + 1]; e += 32) s += (w ? w[e] : 1.0) * x[col[e]];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        s += x[u];
        y[u] = s;
        atomicAdd(&norm2[label[u]], s * s);
    }
}

// x = y / ||y||_component ; delta = max |x_new - x_old| ; lambda[c] = ||y||_c - 1
__global__ void k_normalize(double *__restrict__ x, const double *__restrict__ y, const int32_t *__restrict__ label,
                            const double *__restrict__ norm2, double *__restrict__ lambda, unsigned long long *delta_bits,
                            int32_t n)
{
    int32_t u = blockIdx.x * blockDim.x + threadIdx.x;
    if (u >= n) return;
    const int32_t c = label[u];
    const double nr = sqrt(norm2[c]);
    const double xn = nr > 0.0 ? y[u] / nr : 0.0;
    const double d = fabs(xn - x[u]);
    x[u] = xn;
    if (u == c) lambda[c] = nr - 1.0;
    atomicMax(delta_bits, (unsigned long long)__double_as_longlong(d));  // d >= 0: bit pattern is monotone
}

__global__ void k_fill(double *a, double v, int64_t n)
{
    int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) a[i] = v;
}

__global__ void k_merw_p(const int64_t *__restrict__ rowptr, const int32_t *__restrict__ col, const double *__restrict__ w,
                         const double *__restrict__ x, const double *__restrict__ lambda, const int32_t *__restrict__ label,
                         double *__restrict__ p, int32_t n)
{
    const int32_t u = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int lane = threadIdx.x & 31;
    if (u >= n) return;
    const double denom = lambda[label[u]] * x[u];
    for (int64_t e = rowptr[u] + lane; e < rowptr[u + 1]; e += 32) p[e] = (w ? w[e] : 1.0) * x[col[e]] / denom;
}

}  // namespace pn
using namespace pn;

extern "C" int pn_merw_probabilities(int32_t n, const int64_t *rowptr, const int32_t *col, const double *w, double *p,
                                     double *psi, double *lambda, int32_t *component, double tol, int32_t max_iter,
                                     int32_t *iters_out, void *stream)
{
    PN_REQUIRE(n > 0 && rowptr && col && p && psi && lambda && component, "pn_merw_probabilities: NULL pointer / n <= 0");
    PN_REQUIRE(tol > 0 && max_iter > 0, "pn_merw_probabilities: need tol > 0, max_iter > 0");
    cudaStream_t st = (cudaStream_t)stream;
    double *y = nullptr, *norm2 = nullptr;
    int *changed = nullptr;
    unsigned long long *delta = nullptr;
    PN_CUDA(cudaMallocAsync(&y, sizeof(double) * n, st));
    PN_CUDA(cudaMallocAsync(&norm2, sizeof(double) * n, st));
    PN_CUDA(cudaMallocAsync(&changed, sizeof(int), st));
    PN_CUDA(cudaMallocAsync(&delta, sizeof(unsigned long long), st));
    const unsigned nb = (unsigned)ceil_div(n, 256), nbw = (unsigned)ceil_div((int64_t)n * 32, 256);
    int rc = PN_OK, it = 0;
    // ---- connected components ----
    k_cc_init<<<nb, 256, 0, st>>>(component, n);
    count_launch();
    for (int sweep = 0; sweep < n + 1; sweep++) {
        int h = 0;
        cudaMemsetAsync(changed, 0, sizeof(int), st);
        for (int r = 0; r < 4; r++) { k_cc_sweep<<<nb, 256, 0, st>>>(rowptr, col, component, n, changed); count_launch(); }
        cudaMemcpyAsync(&h, changed, sizeof(int), cudaMemcpyDeviceToHost, st);
        cudaStreamSynchronize(st);
        if (!h) break;
    }
    // ---- power iteration ----
    k_fill<<<nb, 256, 0, st>>>(psi, 1.0, n);
    count_launch();
    for (; it < max_iter;) {
        unsigned long long hbits = 0;
        for (int r = 0; r < 32 && it < max_iter; r++, it++) {
            cudaMemsetAsync(norm2, 0, sizeof(double) * n, st);
            cudaMemsetAsync(delta, 0, sizeof(unsigned long long), st);
            k_spmv_shift<<<nbw, 256, 0, st>>>(rowptr, col, w, psi, y, component, norm2, n);
            k_normalize<<<nb, 256, 0, st>>>(psi, y, component, norm2, lambda, delta, n);
            count_launch(2);
        }
        cudaMemcpyAsync(&hbits, delta, sizeof(hbits), cudaMemcpyDeviceToHost, st);
        if (cudaStreamSynchronize(st) != cudaSuccess) break;
        double d;
        memcpy(&d, &hbits, sizeof d);
        if (it > 1 && d < tol) break;
    }
    k_merw_p<<<nbw, 256, 0, st>>>(rowptr, col, w, psi, lambda, component, p, n);
    count_launch();
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("pn_merw_probabilities: %s", cudaGetErrorString(e)); rc = PN_ERR_CUDA; }
    cudaFreeAsync(y, st); cudaFreeAsync(norm2, st); cudaFreeAsync(changed, st); cudaFreeAsync(delta, st);
    if (iters_out) *iters_out = it;
    return rc;
}
