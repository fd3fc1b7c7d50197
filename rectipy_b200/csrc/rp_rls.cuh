// Recursive-least-squares readout training over a recorded state matrix X[T][n]   (edges.py:227-234,
// network.py:1093-1121).  The update is strictly sequential in t; per step the work is one n x n GEMV and one
// rank-1 update of P, i.e. bound by the read-modify-write of P (L2 resident for the reference sizes, n <= ~2000).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>

#define RP_RLS_MAX_OUT 16

namespace rp {

inline char* rls_err_buf() { static thread_local char buf[256] = ""; return buf; }
inline const char* rls_last_error() { return rls_err_buf(); }

// z[r] = beta_inv * sum_c P[r][c] x[c]     (one warp per row)
__global__ void __launch_bounds__(256) k_rls_z(int n, float beta_inv, const float* __restrict__ P, const float* __restrict__ x, float* __restrict__ z) {
    const int row = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (row >= n) return;
    float acc = 0.f;
    for (int c = lane; c < n; c += 32) acc = fmaf(P[(size_t)row * n + c], __ldg(x + c), acc);
    for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
    if (lane == 0) z[row] = beta_inv * acc;
}

// single block: kappa = 1/(1+x.z); y_hat = W x; W += outer(y - kappa*(W x + y (z.x)), z); loss = |y - y_hat|^2
__global__ void __launch_bounds__(256) k_rls_w(int n, int k, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ z,
                                               float* __restrict__ W, float* kappa_out, float* loss_t, float* pred_t, int do_update) {
    __shared__ float red[8];
    __shared__ float sh[RP_RLS_MAX_OUT + 1];
    float part[RP_RLS_MAX_OUT + 1];
#pragma unroll
    for (int q = 0; q <= RP_RLS_MAX_OUT; ++q) part[q] = 0.f;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const float xc = x[c];
        part[RP_RLS_MAX_OUT] = fmaf(xc, z[c], part[RP_RLS_MAX_OUT]);
#pragma unroll
        for (int q = 0; q < RP_RLS_MAX_OUT; ++q) if (q < k) part[q] = fmaf(W[(size_t)q * n + c], xc, part[q]);
    }
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    for (int q = 0; q <= RP_RLS_MAX_OUT; ++q) {
        if (q < k || q == RP_RLS_MAX_OUT) {
            float v = part[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            __syncthreads();
            if (l == 0) red[w] = v;
            __syncthreads();
            if (threadIdx.x == 0) { float t = 0.f; for (int r = 0; r < 8; ++r) t += red[r]; sh[q] = t; }
        }
    }
    __syncthreads();
    const float xz = sh[RP_RLS_MAX_OUT];
    const float kappa = 1.0f / (1.0f + xz);
    if (threadIdx.x == 0) {
        *kappa_out = do_update ? kappa : 0.f;
        float ls = 0.f;
        for (int q = 0; q < k; ++q) { const float e = y[q] - sh[q]; ls = fmaf(e, e, ls); if (pred_t) pred_t[q] = sh[q]; }
        if (loss_t) *loss_t = ls;
    }
    if (do_update) {
        for (int c = threadIdx.x; c < n; c += blockDim.x) {
            const float zc = z[c];
            for (int q = 0; q < k; ++q) {
                const float coef = y[q] - kappa * (sh[q] + y[q] * xz);
                W[(size_t)q * n + c] = fmaf(coef, zc, W[(size_t)q * n + c]);
            }
        }
    }
}

// P[r][c] -= kappa * z[r] * z[c]
__global__ void __launch_bounds__(256) k_rls_p(int n, const float* __restrict__ z, const float* __restrict__ kappa, float* __restrict__ P) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x, r = blockIdx.y;
    if (c >= n) return;
    const float kp = *kappa;
    if (kp == 0.f) return;
    P[(size_t)r * n + c] = fmaf(-kp * z[r], z[c], P[(size_t)r * n + c]);
}

inline int rls_run(int T, int n, int k, float beta_inv, const float* X, const float* Y, float* W, float* P,
                   float* loss, float* pred, int update_every, cudaStream_t st) {
    float* scratch = nullptr;   // z[n] + kappa
    cudaError_t e = cudaMallocAsync(reinterpret_cast<void**>(&scratch), (size_t)(n + 1) * sizeof(float), st);
    if (e != cudaSuccess) { snprintf(rls_err_buf(), 256, "cudaMallocAsync: %s", cudaGetErrorString(e)); return 1; }
    float* z = scratch; float* kappa = scratch + n;
    for (int t = 0; t < T; ++t) {
        const int upd = (t % update_every == 0) ? 1 : 0;
        const float* x = X + (size_t)t * n;
        k_rls_z<<<(n + 7) / 8, 256, 0, st>>>(n, beta_inv, P, x, z);
        k_rls_w<<<1, 256, 0, st>>>(n, k, x, Y + (size_t)t * k, z, W, kappa, loss ? loss + t : nullptr, pred ? pred + (size_t)t * k : nullptr, upd);
        if (upd) k_rls_p<<<dim3((n + 255) / 256, n), 256, 0, st>>>(n, z, kappa, P);
    }
    e = cudaGetLastError();
    cudaFreeAsync(scratch, st);
    if (e != cudaSuccess) { snprintf(rls_err_buf(), 256, "launch: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

}  // namespace rp
