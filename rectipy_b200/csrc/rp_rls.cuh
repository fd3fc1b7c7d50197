// Recursive-least-squares readout training over a recorded state matrix X[T][n]   (edges.py:227-234,
// network.py:1093-1121).  The update is strictly sequential in t; per step the work is one n x n GEMV and one
// rank-1 update of P, i.e. bound by the read-modify-write of P (L2 resident for the reference sizes, n <= ~2000).
#pragma once
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdlib.h>
#include <algorithm>
#include <map>
#include <mutex>
#include <utility>
#include "rp_persistent.cuh"

#define RP_RLS_MAX_OUT 16

namespace rp {

inline char* rls_err_buf() { static thread_local char buf[256] = ""; return buf; }
inline const char* rls_last_error() { return rls_err_buf(); }

// Scratch of an rp_rls_run call (a few KB: the exchange buffer of the persistent kernel, or z + kappa of the per-step kernels).
// Cached per (device, stream) and only ever grown: the stream-ordered allocator (cudaMallocAsync / cudaFreeAsync per call) gives its
// memory back at every synchronisation and re-maps it on the next call, which showed up as intermittent stalls of 0.1 - 1.4 s on a
// 25 ms call (round 1f, tools/exp_rls_var.py).  Calls on one stream are ordered, so reuse needs no further synchronisation.
inline void* rls_scratch(size_t bytes, cudaStream_t st) {
    struct Buf { void* p = nullptr; size_t cap = 0; };
    static std::mutex mu;
    static std::map<std::pair<int, cudaStream_t>, Buf> cache;
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    Buf& b = cache[std::make_pair(dev, st)];
    if (b.cap < bytes) {
        if (b.p) { cudaStreamSynchronize(st); cudaFree(b.p); b.p = nullptr; b.cap = 0; }
        const size_t want = std::max<size_t>(bytes, 64 * 1024);
        if (cudaMalloc(&b.p, want) != cudaSuccess) { b.p = nullptr; return nullptr; }
        b.cap = want;
    }
    return b.p;
}

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

// ---- persistent variant ---------------------------------------------------------------------------------------------------
// One cooperative launch for all T steps.  CTA c owns rows [c*R, c*R+R) of P in shared memory for the whole run; every CTA
// keeps its own copy of W (k x n, updated identically everywhere) so the prediction needs no communication.  The only
// per-step exchange is the vector z = beta^-1 P x (n floats), published with the flag-in-data protocol of rp_persistent.cuh.
struct RlsArgs {
    int T, n, k, npad, rows_per_cta, update_every;
    float beta_inv;
    const float* X; const float* Y;
    float* W; float* P; float* loss; float* pred;
    uint2* zbuf;            // [2][npad] {value, tag}, tags zeroed before the launch
};

__global__ void __launch_bounds__(PS_THREADS, 1) k_rls_persistent(RlsArgs a) {
    extern __shared__ __align__(16) float rsm[];
    const int n = a.n, npad = a.npad, k = a.k;
    float* s_x = rsm;                      // [npad]
    float* s_z = s_x + npad;               // [npad]
    float* s_red = s_z + npad;             // [32 + 32]
    float* s_W = s_red + 64;               // [k][npad]
    float* s_P = s_W + (size_t)k * npad;   // [R][npad]
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int R = max(0, min(a.rows_per_cta, n - r0));
    for (int idx = tid; idx < k * npad; idx += PS_THREADS) { const int q = idx / npad, c = idx % npad; s_W[idx] = c < n ? a.W[(size_t)q * n + c] : 0.f; }
    for (int idx = tid; idx < R * npad; idx += PS_THREADS) { const int r = idx / npad, c = idx % npad; s_P[idx] = c < n ? a.P[(size_t)(r0 + r) * n + c] : 0.f; }
    if (tid < npad - n) { s_x[n + tid] = 0.f; s_z[n + tid] = 0.f; }
    __syncthreads();
    unsigned int tag = 0;
    for (int t = 0; t < a.T; ++t) {
        const bool upd = (t % a.update_every) == 0;
        for (int c = tid; c < n; c += PS_THREADS) s_x[c] = __ldg(a.X + (size_t)t * n + c);
        __syncthreads();
        // prediction with the weights before this step's update: y_hat = W x  (warp q computes output q)
        for (int q = warp; q < k; q += PS_THREADS / 32) {
            float acc = 0.f;
            for (int c = lane; c < n; c += 32) acc = fmaf(s_W[q * npad + c], s_x[c], acc);
            for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
            if (lane == 0) s_red[q] = acc;
        }
        if (upd) {
            ++tag;
            // z_r = beta^-1 * P[r][:] . x for the owned rows, published to everyone
            for (int r = warp; r < R; r += PS_THREADS / 32) {
                float acc = 0.f;
                for (int c = lane; c < n; c += 32) acc = fmaf(s_P[r * npad + c], s_x[c], acc);
                for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
                if (lane == 0) ll_store(a.zbuf + (size_t)(tag & 1) * npad + r0 + r, a.beta_inv * acc, tag);
            }
            ll_gather(a.zbuf + (size_t)(tag & 1) * npad, s_z, 1, n, npad, tag);
            __syncthreads();
            // x . z (every CTA computes it: n is small)
            float part = 0.f;
            for (int c = tid; c < n; c += PS_THREADS) part = fmaf(s_x[c], s_z[c], part);
            for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
            if (lane == 0) s_red[32 + warp] = part;
            __syncthreads();
            float xz = 0.f;
#pragma unroll
            for (int w8 = 0; w8 < PS_THREADS / 32; ++w8) xz += s_red[32 + w8];
            const float kappa = 1.0f / (1.0f + xz);
            // P[r][c] -= kappa z_r z_c (owned rows) ;  W[q][c] += (y_q - kappa (y_hat_q + y_q xz)) z_c (all rows, every CTA)
            for (int idx = tid; idx < R * n; idx += PS_THREADS) {
                const int r = idx / n, c = idx - r * n;
                s_P[r * npad + c] = fmaf(-kappa * s_z[r0 + r], s_z[c], s_P[r * npad + c]);
            }
            for (int idx = tid; idx < k * n; idx += PS_THREADS) {
                const int q = idx / n, c = idx - q * n;
                const float yq = __ldg(a.Y + (size_t)t * k + q);
                s_W[q * npad + c] = fmaf(yq - kappa * (s_red[q] + yq * xz), s_z[c], s_W[q * npad + c]);
            }
        } else {
            __syncthreads();
        }
        if (blockIdx.x == 0 && tid == 0) {
            float ls = 0.f;
            for (int q = 0; q < k; ++q) {
                const float e = __ldg(a.Y + (size_t)t * k + q) - s_red[q];
                ls = fmaf(e, e, ls);
                if (a.pred) a.pred[(size_t)t * k + q] = s_red[q];
            }
            if (a.loss) a.loss[t] = ls;
        }
        __syncthreads();
    }
    for (int idx = tid; idx < R * n; idx += PS_THREADS) { const int r = idx / n, c = idx - r * n; a.P[(size_t)(r0 + r) * n + c] = s_P[r * npad + c]; }
    if (blockIdx.x == 0) for (int idx = tid; idx < k * n; idx += PS_THREADS) { const int q = idx / n, c = idx - q * n; a.W[(size_t)q * n + c] = s_W[q * npad + c]; }
}

// returns 0 launched, 1 error, 2 shape does not fit the persistent kernel (caller falls back to per-step launches)
inline int rls_run_persistent(int T, int n, int k, float beta_inv, const float* X, const float* Y, float* W, float* P,
                              float* loss, float* pred, int update_every, cudaStream_t st) {
    int dev = 0, sms = 0, coop = 0, smem_max = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    cudaDeviceGetAttribute(&smem_max, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    if (!coop || getenv("RP_NO_PERSISTENT")) return 2;
    const int npad = (n + 3) / 4 * 4;
    int rows = (n + sms - 1) / sms;
    rows = (rows + 7) / 8 * 8;
    const int grid = (n + rows - 1) / rows;
    const size_t smem = ((size_t)2 * npad + 64 + (size_t)k * npad + (size_t)rows * npad) * sizeof(float);
    if (smem > (size_t)std::min(smem_max, 200 * 1024)) return 2;
    if (cudaFuncSetAttribute(k_rls_persistent, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess) return 2;
    int occ = 0;
    cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_rls_persistent, PS_THREADS, smem);
    if (occ < 1 || grid > occ * sms) return 2;
    uint2* zbuf = reinterpret_cast<uint2*>(rls_scratch(2 * (size_t)npad * sizeof(uint2), st));
    if (!zbuf) { snprintf(rls_err_buf(), 256, "cudaMalloc of the RLS exchange buffer failed"); return 1; }
    cudaError_t e = cudaMemsetAsync(zbuf, 0, 2 * (size_t)npad * sizeof(uint2), st);
    RlsArgs a{T, n, k, npad, rows, update_every, beta_inv, X, Y, W, P, loss, pred, zbuf};
    void* args[] = {&a};
    if (e == cudaSuccess) e = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(k_rls_persistent), dim3(grid), dim3(PS_THREADS), args, smem, st);
    if (e != cudaSuccess) { snprintf(rls_err_buf(), 256, "cooperative launch: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

inline int rls_run(int T, int n, int k, float beta_inv, const float* X, const float* Y, float* W, float* P,
                   float* loss, float* pred, int update_every, cudaStream_t st) {
    if (T > 0) {
        const int rc = rls_run_persistent(T, n, k, beta_inv, X, Y, W, P, loss, pred, update_every, st);
        if (rc != 2) return rc;
    }
    float* scratch = reinterpret_cast<float*>(rls_scratch((size_t)(n + 1) * sizeof(float), st));   // z[n] + kappa
    if (!scratch) { snprintf(rls_err_buf(), 256, "cudaMalloc of the RLS scratch failed"); return 1; }
    float* z = scratch; float* kappa = scratch + n;
    for (int t = 0; t < T; ++t) {
        const int upd = (t % update_every == 0) ? 1 : 0;
        const float* x = X + (size_t)t * n;
        k_rls_z<<<(n + 7) / 8, 256, 0, st>>>(n, beta_inv, P, x, z);
        k_rls_w<<<1, 256, 0, st>>>(n, k, x, Y + (size_t)t * k, z, W, kappa, loss ? loss + t : nullptr, pred ? pred + (size_t)t * k : nullptr, upd);
        if (upd) k_rls_p<<<dim3((n + 255) / 256, n), 256, 0, st>>>(n, z, kappa, P);
    }
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { snprintf(rls_err_buf(), 256, "launch: %s", cudaGetErrorString(e)); return 1; }
    return 0;
}

}  // namespace rp
