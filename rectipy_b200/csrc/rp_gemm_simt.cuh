// fp32 FFMA contractions (exact-fp32 path, any shape).  Used for small / ragged networks, for B < 16 trials,
// and as the strict-fp32 comparator of the tcgen05 3xTF32 path (rp_gemm_tc.cuh).
//
// All three contractions of the engine have the form   C[q][p] (+)= sum_k Aop(p,k) * Bop(q,k)   (p fastest in C):
//   forward   u[b][i]   = sum_j (kW)[i][j]   * src[b][j]       A,B "K-major"  (k contiguous)
//   dgrad     Z[b][j]   = sum_i (kW)^T[j][i] * g[b][i]         A,B K-major
//   wgrad     dW[i][j] += sum_b src[b][j]    * g[b][i]         A,B "MN-major" (p / q contiguous), p=j, q=i
#pragma once
#include <cuda_runtime.h>

namespace rp {

constexpr int SG_BM = 128, SG_BN = 128, SG_BK = 16, SG_PAD = 4;

// KMAJOR: A[p*lda + k], B[q*ldb + k]       else: A[k*lda + p], B[k*ldb + q]
// VEC: 16-byte loads allowed (all leading dims and extents along the contiguous dim are multiples of 4, bases aligned)
template <bool KMAJOR, bool VEC>
__global__ void __launch_bounds__(256) k_sgemm(int P, int Q, int K, const float* __restrict__ A, int lda,
                                               const float* __restrict__ Bm, int ldb, float* __restrict__ C, int ldc,
                                               int accumulate) {
    __shared__ __align__(16) float As[2][SG_BK][SG_BM + SG_PAD];
    __shared__ __align__(16) float Bs[2][SG_BK][SG_BN + SG_PAD];
    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;             // tx -> p (fast output dim), ty -> q
    const int p0 = blockIdx.x * SG_BM, q0 = blockIdx.y * SG_BN;

    float acc[8][8];
#pragma unroll
    for (int r = 0; r < 8; ++r)
#pragma unroll
        for (int c = 0; c < 8; ++c) acc[r][c] = 0.f;

    float4 ra[2], rb[2];
    const int nk = (K + SG_BK - 1) / SG_BK;

    auto load_tile = [&](int kt) {
        const int k0 = kt * SG_BK;
        if constexpr (KMAJOR) {
            // 128 rows x 16 k : 4 float4 per row, 512 float4, 2 per thread
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = (tid >> 2) + 64 * r, kv = (tid & 3) * 4;
                float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                const int pa = p0 + row, qb = q0 + row, kk = k0 + kv;
                if (VEC) {
                    if (pa < P && kk < K) va = *reinterpret_cast<const float4*>(A + (size_t)pa * lda + kk);
                    if (qb < Q && kk < K) vb = *reinterpret_cast<const float4*>(Bm + (size_t)qb * ldb + kk);
                } else {
                    float* fa = reinterpret_cast<float*>(&va); float* fb = reinterpret_cast<float*>(&vb);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (pa < P && kk + e < K) fa[e] = A[(size_t)pa * lda + kk + e];
                        if (qb < Q && kk + e < K) fb[e] = Bm[(size_t)qb * ldb + kk + e];
                    }
                }
                ra[r] = va; rb[r] = vb;
            }
        } else {
            // 16 k-rows x 128 contiguous : 32 float4 per row, 512 float4, 2 per thread
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int krow = (tid >> 5) + 8 * r, cv = (tid & 31) * 4;
                float4 va = make_float4(0.f, 0.f, 0.f, 0.f), vb = va;
                const int kk = k0 + krow;
                if (VEC) {
                    if (kk < K && p0 + cv < P) va = *reinterpret_cast<const float4*>(A + (size_t)kk * lda + p0 + cv);
                    if (kk < K && q0 + cv < Q) vb = *reinterpret_cast<const float4*>(Bm + (size_t)kk * ldb + q0 + cv);
                } else {
                    float* fa = reinterpret_cast<float*>(&va); float* fb = reinterpret_cast<float*>(&vb);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        if (kk < K && p0 + cv + e < P) fa[e] = A[(size_t)kk * lda + p0 + cv + e];
                        if (kk < K && q0 + cv + e < Q) fb[e] = Bm[(size_t)kk * ldb + q0 + cv + e];
                    }
                }
                ra[r] = va; rb[r] = vb;
            }
        }
    };
    auto store_tile = [&](int buf) {
        if constexpr (KMAJOR) {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int row = (tid >> 2) + 64 * r, kv = (tid & 3) * 4;
                const float* fa = reinterpret_cast<const float*>(&ra[r]);
                const float* fb = reinterpret_cast<const float*>(&rb[r]);
#pragma unroll
                for (int e = 0; e < 4; ++e) { As[buf][kv + e][row] = fa[e]; Bs[buf][kv + e][row] = fb[e]; }
            }
        } else {
#pragma unroll
            for (int r = 0; r < 2; ++r) {
                const int krow = (tid >> 5) + 8 * r, cv = (tid & 31) * 4;
                *reinterpret_cast<float4*>(&As[buf][krow][cv]) = ra[r];
                *reinterpret_cast<float4*>(&Bs[buf][krow][cv]) = rb[r];
            }
        }
    };

    load_tile(0);
    store_tile(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) load_tile(kt + 1);
#pragma unroll
        for (int kk = 0; kk < SG_BK; ++kk) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][kk][tx * 4]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][kk][64 + tx * 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][kk][ty * 4]);
            const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][kk][64 + ty * 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int c = 0; c < 8; ++c) acc[r][c] = fmaf(bv[r], av[c], acc[r][c]);   // acc[q][p]
        }
        if (kt + 1 < nk) {
            store_tile(buf ^ 1);
            __syncthreads();
        }
    }

    // epilogue: C[q*ldc + p], p contiguous -> float4 stores along p
#pragma unroll
    for (int r = 0; r < 8; ++r) {
        const int q = q0 + (r < 4 ? ty * 4 + r : 64 + ty * 4 + (r - 4));
        if (q >= Q) continue;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
            const int p = p0 + h * 64 + tx * 4;
            float* dst = C + (size_t)q * ldc + p;
            if (VEC && p + 3 < P) {
                float4 o = make_float4(acc[r][h * 4 + 0], acc[r][h * 4 + 1], acc[r][h * 4 + 2], acc[r][h * 4 + 3]);
                if (accumulate) { const float4 c = *reinterpret_cast<const float4*>(dst); o.x += c.x; o.y += c.y; o.z += c.z; o.w += c.w; }
                *reinterpret_cast<float4*>(dst) = o;
            } else {
#pragma unroll
                for (int e = 0; e < 4; ++e)
                    if (p + e < P) dst[e] = accumulate ? dst[e] + acc[r][h * 4 + e] : acc[r][h * 4 + e];
            }
        }
    }
}

// Few trials (Q <= 8 per pass): one warp per output row p streams A[p][:] once and keeps Q accumulators.
// HBM/L2-bound on A; this is the reference's own GEMV shape (edges.py:49, nodes.py:169) for B = 1.
template <bool VEC>
__global__ void __launch_bounds__(256) k_gemv_rows(int P, int Q, int K, const float* __restrict__ A, int lda,
                                                   const float* __restrict__ Bm, int ldb, float* __restrict__ C, int ldc) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= P) return;
    const float* arow = A + (size_t)warp * lda;
    float acc[8];
#pragma unroll
    for (int q = 0; q < 8; ++q) acc[q] = 0.f;
    if (VEC) {
        for (int k = lane * 4; k < K; k += 128) {
            const float4 a = *reinterpret_cast<const float4*>(arow + k);
#pragma unroll
            for (int q = 0; q < 8; ++q) {
                if (q < Q) {
                    const float4 b = *reinterpret_cast<const float4*>(Bm + (size_t)q * ldb + k);
                    acc[q] = fmaf(a.x, b.x, acc[q]); acc[q] = fmaf(a.y, b.y, acc[q]);
                    acc[q] = fmaf(a.z, b.z, acc[q]); acc[q] = fmaf(a.w, b.w, acc[q]);
                }
            }
        }
    } else {
        for (int k = lane; k < K; k += 32) {
            const float a = arow[k];
#pragma unroll
            for (int q = 0; q < 8; ++q)
                if (q < Q) acc[q] = fmaf(a, Bm[(size_t)q * ldb + k], acc[q]);
        }
    }
#pragma unroll
    for (int q = 0; q < 8; ++q) {
        if (q < Q) {
            float v = acc[q];
            for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
            if (lane == 0) C[(size_t)q * ldc + warp] = v;
        }
    }
}

// Few trials: dW[i][j] += sum_b g[b][i] * src[b][j]   (rank-B update, bound by the read-modify-write of dW)
static __global__ void __launch_bounds__(256) k_outer_acc(int N, int Bq, const float* __restrict__ g, int ldg,
                                                   const float* __restrict__ src, int lds, float* __restrict__ dW, int ldw) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    const int i = blockIdx.y;
    if (j >= N) return;
    float acc = dW[(size_t)i * ldw + j];
    for (int b = 0; b < Bq; ++b) acc = fmaf(__ldg(g + (size_t)b * ldg + i), __ldg(src + (size_t)b * lds + j), acc);
    dW[(size_t)i * ldw + j] = acc;
}

}  // namespace rp
