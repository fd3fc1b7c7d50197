// Second translation unit of librectipy_b200.so: the persistent few-trial kernels (rp_persistent.cuh) and the FFMA contractions
// (rp_gemm_simt.cuh), compiled WITHOUT -split-compile (see rp_fp32_paths.cuh for the reason).  Only host launchers cross the unit
// boundary; the kernels shared through the headers have internal linkage.
#include <cuda_runtime.h>
#include <algorithm>

#include "../../include/rectipy_b200.h"
#include "rp_kernels.cuh"
#include "rp_gemm_simt.cuh"
#include "rp_persistent.cuh"
#include "rp_fp32_paths.cuh"

namespace rp {
namespace {

template <typename K, typename A>
int ps_launch(K kernel, const A* pa, int grid, size_t smem, cudaStream_t st, cudaError_t* err, int* occ_out, int* sms_out) {
    if ((*err = cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem)) != cudaSuccess) return 1;
    int occ = 0, dev = 0, sms = 0;
    if ((*err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, PS_THREADS, smem)) != cudaSuccess) return 1;
    if ((*err = cudaGetDevice(&dev)) != cudaSuccess) return 1;
    if ((*err = cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev)) != cudaSuccess) return 1;
    *occ_out = occ; *sms_out = sms;
    if (occ < 1 || grid > occ * sms) return 2;
    A copy = *pa;
    void* args[] = {&copy};
    *err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), dim3(grid), dim3(PS_THREADS), args, smem, st);
    return *err == cudaSuccess ? 0 : 1;
}

inline bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

}  // namespace

#define RP_PS_DISPATCH(KERNEL)                                                                                   \
    switch (model) {                                                                                             \
        case RP_LI_TANH:    return ps_launch(KERNEL<RP_LI_TANH>, pa, grid, smem, st, err, occ_out, sms_out);     \
        case RP_LI_SIGMOID: return ps_launch(KERNEL<RP_LI_SIGMOID>, pa, grid, smem, st, err, occ_out, sms_out);  \
        case RP_QIF:        return ps_launch(KERNEL<RP_QIF>, pa, grid, smem, st, err, occ_out, sms_out);         \
        case RP_QIF_SFA:    return ps_launch(KERNEL<RP_QIF_SFA>, pa, grid, smem, st, err, occ_out, sms_out);     \
        case RP_LIF:        return ps_launch(KERNEL<RP_LIF>, pa, grid, smem, st, err, occ_out, sms_out);         \
        case RP_IK:         return ps_launch(KERNEL<RP_IK>, pa, grid, smem, st, err, occ_out, sms_out);          \
        default: return 3;                                                                                       \
    }

int ps_launch_fwd(int model, const PersistFwdArgs* pa, int grid, size_t smem, cudaStream_t st, cudaError_t* err, int* occ_out, int* sms_out) {
    RP_PS_DISPATCH(k_persist_fwd)
}
int ps_launch_bwd(int model, const PersistBwdArgs* pa, int grid, size_t smem, cudaStream_t st, cudaError_t* err, int* occ_out, int* sms_out) {
    RP_PS_DISPATCH(k_persist_bwd)
}

cudaError_t gemm_fp32_launch(bool kmajor, int P, int Q, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int accumulate,
                             cudaStream_t st, long long* launches) {
    if (P <= 0 || Q <= 0) return cudaSuccess;
    if (K <= 0) return accumulate ? cudaSuccess : cudaMemset2DAsync(C, (size_t)ldc * sizeof(float), 0, (size_t)P * sizeof(float), Q, st);
    const bool al = aligned16(A) && aligned16(B) && aligned16(C) && (lda % 4 == 0) && (ldb % 4 == 0) && (ldc % 4 == 0);
    if (kmajor && Q <= 16 && !accumulate) {
        const bool vec = al && (K % 4 == 0);
        for (int q0 = 0; q0 < Q; q0 += 8) {
            const int qn = std::min(8, Q - q0);
            const int blocks = (P + 7) / 8;   // 8 warps per block
            if (vec) k_gemv_rows<true><<<blocks, 256, 0, st>>>(P, qn, K, A, lda, B + (size_t)q0 * ldb, ldb, C + (size_t)q0 * ldc, ldc);
            else     k_gemv_rows<false><<<blocks, 256, 0, st>>>(P, qn, K, A, lda, B + (size_t)q0 * ldb, ldb, C + (size_t)q0 * ldc, ldc);
            ++*launches;
        }
        return cudaGetLastError();
    }
    dim3 grid((P + SG_BM - 1) / SG_BM, (Q + SG_BN - 1) / SG_BN);
    if (kmajor) {
        const bool vec = al && (K % 4 == 0) && (P % 4 == 0);
        if (vec) k_sgemm<true, true><<<grid, 256, 0, st>>>(P, Q, K, A, lda, B, ldb, C, ldc, accumulate);
        else     k_sgemm<true, false><<<grid, 256, 0, st>>>(P, Q, K, A, lda, B, ldb, C, ldc, accumulate);
    } else {
        const bool vec = al && (P % 4 == 0) && (Q % 4 == 0);
        if (vec) k_sgemm<false, true><<<grid, 256, 0, st>>>(P, Q, K, A, lda, B, ldb, C, ldc, accumulate);
        else     k_sgemm<false, false><<<grid, 256, 0, st>>>(P, Q, K, A, lda, B, ldb, C, ldc, accumulate);
    }
    ++*launches;
    return cudaGetLastError();
}

cudaError_t outer_acc_launch(int N, int Bq, const float* g, int ldg, const float* src, int lds, float* dWraw, int ldw, cudaStream_t st) {
    dim3 og((N + 255) / 256, N);
    k_outer_acc<<<og, 256, 0, st>>>(N, Bq, g, ldg, src, lds, dWraw, ldw);
    return cudaGetLastError();
}

}  // namespace rp
