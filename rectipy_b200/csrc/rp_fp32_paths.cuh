// Host launchers of the fp32 execution paths that live in their own translation unit (rp_fp32_paths.cu): the persistent few-trial
// kernels and the FFMA contractions.  Why a second unit: the engine's main unit (rp_api.cu, ~1 min with -split-compile) and these
// latency-bound kernels do not mix -- with split compilation the register allocation of k_persist_fwd / k_persist_bwd came out at 164
// or at 254 registers depending on what ELSE the unit contained (unrelated edits to rp_api.cu moved C1 between 2.6 and 3.3 us/step and
// the 32-trial sweep between 7.8 and 11.1 us/step, profiles/r2_split_compile.md); compiled alone and unsplit the result is stable.
#pragma once
#include <cuda_runtime.h>

namespace rp {

struct PersistFwdArgs;
struct PersistBwdArgs;

// 0 ok; 1 CUDA error (*err receives it); 2 the grid is not co-resident (occ_out / sms_out filled); 3 unknown model id
int ps_launch_fwd(int model, const PersistFwdArgs* pa, int grid, size_t smem, cudaStream_t st, cudaError_t* err, int* occ_out, int* sms_out);
int ps_launch_bwd(int model, const PersistBwdArgs* pa, int grid, size_t smem, cudaStream_t st, cudaError_t* err, int* occ_out, int* sms_out);

// C[q*ldc+p] (+)= sum_k Aop(p,k) Bop(q,k) on the FFMA / GEMV kernels; returns the cudaError_t of the launch, adds to *launches
cudaError_t gemm_fp32_launch(bool kmajor, int P, int Q, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc, int accumulate,
                             cudaStream_t st, long long* launches);
// dWraw[i][j] += sum_b g[b][i] src[b][j]  (few trials)
cudaError_t outer_acc_launch(int N, int Bq, const float* g, int ldg, const float* src, int lds, float* dWraw, int ldw, cudaStream_t st);

}  // namespace rp
