// Plain-data argument records of the per-step kernels and the two device helpers that read them -- shared between the engine
// (nvcc, included by rp_kernels.cuh) and the vector fields compiled at run time from user templates (NVRTC, rectipy_b200/jit.py).
// No includes: `rectipy_b200.h` (constants) must precede this header; only built-in CUDA types are used.
#pragma once

namespace rp {

struct ModelParams {
    const float* p[RP_NUM_PARAMS];
    int stride[RP_NUM_PARAMS];   // neuron stride: 0 shared, 1 per neuron
    int bstride[RP_NUM_PARAMS];  // trial stride: 0 shared across trials, 1 per trial ([B]), N per trial and neuron ([B][N])
};

// parameter `which` of neuron i in trial b (parameter sweeps give every trial its own value)
__device__ __forceinline__ float ldp(const ModelParams& mp, int which, int i, int b = 0) {
    return __ldg(mp.p[which] + (size_t)b * mp.bstride[which] + (size_t)i * mp.stride[which]);
}


struct ScaleRef {
    const float* bound;   // device scalar: an upper bound (or the exact maximum) of |x| over the operand; nullptr: unscaled
    float add;            // added to *bound (growth allowance of one step, e.g. +1 for a synapse that gains one spike)
    int H;                // the bound is mapped into [2^(H-1), 2^H)
};

struct FwdStepArgs {
    int N, B, m, in_mode, in_target;
    float dt, theta, v_reset;
    const float* y_cur;   // [nsv][B][N]
    float* y_next;        // [nsv][B][N]
    const float* u;       // [B][ldu]  recurrent drive  (k_i W) . src_t
    int ldu;
    const float* x_t;     // dense: [B][N]   proj: [B][m]
    const float* W_in;    // [N][m]
    ModelParams mp;
    float* src_next;      // rate models, fp32 path: act(v_{t+1}) [B][N], else nullptr
    void* src_hi;         // tensor-core path: source operand of the next step, split, [B][ld_src] (fp32 words or binary16)
    void* src_lo;
    int ld_src;
    ScaleRef sc_out;      // binary16 operands: scale of the operand written here; amax_out receives max |src_{t+1}|
    float* amax_out;
    float* urec_out;      // ik: checkpoint plane receiving the recurrent drive of this step [B][N], or nullptr
    const float2* mf;     // iku: per trial {mean_i v_t, mean_i spike_t} of the state being stepped (k_trial_means), else nullptr
    int per_trial;        // 1: some parameter differs between trials (no per-neuron hoisting)
    int no_lean;          // 1: fused forward epilogue uses the generic element loop instead of the lean one (A/B runs)
};


__device__ __forceinline__ float input_current(int in_mode, int m, const float* __restrict__ x_t,
                                               const float* __restrict__ W_in, int N, int b, int i) {
    if (in_mode == RP_IN_DENSE) return __ldg(x_t + (size_t)b * N + i);
    if (in_mode == RP_IN_PROJ) {
        float acc = 0.f;
        for (int j = 0; j < m; ++j) acc = fmaf(__ldg(W_in + (size_t)i * m + j), __ldg(x_t + (size_t)b * m + j), acc);
        return acc;
    }
    return 0.f;
}


struct AdjArgs {
    int N, B, m, k, in_mode, in_target, out_mode, out_var;
    float dt, theta, slope;
    int do_post, do_pre, zero_after_post;   // zero_after_post: truncated-BPTT cut between step t-1 and t
    const float* y_t;      // history slot t      [nsv][B][N]  (post)
    const float* y_tm1;    // history slot t-1                 (pre)
    const float* urec_t;   // ik: recurrent drive of step t [B][N] (checkpoint plane), else nullptr
    const float2* mf_t;    // iku: per trial {mean v_t, mean spike_t}
    const float2* asum;    // iku: per trial {mean_i(ax_i dt b_i / tau_u_i), mean_i(ax_i kappa_i)} of the incoming adjoint of u
    float* adj;            // [nsv][B][N] adjoint of the state, updated in place (t+1 -> t)
    const float* Z;        // [B][ldz]  (kW)^T g_t
    int ldz;
    const float* x_t;      // input of step t (dense [B][N] | proj [B][m])
    const float* W_in;
    const float* W_out;
    const float* e_t;      // dL/d out_rec[j] for the record window that contains t: [B][k] | [B][N], or nullptr
    float e_scale;         // 1/|window|
    ModelParams mp;
    float* g;              // [B][N]   operand for the next GEMMs (fp32 path)
    float* src;            // [B][N]   rate models: act(v_{t-1}) for the weight gradient (fp32 path), else nullptr
    float* g_amax;         // optional: receives max |g_{t-1}| (binary16 operand scaling), accumulated with atomicMax
    float* g_hi; float* g_lo; int ld_g;           // 3xTF32 operands  [B][ld_g]
    float* gT_hi; float* gT_lo;                   // transposed copies [N][ld_t] for the weight gradient
    float* srcT_hi; float* srcT_lo; int ld_t; int t_col0;  // column offset (trial index base) inside the K-chunk
    float* dparams[RP_NUM_PARAMS];  // [N] accumulators or nullptr
    float* dW_in;          // [N][m] or nullptr
    float* dW_out;         // [k][N] or nullptr
    float* g_x_t;          // dense input gradient of step t [B][N] or nullptr
    int any_param_grad;
    int per_trial;         // 1: some parameter differs between trials (no per-neuron hoisting)
    // run-time compiled MultiSpikeResetNet fields (outputs are POST-update slices, rectipy/nodes.py:451-465): dL/d out_rec of the record
    // window that contains step t-1, added to the adjoint of y_t in the launch that finishes step t; y_t is set on every launch
    const float* e_tm1;
    float e_scale_tm1;
};


// run-time compiled forward step with the recurrent contraction inside (few trials: one warp per neuron row, one launch per step)
struct JitRowsArgs {
    FwdStepArgs a;
    const float* W;        // [N][ldw]  recurrent weights (coupling constants live in the equations)
    int ldw;
    const float* src;      // [B][N]    source values of the step
};

}  // namespace rp
