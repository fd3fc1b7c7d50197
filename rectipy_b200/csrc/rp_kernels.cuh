// Element-wise kernels of the RectiPy time-stepping engine (sm_100a).
//
// Everything that is not a contraction lives here: the Euler step of the template's vector field with
// spike threshold/reset (nodes.py:166-170,382-392; neuron_model_templates/**.yaml), the input projection
// (edges.py:48-49, m <= RP_MAX_IN), the Observer window means (network.py:590-597), and the reverse-time
// adjoint of all of it including the surrogate spike gradient (nodes.py:478-481).  The contractions
// (W.r, W^T.lambda, lambda (x) r) are in rp_gemm_simt.cuh / rp_gemm_tc.cuh.
//
// Layout: state planes y[var][trial][neuron] (plane stride B*N, neuron fastest) so that a warp touches
// 128 contiguous bytes per access.  All kernels are HBM/L2-bandwidth bound streaming kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/rectipy_b200.h"

namespace rp {

struct ModelParams {
    const float* p[RP_NUM_PARAMS];
    int stride[RP_NUM_PARAMS];   // 0: shared scalar, 1: per neuron
};

__device__ __forceinline__ float ldp(const ModelParams& mp, int which, int i) {
    return __ldg(mp.p[which] + (size_t)i * mp.stride[which]);
}

template <int MODEL> struct ModelTraits;
template <> struct ModelTraits<RP_LI_TANH>    { static constexpr int NSV = 1; static constexpr bool SPIKING = false; };
template <> struct ModelTraits<RP_LI_SIGMOID> { static constexpr int NSV = 1; static constexpr bool SPIKING = false; };
template <> struct ModelTraits<RP_QIF>        { static constexpr int NSV = 2; static constexpr bool SPIKING = true; };
template <> struct ModelTraits<RP_QIF_SFA>    { static constexpr int NSV = 3; static constexpr bool SPIKING = true; };
template <> struct ModelTraits<RP_LIF>        { static constexpr int NSV = 2; static constexpr bool SPIKING = true; };

// activation of the rate templates (leaky_integrator.yaml:20-36) and its derivative w.r.t. v
template <int MODEL>
__device__ __forceinline__ float rate_act(const ModelParams& mp, int i, float v) {
    if constexpr (MODEL == RP_LI_TANH) {
        return tanhf(v);
    } else {
        float rmax = ldp(mp, RP_P_RMAX, i), s = ldp(mp, RP_P_SIG_S, i), v0 = ldp(mp, RP_P_V0, i);
        return rmax / (1.0f + expf(s * (v0 - v)));
    }
}
template <int MODEL>
__device__ __forceinline__ float rate_act_grad(const ModelParams& mp, int i, float v) {
    if constexpr (MODEL == RP_LI_TANH) {
        float r = tanhf(v);
        return 1.0f - r * r;
    } else {
        float rmax = ldp(mp, RP_P_RMAX, i), s = ldp(mp, RP_P_SIG_S, i), v0 = ldp(mp, RP_P_V0, i);
        float r = rmax / (1.0f + expf(s * (v0 - v)));
        return s * r * (1.0f - r / rmax);
    }
}

// TF32 split used by the 3xTF32 tensor-core path: x ~= hi + lo with hi, lo exactly representable in tf32
// (round-to-nearest on both parts: |x - hi - lo| <= 2^-24 |x|, and the dropped lo*lo product is <= 2^-24 relative)
__device__ __forceinline__ float round_tf32(float x) {
    uint32_t u;
    asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(u) : "f"(x));
    return __uint_as_float(u);
}
__device__ __forceinline__ void split_tf32(float x, float& hi, float& lo) {
    hi = round_tf32(x);
    lo = round_tf32(x - hi);
}

// ------------------------------------------------------------------------------------------------------
// forward step
// ------------------------------------------------------------------------------------------------------
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
    float* src_hi;        // 3xTF32 path: source operand of the next step, split, [B][ld_src]
    float* src_lo;
    int ld_src;
};

template <int MODEL>
__device__ __forceinline__ void fwd_elem(const FwdStepArgs& a, int i, float u, float Iin,
                                         float v, float s, float x, float& v1, float& s1, float& x1) {
    const float dt = a.dt;
    const float tau = ldp(a.mp, RP_P_TAU, i), eta = ldp(a.mp, RP_P_ETA, i);
    if constexpr (!ModelTraits<MODEL>::SPIKING) {
        // li_op: v' = -v/tau + k*r_in + I_ext + eta          (leaky_integrator.yaml:10)
        v1 = v + dt * (-v / tau + u + Iin + eta);
        s1 = 0.f; x1 = 0.f;
    } else {
        const float tau_s = ldp(a.mp, RP_P_TAU_S, i);
        const bool p = v >= a.theta;                           // heaviside(v-theta, 1.0)   nodes.py:383,476
        const float pf = p ? 1.0f : 0.0f;
        float vt;
        if constexpr (MODEL == RP_LIF) {
            // lif_op: v' = -v/tau + k*s_in + I_ext + eta ; s' = -s/tau_s + spike + s_ext   (lif.yaml:10-15)
            const float Iv = a.in_target == 0 ? Iin : 0.f, Is = a.in_target == 1 ? Iin : 0.f;
            vt = v + dt * (-v / tau + u + Iv + eta);
            s1 = s + dt * (-s / tau_s + Is) + pf;
            x1 = 0.f;
        } else {
            // qif_op: v' = (v^2 + eta + I_ext)/tau + k*s_in ; s' = -s/tau_s + spike        (qif.yaml:10-12)
            float xx = 0.f;
            if constexpr (MODEL == RP_QIF_SFA) xx = x;         // eta -> eta - x                 (qif.yaml:28-29)
            vt = v + dt * ((v * v + eta - xx + Iin) / tau + u);
            s1 = s + dt * (-s / tau_s) + pf;                  // dt * (spike/dt) == 1 per spike  nodes.py:385
            if constexpr (MODEL == RP_QIF_SFA) {
                const float tau_x = ldp(a.mp, RP_P_TAU_X, i), alpha = ldp(a.mp, RP_P_ALPHA, i);
                x1 = x + dt * (-x / tau_x) + alpha * pf;      // x' = -x/tau_x + alpha*spike     (qif.yaml:31)
            } else {
                x1 = 0.f;
            }
        }
        v1 = p ? a.v_reset : vt;                               // reset blend                     nodes.py:390
    }
}

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

template <int MODEL>
__global__ void __launch_bounds__(256) k_fwd_step(FwdStepArgs a) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    const size_t plane = (size_t)a.B * a.N;
    const size_t total = plane;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / a.N), i = (int)(idx - (size_t)b * a.N);
        const float v = a.y_cur[idx];
        const float s = NSV > 1 ? a.y_cur[plane + idx] : 0.f;
        const float x = NSV > 2 ? a.y_cur[2 * plane + idx] : 0.f;
        const float u = a.u[(size_t)b * a.ldu + i];
        const float Iin = input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i);
        float v1, s1, x1;
        fwd_elem<MODEL>(a, i, u, Iin, v, s, x, v1, s1, x1);
        a.y_next[idx] = v1;
        if (NSV > 1) a.y_next[plane + idx] = s1;
        if (NSV > 2) a.y_next[2 * plane + idx] = x1;
        float src1;
        if constexpr (ModelTraits<MODEL>::SPIKING) src1 = s1; else src1 = rate_act<MODEL>(a.mp, i, v1);
        if (a.src_next) a.src_next[idx] = src1;
        if (a.src_hi) {
            float hi, lo;
            split_tf32(src1, hi, lo);
            a.src_hi[(size_t)b * a.ld_src + i] = hi;
            a.src_lo[(size_t)b * a.ld_src + i] = lo;
        }
    }
}

// source operand of the very first step (and of re-started runs): src_0 from y_0
template <int MODEL>
__global__ void __launch_bounds__(256) k_init_src(int N, int B, const float* y, ModelParams mp,
                                                   float* src, int ld_plain, float* src_hi, float* src_lo, int ld_src) {
    const size_t plane = (size_t)B * N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / N), i = (int)(idx - (size_t)b * N);
        float r;
        if constexpr (ModelTraits<MODEL>::SPIKING) r = y[plane + idx]; else r = rate_act<MODEL>(mp, i, y[idx]);
        if (src) src[(size_t)b * ld_plain + i] = r;
        if (src_hi) {
            float hi, lo;
            split_tf32(r, hi, lo);
            src_hi[(size_t)b * ld_src + i] = hi;
            src_lo[(size_t)b * ld_src + i] = lo;
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// Observer: window means of the output and instantaneous state records     (network.py:590-597)
// ------------------------------------------------------------------------------------------------------
struct ObsArgs {
    int N, B, k, out_mode, out_var, model;
    const float* y_pre;     // state before the step (slot t)       [nsv][B][N]
    const float* y_post;    // state after the step  (slot t+1)
    const float* W_out;     // [k][N]
    ModelParams mp;
    float* win_acc;         // [B][N] running window sum of y[out] (only used when the window is longer than 1)
    int win_first;          // 1: this step opens a window (ignore win_acc on read)
    int win_close;          // 1: this step is a record step
    float inv_len;          // 1/|window|
    float* out_rec_j;       // [B][k] | [B][N]  record slot j, or nullptr
    int n_rec_vars;
    int rec_var[RP_MAX_REC];
    int rec_reduce[RP_MAX_REC];
    float* rec_buf_j[RP_MAX_REC];   // [B][N] | [B]
    int rec_post;           // 1: recorded vars are post-update (RateNet), 0: pre-update (SpikeResetNet)
};

template <int MODEL>
__device__ __forceinline__ float out_value(const ObsArgs& a, const float* y, size_t plane, int b, int i) {
    const size_t idx = (size_t)b * a.N + i;
    if (a.out_var == RP_VAR_R) {
        if constexpr (!ModelTraits<MODEL>::SPIKING) return rate_act<MODEL>(a.mp, i, y[idx]);
        else return 0.f;
    }
    return y[(size_t)a.out_var * plane + idx];
}

__device__ __forceinline__ float block_sum(float v, float* red) {
    // blockDim.x == 256
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    const int w = threadIdx.x >> 5, l = threadIdx.x & 31;
    __syncthreads();
    if (l == 0) red[w] = v;
    __syncthreads();
    float t = (threadIdx.x < 8) ? red[threadIdx.x] : 0.f;
    if (w == 0) {
        for (int o = 4; o > 0; o >>= 1) t += __shfl_xor_sync(0xffffffffu, t, o);
        if (l == 0) red[8] = t;
    }
    __syncthreads();
    return red[8];
}

// one block per trial b
template <int MODEL>
__global__ void __launch_bounds__(256) k_observe(ObsArgs a) {
    __shared__ float red[9];
    const int b = blockIdx.x;
    const size_t plane = (size_t)a.B * a.N;
    float part[RP_MAX_OUT];
#pragma unroll
    for (int q = 0; q < RP_MAX_OUT; ++q) part[q] = 0.f;
    const bool need_out = a.out_rec_j != nullptr || !a.win_close;
    if (need_out) {
        for (int i = threadIdx.x; i < a.N; i += blockDim.x) {
            const size_t idx = (size_t)b * a.N + i;
            float acc = out_value<MODEL>(a, a.y_pre, plane, b, i);
            if (!a.win_first) acc += a.win_acc[idx];
            if (a.win_close) {
                if (a.out_mode == RP_OUT_DENSE) {
                    if (a.out_rec_j) a.out_rec_j[idx] = acc * a.inv_len;
                } else {
#pragma unroll
                    for (int q = 0; q < RP_MAX_OUT; ++q)
                        if (q < a.k) part[q] = fmaf(__ldg(a.W_out + (size_t)q * a.N + i), acc, part[q]);
                }
            } else {
                a.win_acc[idx] = acc;
            }
        }
        if (a.win_close && a.out_mode == RP_OUT_READOUT && a.out_rec_j) {
            for (int q = 0; q < a.k; ++q) {
                const float tot = block_sum(part[q], red);
                if (threadIdx.x == 0) a.out_rec_j[(size_t)b * a.k + q] = tot * a.inv_len;
            }
        }
    }
    if (a.win_close) {
        const float* ysrc = a.rec_post ? a.y_post : a.y_pre;
        for (int r = 0; r < a.n_rec_vars; ++r) {
            const float* pl = ysrc + (size_t)a.rec_var[r] * plane + (size_t)b * a.N;
            if (a.rec_reduce[r]) {
                float acc = 0.f;
                for (int i = threadIdx.x; i < a.N; i += blockDim.x) acc += pl[i];
                const float tot = block_sum(acc, red);
                if (threadIdx.x == 0) a.rec_buf_j[r][b] = tot / (float)a.N;
            } else {
                float* dst = a.rec_buf_j[r] + (size_t)b * a.N;
                for (int i = threadIdx.x; i < a.N; i += blockDim.x) dst[i] = pl[i];
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// adjoint step (reverse time).  One launch finishes step t ("post": needs Z_t = (kW)^T g_t) and prepares
// step t-1 ("pre": g_{t-1} = dt * gate_{t-1} * a_t, plus the source operand of the weight gradient).
// Recurrences: SURVEY.md Appendix A (verified against the reference's autograd).
// ------------------------------------------------------------------------------------------------------
struct AdjArgs {
    int N, B, m, k, in_mode, in_target, out_mode, out_var;
    float dt, theta, slope;
    int do_post, do_pre, zero_after_post;   // zero_after_post: truncated-BPTT cut between step t-1 and t
    const float* y_t;      // history slot t      [nsv][B][N]  (post)
    const float* y_tm1;    // history slot t-1                 (pre)
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
    float* g_hi; float* g_lo; int ld_g;           // 3xTF32 operands  [B][ld_g]
    float* gT_hi; float* gT_lo;                   // transposed copies [N][ld_t] for the weight gradient
    float* srcT_hi; float* srcT_lo; int ld_t; int t_col0;  // column offset (trial index base) inside the K-chunk
    float* dparams[RP_NUM_PARAMS];  // [N] accumulators or nullptr
    float* dW_in;          // [N][m] or nullptr
    float* dW_out;         // [k][N] or nullptr
    float* g_x_t;          // dense input gradient of step t [B][N] or nullptr
    int any_param_grad;
};

constexpr int ADJ_TX = 32, ADJ_TY = 8, ADJ_BPT = 8;   // block covers 32 neurons x 64 trials

template <int MODEL>
__global__ void __launch_bounds__(ADJ_TX * ADJ_TY) k_adj_step(AdjArgs a) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr bool SPK = ModelTraits<MODEL>::SPIKING;
    constexpr int NACC = RP_NUM_PARAMS + RP_MAX_IN + RP_MAX_OUT;
    __shared__ float red[ADJ_TY][ADJ_TX + 1];
    const int i = blockIdx.x * ADJ_TX + threadIdx.x;
    const int b0 = blockIdx.y * (ADJ_TY * ADJ_BPT) + threadIdx.y;
    const size_t plane = (size_t)a.B * a.N;
    const bool valid_i = i < a.N;

    float acc[NACC];
#pragma unroll
    for (int q = 0; q < NACC; ++q) acc[q] = 0.f;

    float tau = 1.f, tau_s = 1.f, tau_x = 1.f, alpha = 0.f;
    if (valid_i) {
        tau = ldp(a.mp, RP_P_TAU, i);
        if (SPK) tau_s = ldp(a.mp, RP_P_TAU_S, i);
        if (MODEL == RP_QIF_SFA) { tau_x = ldp(a.mp, RP_P_TAU_X, i); alpha = ldp(a.mp, RP_P_ALPHA, i); }
    }
    const float dt = a.dt;

    for (int l = 0; l < ADJ_BPT; ++l) {
        const int b = b0 + l * ADJ_TY;
        if (!valid_i || b >= a.B) continue;
        const size_t idx = (size_t)b * a.N + i;
        float av = a.adj[idx];
        float as = NSV > 1 ? a.adj[plane + idx] : 0.f;
        float ax = NSV > 2 ? a.adj[2 * plane + idx] : 0.f;

        if (a.do_post) {
            const float v = a.y_t[idx];
            const float s = NSV > 1 ? a.y_t[plane + idx] : 0.f;
            const float x = NSV > 2 ? a.y_t[2 * plane + idx] : 0.f;
            const float Z = a.Z[(size_t)b * a.ldz + i];
            // readout / record gradient flowing into y_t[out]
            float ro = 0.f, yout = 0.f;
            if (a.e_t) {
                if (a.out_var == RP_VAR_V) yout = v; else if (a.out_var == RP_VAR_S) yout = s;
                else if (a.out_var == RP_VAR_X) yout = x;
                else { if constexpr (!SPK) yout = rate_act<MODEL>(a.mp, i, v); }
                if (a.out_mode == RP_OUT_DENSE) {
                    ro = a.e_t[idx] * a.e_scale;
                } else {
#pragma unroll
                    for (int q = 0; q < RP_MAX_OUT; ++q) {
                        if (q < a.k) {
                            const float e = __ldg(a.e_t + (size_t)b * a.k + q) * a.e_scale;
                            ro = fmaf(__ldg(a.W_out + (size_t)q * a.N + i), e, ro);
                            acc[RP_NUM_PARAMS + RP_MAX_IN + q] = fmaf(e, yout, acc[RP_NUM_PARAMS + RP_MAX_IN + q]);
                        }
                    }
                }
            }
            float dI = 0.f;      // dL/d(input current of step t)
            float nav, nas = 0.f, nax = 0.f;
            if constexpr (!SPK) {
                const float rg = rate_act_grad<MODEL>(a.mp, i, v);
                nav = av * (1.0f - dt / tau) + rg * Z;
                if (a.out_var == RP_VAR_V) nav += ro; else if (a.out_var == RP_VAR_R) nav += rg * ro;
                dI = dt * av;
                acc[RP_P_ETA] += dt * av;
                acc[RP_P_TAU] += dt * av * v / (tau * tau);
            } else {
                const bool p = v >= a.theta;
                const float gv = p ? 0.f : av;
                const float d = 1.0f + a.slope * fabsf(v - a.theta);
                const float sg = 1.0f / (d * d);                 // Spike.backward          nodes.py:478-481
                if constexpr (MODEL == RP_LIF) {
                    nav = gv * (1.0f - dt / tau) + sg * as;
                    nas = as * (1.0f - dt / tau_s) + Z;
                    dI = a.in_target == 0 ? dt * gv : dt * as;
                    acc[RP_P_ETA] += dt * gv;
                    acc[RP_P_TAU] += dt * gv * v / (tau * tau);
                    acc[RP_P_TAU_S] += as * s * dt / (tau_s * tau_s);
                } else {
                    const float Iin = a.dparams[RP_P_TAU] ? input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i) : 0.f;
                    const float eta = ldp(a.mp, RP_P_ETA, i);
                    nav = gv * (1.0f + 2.0f * dt * v / tau) + sg * (as + alpha * ax);
                    nas = as * (1.0f - dt / tau_s) + Z;
                    dI = dt / tau * gv;
                    acc[RP_P_ETA] += dI;
                    acc[RP_P_TAU] -= dt * gv * (v * v + eta - x + Iin) / (tau * tau);
                    acc[RP_P_TAU_S] += as * s * dt / (tau_s * tau_s);
                    if constexpr (MODEL == RP_QIF_SFA) {
                        nax = ax * (1.0f - dt / tau_x) - dI;
                        acc[RP_P_TAU_X] += ax * x * dt / (tau_x * tau_x);
                        acc[RP_P_ALPHA] += ax * (p ? 1.0f : 0.0f);
                    }
                }
                if (a.out_var == RP_VAR_V) nav += ro; else if (a.out_var == RP_VAR_S) nas += ro; else if (a.out_var == RP_VAR_X) nax += ro;
            }
            if (a.in_mode == RP_IN_PROJ && a.dW_in) {
#pragma unroll
                for (int j = 0; j < RP_MAX_IN; ++j)
                    if (j < a.m) acc[RP_NUM_PARAMS + j] = fmaf(dI, __ldg(a.x_t + (size_t)b * a.m + j), acc[RP_NUM_PARAMS + j]);
            }
            if (a.g_x_t) a.g_x_t[idx] = dI;
            av = nav; as = nas; ax = nax;
            if (a.zero_after_post) { av = 0.f; as = 0.f; ax = 0.f; }
            a.adj[idx] = av;
            if (NSV > 1) a.adj[plane + idx] = as;
            if (NSV > 2) a.adj[2 * plane + idx] = ax;
        }

        if (a.do_pre) {
            const float vm = a.y_tm1[idx];
            float gate = 1.0f, srcv;
            if constexpr (SPK) { gate = (vm >= a.theta) ? 0.f : 1.0f; srcv = a.y_tm1[plane + idx]; }
            else srcv = rate_act<MODEL>(a.mp, i, vm);
            const float g = dt * gate * av;
            if (a.g) a.g[idx] = g;
            if (a.src) a.src[idx] = srcv;
            if (a.g_hi) {
                float hi, lo;
                split_tf32(g, hi, lo);
                a.g_hi[(size_t)b * a.ld_g + i] = hi;
                a.g_lo[(size_t)b * a.ld_g + i] = lo;
                if (a.gT_hi) {
                    a.gT_hi[(size_t)i * a.ld_t + a.t_col0 + b] = hi;
                    a.gT_lo[(size_t)i * a.ld_t + a.t_col0 + b] = lo;
                    split_tf32(srcv, hi, lo);
                    a.srcT_hi[(size_t)i * a.ld_t + a.t_col0 + b] = hi;
                    a.srcT_lo[(size_t)i * a.ld_t + a.t_col0 + b] = lo;
                }
            }
        }
    }

    if (a.do_post && a.any_param_grad) {
        // reduce the per-thread partial sums over the 8 trial lanes, then one atomic per (neuron, quantity)
#pragma unroll
        for (int q = 0; q < NACC; ++q) {
            float* dst = nullptr;
            if (q < RP_NUM_PARAMS) dst = a.dparams[q];
            else if (q < RP_NUM_PARAMS + RP_MAX_IN) { if (a.dW_in && (q - RP_NUM_PARAMS) < a.m) dst = a.dW_in; }
            else { if (a.dW_out && (q - RP_NUM_PARAMS - RP_MAX_IN) < a.k && a.out_mode == RP_OUT_READOUT) dst = a.dW_out; }
            if (dst == nullptr) continue;     // uniform across the block
            __syncthreads();
            red[threadIdx.y][threadIdx.x] = acc[q];
            __syncthreads();
            if (threadIdx.y == 0 && valid_i) {
                float t = 0.f;
#pragma unroll
                for (int r = 0; r < ADJ_TY; ++r) t += red[r][threadIdx.x];
                size_t off;
                if (q < RP_NUM_PARAMS) off = i;
                else if (q < RP_NUM_PARAMS + RP_MAX_IN) off = (size_t)i * a.m + (q - RP_NUM_PARAMS);
                else off = (size_t)(q - RP_NUM_PARAMS - RP_MAX_IN) * a.N + i;
                atomicAdd(dst + off, t);
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------------
// weight preparation / finalisation
// ------------------------------------------------------------------------------------------------------
// Wk[i][j] = k_i * W[i][j]  and its transpose; optionally also the tf32 hi/lo splits (padded leading dims)
__global__ void __launch_bounds__(256) k_prepare_weights(int N, const float* __restrict__ W, const float* __restrict__ kp,
                                                          int k_stride, float* Wk, float* WkT, int ldw,
                                                          float* Wk_hi, float* Wk_lo, float* WkT_hi, float* WkT_lo) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;     // bx: column (j) base, by: row (i) base
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;    // 32 x 8
    for (int r = ty; r < 32; r += 8) {
        const int i = by + r, j = bx + tx;
        float w = 0.f;
        if (i < N && j < N) w = W[(size_t)i * N + j] * __ldg(kp + (size_t)i * k_stride);
        tile[r][tx] = w;
        if (i < N && j < N) {
            if (Wk) Wk[(size_t)i * ldw + j] = w;
            if (Wk_hi) { float hi, lo; split_tf32(w, hi, lo); Wk_hi[(size_t)i * ldw + j] = hi; Wk_lo[(size_t)i * ldw + j] = lo; }
        }
    }
    __syncthreads();
    for (int r = ty; r < 32; r += 8) {
        const int j = bx + r, i = by + tx;       // transposed element (j, i)
        if (i < N && j < N) {
            const float w = tile[tx][r];
            if (WkT) WkT[(size_t)j * ldw + i] = w;
            if (WkT_hi) { float hi, lo; split_tf32(w, hi, lo); WkT_hi[(size_t)j * ldw + i] = hi; WkT_lo[(size_t)j * ldw + i] = lo; }
        }
    }
}

// dW[i][j] = k_i * dWraw[i][j] ;  dk[i] = sum_j dWraw[i][j] * W[i][j]        (one block per row)
__global__ void __launch_bounds__(256) k_finish_wgrad(int N, const float* __restrict__ dWraw, int ldr, const float* __restrict__ W,
                                                       const float* __restrict__ kp, int k_stride, float* dW, float* dk) {
    __shared__ float red[9];
    const int i = blockIdx.x;
    const float kv = __ldg(kp + (size_t)i * k_stride);
    float acc = 0.f;
    for (int j = threadIdx.x; j < N; j += blockDim.x) {
        const float r = dWraw[(size_t)i * ldr + j];
        if (dW) dW[(size_t)i * N + j] = kv * r;
        acc = fmaf(r, W[(size_t)i * N + j], acc);
    }
    if (dk) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) dk[i] = tot;
    }
}

__global__ void __launch_bounds__(256) k_fill(float* p, size_t n, float v) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) p[idx] = v;
}

}  // namespace rp
