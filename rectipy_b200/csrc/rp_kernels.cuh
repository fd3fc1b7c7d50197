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
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <utility>
#include "../../include/rectipy_b200.h"
#include "rp_jit_abi.cuh"

namespace rp {

// ---- optional CTA timeline (debug / profiles): one record per traced CTA when rp_trace_enable() armed a buffer ---------------
struct TraceRec { unsigned tag, smid; unsigned long long t0, t1; };
// (internal linkage: the headers are compiled into two translation units, see rp_fp32_paths.cu)
static __device__ TraceRec* g_trace_buf = nullptr;
static __device__ unsigned g_trace_cap = 0;
static __device__ unsigned g_trace_n = 0;
enum { TR_GEMM_STORE = 1, TR_GEMM_WGRAD = 2, TR_GEMM_FWD = 3, TR_GEMM_ADJ = 4, TR_ADJ_STEP = 5, TR_ADJ_CONVERT = 6 };
__device__ __forceinline__ unsigned long long trace_now() { unsigned long long t; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t)); return t; }
// call from ONE thread of the CTA; returns nullptr when tracing is off
__device__ __forceinline__ TraceRec* trace_begin(unsigned tag) {
    TraceRec* buf = g_trace_buf;
    if (buf == nullptr) return nullptr;
    const unsigned slot = atomicAdd(&g_trace_n, 1u);
    if (slot >= g_trace_cap) return nullptr;
    unsigned sm;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(sm));
    buf[slot].tag = tag; buf[slot].smid = sm; buf[slot].t0 = trace_now(); buf[slot].t1 = 0;
    return buf + slot;
}
__device__ __forceinline__ void trace_end(TraceRec* r) { if (r) r->t1 = trace_now(); }

// ---- programmatic dependent launch (PDL) -------------------------------------------------------------------------------------
// The kernels of the per-step chains are launched with cudaLaunchAttributeProgrammaticStreamSerialization: a kernel's CTAs may
// start (barrier / TMEM set-up, parameter loads) while the previous kernel of the stream drains; pdl_wait() returns once that
// kernel has completed and its writes are visible, and must precede the first access to anything it produced.
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Opt-in (RP_PDL=1): measured on B200 the inter-kernel gaps shrink from 3-5 us to 0-2.5 us, but the kernels that start early run
// correspondingly longer (adjoint step 35 -> 45 us) and the pass time is unchanged (35.1 vs 35.0 ms) -- the "gaps" are drain and
// ramp time of lock-step single-wave kernels, not launch latency.
inline bool pdl_enabled() { static const bool on = getenv("RP_PDL") != nullptr; return on; }
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = pdl_enabled() ? 1 : 0;
    cfg.attrs = attr; cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// record windows of Network.run (network.py:590-597), usable on host and device
struct PWindow { int j, first, close, len; };
__host__ __device__ inline PWindow pwindow_of(int t, int T, int S, int cutoff) {
    PWindow w{-1, 0, 0, 0};
    if (t < cutoff) return w;
    const int r0 = ((cutoff + S - 1) / S) * S;
    int j, start, rec;
    if (t <= r0) { j = 0; start = cutoff; rec = r0; }
    else { j = (t - r0 + S - 1) / S; rec = r0 + j * S; start = rec - S + 1; }
    if (rec >= T) return w;
    w.j = j; w.first = (t == start); w.close = (t == rec); w.len = rec - start + 1;
    return w;
}

template <int MODEL> struct ModelTraits;
template <> struct ModelTraits<RP_LI_TANH>    { static constexpr int NSV = 1; static constexpr bool SPIKING = false; };
template <> struct ModelTraits<RP_LI_SIGMOID> { static constexpr int NSV = 1; static constexpr bool SPIKING = false; };
template <> struct ModelTraits<RP_QIF>        { static constexpr int NSV = 2; static constexpr bool SPIKING = true; };
template <> struct ModelTraits<RP_QIF_SFA>    { static constexpr int NSV = 3; static constexpr bool SPIKING = true; };
template <> struct ModelTraits<RP_LIF>        { static constexpr int NSV = 2; static constexpr bool SPIKING = true; };
template <> struct ModelTraits<RP_IK>         { static constexpr int NSV = 3; static constexpr bool SPIKING = true; };   // planes v, s, u
template <> struct ModelTraits<RP_IKU>        { static constexpr int NSV = 3; static constexpr bool SPIKING = true; };   // planes v, s, u (mean-field u)
template <> struct ModelTraits<RP_IK_BIEXP>   { static constexpr int NSV = 4; static constexpr bool SPIKING = true; };   // planes v, s, u (mean-field), x (synaptic rise)
__host__ __device__ constexpr bool is_ik(int model) { return model == RP_IK || model == RP_IKU || model == RP_IK_BIEXP; }
// templates whose recovery variable is driven by the per-trial population means (k_trial_means / k_trial_adj_sums every step)
__host__ __device__ constexpr bool is_mean_field(int model) { return model == RP_IKU || model == RP_IK_BIEXP; }
// checkpoint planes per step: the ik conductance synapse makes d(v')/dv depend on the recurrent drive, which is stored too
template <int MODEL> struct HistPlanes { static constexpr int N = ModelTraits<MODEL>::NSV + (is_ik(MODEL) ? 1 : 0); };
// parameter slot of the coupling constant that is folded into the weights
__host__ __device__ constexpr int fold_slot(int model) { return is_ik(model) ? RP_P_G : RP_P_K; }

// activation of the rate templates (leaky_integrator.yaml:20-36) and its derivative w.r.t. v
template <int MODEL>
__device__ __forceinline__ float rate_act(const ModelParams& mp, int i, float v, int b = 0) {
    if constexpr (MODEL == RP_LI_TANH) {
        return tanhf(v);
    } else {
        float rmax = ldp(mp, RP_P_RMAX, i, b), s = ldp(mp, RP_P_SIG_S, i, b), v0 = ldp(mp, RP_P_V0, i, b);
        return rmax / (1.0f + expf(s * (v0 - v)));
    }
}
template <int MODEL>
__device__ __forceinline__ float rate_act_grad(const ModelParams& mp, int i, float v, int b = 0) {
    if constexpr (MODEL == RP_LI_TANH) {
        float r = tanhf(v);
        return 1.0f - r * r;
    } else {
        float rmax = ldp(mp, RP_P_RMAX, i, b), s = ldp(mp, RP_P_SIG_S, i, b), v0 = ldp(mp, RP_P_V0, i, b);
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

// ---- binary16 split (3xFP16 tensor-core path) ------------------------------------------------------------------------------
// binary16 has the same 11-bit significand as tf32, so hi = rn16(x), lo = rn16(x - hi) represents x to 2^-22 |x| as long as
// both parts are normal binary16 numbers; below that the error is the absolute 2^-25 of the subnormal grid.  Every operand is
// therefore multiplied by an exact power of two that puts its largest magnitude near the top of the binary16 range (65504):
// the worst absolute error is then < 2^-37 of the operand's maximum -- far below fp32 rounding of the contraction itself.
// The scale is never stored: producer and consumer both derive the exponent from the same device-resident bound.
__host__ __device__ inline ScaleRef no_scale() { return ScaleRef{nullptr, 0.f, 0}; }
__device__ __forceinline__ int expo_for(float bound, int H) {
    if (!(bound > 0.f) || bound > 3.0e38f) return 0;
    // exponent field instead of ilogbf(): identical for normal numbers, and a subnormal bound clamps to +100 either way
    const int e = H - 1 - (int)((__float_as_uint(bound) >> 23) & 0xffu) + 127;
    return max(-100, min(100, e));
}
__device__ __forceinline__ float exp2i(int e) { return __int_as_float((e + 127) << 23); }       // 2^e, |e| <= 126
__device__ __forceinline__ int scale_expo(const ScaleRef& r) { return r.bound ? expo_for(*r.bound + r.add, r.H) : 0; }
__device__ __forceinline__ void split_f16(float xs, __half& hi, __half& lo) {
    hi = __float2half_rn(xs);
    lo = __float2half_rn(xs - __half2float(hi));
}
// four consecutive elements -> one 8-byte store per part
__device__ __forceinline__ void store_split4_f16(void* hi_base, void* lo_base, size_t elem_off, const float* x, float scale) {
    __half h[4], l[4];
#pragma unroll
    for (int r = 0; r < 4; ++r) split_f16(x[r] * scale, h[r], l[r]);
    uint2 ph, pl;
    ph.x = (uint32_t)__half_as_ushort(h[0]) | ((uint32_t)__half_as_ushort(h[1]) << 16);
    ph.y = (uint32_t)__half_as_ushort(h[2]) | ((uint32_t)__half_as_ushort(h[3]) << 16);
    pl.x = (uint32_t)__half_as_ushort(l[0]) | ((uint32_t)__half_as_ushort(l[1]) << 16);
    pl.y = (uint32_t)__half_as_ushort(l[2]) | ((uint32_t)__half_as_ushort(l[3]) << 16);
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(hi_base) + elem_off) = ph;
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(lo_base) + elem_off) = pl;
}
__device__ __forceinline__ void store_split1_f16(void* hi_base, void* lo_base, size_t elem_off, float x, float scale) {
    __half h, l;
    split_f16(x * scale, h, l);
    reinterpret_cast<__half*>(hi_base)[elem_off] = h;
    reinterpret_cast<__half*>(lo_base)[elem_off] = l;
}
// running maximum of non-negative floats (bit patterns of non-negative floats order like unsigned integers)
__device__ __forceinline__ void atomic_max_nonneg(float* dst, float v) { atomicMax(reinterpret_cast<unsigned int*>(dst), __float_as_uint(v)); }
__device__ __forceinline__ float warp_max(float v) {
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// ------------------------------------------------------------------------------------------------------
// forward step
// ------------------------------------------------------------------------------------------------------
template <int MODEL>
__device__ __forceinline__ void fwd_elem(const FwdStepArgs& a, int i, float u, float Iin,
                                         float v, float s, float x, float& v1, float& s1, float& x1, int b = 0,
                                         float w = 0.f, float* w1 = nullptr) {      // (w, w1): fourth state plane (ik_biexp_op: x)
    const float dt = a.dt;
    const float tau = is_ik(MODEL) ? 1.f : ldp(a.mp, RP_P_TAU, i, b), eta = ldp(a.mp, RP_P_ETA, i, b);
    if constexpr (!ModelTraits<MODEL>::SPIKING) {
        // li_op: v' = -v/tau + k*r_in + I_ext + eta          (leaky_integrator.yaml:10)
        v1 = v + dt * (-v / tau + u + Iin + eta);
        s1 = 0.f; x1 = 0.f;
    } else {
        const float tau_s = ldp(a.mp, RP_P_TAU_S, i, b);
        const bool p = v >= a.theta;                           // heaviside(v-theta, 1.0)   nodes.py:383,476
        const float pf = p ? 1.0f : 0.0f;
        float vt;
        if constexpr (is_ik(MODEL)) {
            // ik_op (ik.yaml:10-13): v' = (k (v-v_r)(v-v_theta) - u + I_ext + eta + g s_in (E_r - v)) / C
            //                        u' = (b (v-v_r) - u)/tau_u + kappa*spike ;  s' = -s/tau_s + spike      (x holds u, u holds g*W.s)
            // iku_op (ik.yaml:33-39): u' = (b (mean(v)-v_r) - u)/tau_u + kappa*mean(spike)   (population means per trial)
            // ik_biexp_op (ik.yaml:42-49): iku_op with  s' = -s/tau_d + x ;  x' = -x/tau_r + spike   (w holds x; tau_d, tau_r in the
            //                              tau_s, tau_x slots)
            const float C = ldp(a.mp, RP_P_C, i, b), kq = ldp(a.mp, RP_P_K, i, b), vr = ldp(a.mp, RP_P_VR, i, b), vth = ldp(a.mp, RP_P_VTH, i, b);
            const float Er = ldp(a.mp, RP_P_ER, i, b), bb = ldp(a.mp, RP_P_B, i, b), tau_u = ldp(a.mp, RP_P_TAU_U, i, b), kappa = ldp(a.mp, RP_P_KAPPA, i, b);
            vt = v + dt * ((kq * (v - vr) * (v - vth) - x + Iin + eta + u * (Er - v)) / C);
            if constexpr (is_mean_field(MODEL)) {
                const float2 mfb = a.mf[b];
                x1 = x + dt * ((bb * (mfb.x - vr) - x) / tau_u) + kappa * mfb.y;
            } else {
                x1 = x + dt * ((bb * (v - vr) - x) / tau_u) + kappa * pf;
            }
            if constexpr (MODEL == RP_IK_BIEXP) {
                s1 = s + dt * (-s / tau_s + w);
                if (w1) *w1 = w + dt * (-w / ldp(a.mp, RP_P_TAU_X, i, b)) + pf;
            } else {
                s1 = s + dt * (-s / tau_s) + pf;
            }
        } else if constexpr (MODEL == RP_LIF) {
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
                const float tau_x = ldp(a.mp, RP_P_TAU_X, i, b), alpha = ldp(a.mp, RP_P_ALPHA, i, b);
                x1 = x + dt * (-x / tau_x) + alpha * pf;      // x' = -x/tau_x + alpha*spike     (qif.yaml:31)
            } else {
                x1 = 0.f;
            }
        }
        v1 = p ? a.v_reset : vt;                               // reset blend                     nodes.py:390
    }
}

// Per-neuron constants of the forward step, hoisted out of the trial loop by the fused epilogue.  Reciprocals replace the
// per-element divisions (x * (1/tau) differs from x / tau by at most 1 ulp; exact for the template default tau = 1).
struct FwdRow { float inv_tau, eta, inv_tau_s, inv_tau_x, alpha, wi0, wi1; };

template <int MODEL>
__device__ __forceinline__ FwdRow fwd_row(const FwdStepArgs& a, int i) {
    FwdRow r{1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f};
    if (!is_ik(MODEL)) r.inv_tau = 1.0f / ldp(a.mp, RP_P_TAU, i);
    r.eta = ldp(a.mp, RP_P_ETA, i);
    if (ModelTraits<MODEL>::SPIKING) r.inv_tau_s = 1.0f / ldp(a.mp, RP_P_TAU_S, i);
    if (MODEL == RP_QIF_SFA) { r.inv_tau_x = 1.0f / ldp(a.mp, RP_P_TAU_X, i); r.alpha = ldp(a.mp, RP_P_ALPHA, i); }
    if (a.in_mode == RP_IN_PROJ) {
        r.wi0 = __ldg(a.W_in + (size_t)i * a.m);
        if (a.m > 1) r.wi1 = __ldg(a.W_in + (size_t)i * a.m + 1);
    }
    return r;
}

// x0, x1: the first two input channels of trial b (broadcast), further channels are fetched here
template <int MODEL>
__device__ __forceinline__ void fwd_elem_fast(const FwdStepArgs& a, const FwdRow& r, int i, int b, float u, float x0, float x1,
                                              float xdense, float v, float s, float x, float& v1, float& s1, float& x1o) {
    const float dt = a.dt;
    float Iin = xdense;
    if (a.in_mode == RP_IN_PROJ) {
        Iin = fmaf(r.wi1, x1, r.wi0 * x0);
        for (int j = 2; j < a.m; ++j) Iin = fmaf(__ldg(a.W_in + (size_t)i * a.m + j), __ldg(a.x_t + (size_t)b * a.m + j), Iin);
    }
    if constexpr (!ModelTraits<MODEL>::SPIKING) {
        v1 = v + dt * (-v * r.inv_tau + u + Iin + r.eta);
        s1 = 0.f; x1o = 0.f;
    } else {
        const bool p = v >= a.theta;
        const float pf = p ? 1.0f : 0.0f;
        float vt;
        if constexpr (MODEL == RP_LIF) {
            const float Iv = a.in_target == 0 ? Iin : 0.f, Is = a.in_target == 1 ? Iin : 0.f;
            vt = v + dt * (-v * r.inv_tau + u + Iv + r.eta);
            s1 = s + dt * (-s * r.inv_tau_s + Is) + pf;
            x1o = 0.f;
        } else {
            float xx = 0.f;
            if constexpr (MODEL == RP_QIF_SFA) xx = x;
            vt = v + dt * ((v * v + r.eta - xx + Iin) * r.inv_tau + u);
            s1 = s + dt * (-s * r.inv_tau_s) + pf;
            if constexpr (MODEL == RP_QIF_SFA) x1o = x + dt * (-x * r.inv_tau_x) + r.alpha * pf; else x1o = 0.f;
        }
        v1 = p ? a.v_reset : vt;
    }
}

// everything step t does for element (neuron i, trial b) once its recurrent drive u is known
template <int MODEL>
__device__ __forceinline__ void fwd_element(const FwdStepArgs& a, int i, int b, float u) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    const size_t plane = (size_t)a.B * a.N;
    const size_t idx = (size_t)b * a.N + i;
    const float v = a.y_cur[idx];
    const float s = NSV > 1 ? a.y_cur[plane + idx] : 0.f;
    const float x = NSV > 2 ? a.y_cur[2 * plane + idx] : 0.f;
    const float w = NSV > 3 ? a.y_cur[3 * plane + idx] : 0.f;
    const float Iin = input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i);
    float v1, s1, x1, w1 = 0.f;
    fwd_elem<MODEL>(a, i, u, Iin, v, s, x, v1, s1, x1, b, w, &w1);
    if (is_ik(MODEL) && a.urec_out) a.urec_out[idx] = u;
    a.y_next[idx] = v1;
    if (NSV > 1) a.y_next[plane + idx] = s1;
    if (NSV > 2) a.y_next[2 * plane + idx] = x1;
    if (NSV > 3) a.y_next[3 * plane + idx] = w1;
    float src1;
    if constexpr (ModelTraits<MODEL>::SPIKING) src1 = s1; else src1 = rate_act<MODEL>(a.mp, i, v1, b);
    if (a.src_next) a.src_next[idx] = src1;
    if (a.src_hi) {
        float hi, lo;
        split_tf32(src1, hi, lo);
        reinterpret_cast<float*>(a.src_hi)[(size_t)b * a.ld_src + i] = hi;
        reinterpret_cast<float*>(a.src_lo)[(size_t)b * a.ld_src + i] = lo;
    }
}

template <int MODEL>
__global__ void __launch_bounds__(256) k_fwd_step(FwdStepArgs a) {
    const size_t total = (size_t)a.B * a.N;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / a.N), i = (int)(idx - (size_t)b * a.N);
        fwd_element<MODEL>(a, i, b, a.u[(size_t)b * a.ldu + i]);
    }
}

// source operand of the very first step (and of re-started runs): src_0 from y_0
// f16: 0 -> tf32 split into fp32 words, 1 -> binary16 split with scale `sc`.  amax_out (optional): max |src_0|.
template <int MODEL>
__global__ void __launch_bounds__(256) k_init_src(int N, int B, const float* y, ModelParams mp,
                                                   float* src, int ld_plain, void* src_hi, void* src_lo, int ld_src,
                                                   int f16, ScaleRef sc, float* amax_out) {
    const size_t plane = (size_t)B * N;
    const float scale = f16 ? exp2i(scale_expo(sc)) : 1.f;
    float amax = 0.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < plane; idx += (size_t)gridDim.x * blockDim.x) {
        const int b = (int)(idx / N), i = (int)(idx - (size_t)b * N);
        float r;
        if constexpr (ModelTraits<MODEL>::SPIKING) r = y[plane + idx]; else r = rate_act<MODEL>(mp, i, y[idx], b);
        amax = fmaxf(amax, fabsf(r));
        if (src) src[(size_t)b * ld_plain + i] = r;
        if (src_hi) {
            if (f16) {
                store_split1_f16(src_hi, src_lo, (size_t)b * ld_src + i, r, scale);
            } else {
                float hi, lo;
                split_tf32(r, hi, lo);
                reinterpret_cast<float*>(src_hi)[(size_t)b * ld_src + i] = hi;
                reinterpret_cast<float*>(src_lo)[(size_t)b * ld_src + i] = lo;
            }
        }
    }
    if (amax_out) {
        amax = warp_max(amax);
        if ((threadIdx.x & 31) == 0) atomic_max_nonneg(amax_out, amax);
    }
}

// ---- absolute maxima that the binary16 operand scales are derived from --------------------------------------------------------
// dst = max(dst, max |src[r*ld + c] * rowscale[r*rs_stride]|) over rows x cols   (rowscale == nullptr: 1)
static __global__ void __launch_bounds__(256) k_amax_2d(int rows, int cols, const float* __restrict__ src, size_t ld,
                                                  const float* __restrict__ rowscale, int rs_stride, float* dst) {
    float amax = 0.f;
    const size_t total = (size_t)rows * cols;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t r = idx / cols, c = idx - r * cols;
        float v = src[r * ld + c];
        if (rowscale) v *= __ldg(rowscale + r * rs_stride);
        amax = fmaxf(amax, fabsf(v));
    }
    amax = warp_max(amax);
    if ((threadIdx.x & 31) == 0 && amax > 0.f) atomic_max_nonneg(dst, amax);
}

// ------------------------------------------------------------------------------------------------------
// iku_op mean field: one block per trial.  Deterministic (fixed-order block reduction), so trials stay bit-independent.
// ------------------------------------------------------------------------------------------------------
// mf[b] = { mean_i v[b][i], mean_i [v[b][i] >= theta] }
static __global__ void __launch_bounds__(256) k_trial_means(int N, const float* __restrict__ v, float theta, float2* mf) {
    __shared__ float red[2][8];
    const int b = blockIdx.x;
    float sv = 0.f, sp = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) { const float x = v[(size_t)b * N + i]; sv += x; sp += (x >= theta) ? 1.f : 0.f; }
    for (int o = 16; o > 0; o >>= 1) { sv += __shfl_xor_sync(0xffffffffu, sv, o); sp += __shfl_xor_sync(0xffffffffu, sp, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv; red[1][threadIdx.x >> 5] = sp; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
        mf[b] = make_float2(a / (float)N, c / (float)N);
    }
}
// asum[b] = { mean_i(ax[b][i] dt b_i / tau_u_i), mean_i(ax[b][i] kappa_i) }     (ax = adjoint of the recovery variable, plane 2)
static __global__ void __launch_bounds__(256) k_trial_adj_sums(int N, int B, const float* __restrict__ ax, ModelParams mp, float dt, float2* asum) {
    __shared__ float red[2][8];
    const int b = blockIdx.x;
    float s0 = 0.f, s1 = 0.f;
    for (int i = threadIdx.x; i < N; i += blockDim.x) {
        const float a = ax[(size_t)b * N + i];
        s0 += a * dt * ldp(mp, RP_P_B, i, b) / ldp(mp, RP_P_TAU_U, i, b);
        s1 += a * ldp(mp, RP_P_KAPPA, i, b);
    }
    for (int o = 16; o > 0; o >>= 1) { s0 += __shfl_xor_sync(0xffffffffu, s0, o); s1 += __shfl_xor_sync(0xffffffffu, s1, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s0; red[1][threadIdx.x >> 5] = s1; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, c = 0.f;
        for (int w = 0; w < 8; ++w) { a += red[0][w]; c += red[1][w]; }
        asum[b] = make_float2(a / (float)N, c / (float)N);
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
    int skip_out;           // 1: the output is produced elsewhere (fused readout rows); only record state variables
};

template <int MODEL>
__device__ __forceinline__ float out_value(const ObsArgs& a, const float* y, size_t plane, int b, int i) {
    const size_t idx = (size_t)b * a.N + i;
    if (a.out_var == RP_VAR_R) {
        if constexpr (!ModelTraits<MODEL>::SPIKING) return rate_act<MODEL>(a.mp, i, y[idx], b);
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
    const bool need_out = !a.skip_out && (a.out_rec_j != nullptr || !a.win_close);
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
constexpr int ADJ_TX = 32, ADJ_TY = 8, ADJ_BPT = 8;   // block covers 32 neurons x 64 trials
constexpr int ADJ_NACC = RP_NUM_PARAMS + RP_MAX_IN + RP_MAX_OUT;

// per-thread context of the adjoint: one neuron i, running sums of its parameter-gradient contributions
template <int MODEL>
struct AdjCtx {
    int i;
    float tau, tau_s, tau_x, alpha;
    float acc[ADJ_NACC];     // [0,RP_NUM_PARAMS): dparams ; then dW_in[i][0..m) ; then dW_out[0..k)[i]
};

template <int MODEL>
__device__ __forceinline__ void adj_ctx_init(const AdjArgs& a, AdjCtx<MODEL>& c, int i) {
    c.i = i;
    c.tau = 1.f; c.tau_s = 1.f; c.tau_x = 1.f; c.alpha = 0.f;
#pragma unroll
    for (int q = 0; q < ADJ_NACC; ++q) c.acc[q] = 0.f;
    if (i < a.N) {
        if (!is_ik(MODEL)) c.tau = ldp(a.mp, RP_P_TAU, i);
        if (ModelTraits<MODEL>::SPIKING) c.tau_s = ldp(a.mp, RP_P_TAU_S, i);
        if (MODEL == RP_QIF_SFA) { c.tau_x = ldp(a.mp, RP_P_TAU_X, i); c.alpha = ldp(a.mp, RP_P_ALPHA, i); }
    }
}

// Accumulation policies for the parameter-gradient contributions of one neuron
struct RegAcc {            // private running sums (stand-alone kernel: reduced over trial lanes afterwards)
    float* acc;
    __device__ __forceinline__ void add(int q, float v) const { acc[q] += v; }
};
struct SmemAcc {           // shared-memory atomics, one column of [ADJ_NACC][stride] per neuron (fused epilogue)
    float* base; int stride; unsigned mask;
    __device__ __forceinline__ void add(int q, float v) const { if (mask & (1u << q)) atomicAdd(base + q * stride, v); }
};

struct RowAcc {            // five per-neuron template-parameter sums in registers (fused epilogue); edge gradients handled elsewhere
    float* pacc;           // [5]: eta, tau, tau_s, tau_x, alpha
    __device__ __forceinline__ void add(int q, float v) const {
        if (q == RP_P_ETA) pacc[0] += v; else if (q == RP_P_TAU) pacc[1] += v; else if (q == RP_P_TAU_S) pacc[2] += v;
        else if (q == RP_P_TAU_X) pacc[3] += v; else if (q == RP_P_ALPHA) pacc[4] += v;
    }
};
struct NoAcc { __device__ __forceinline__ void add(int, float) const {} };

struct AdjRowParams { float tau, tau_s, tau_x, alpha; };

template <int MODEL>
__device__ __forceinline__ AdjRowParams adj_row_params(const AdjArgs& a, int i, int b = 0) {
    AdjRowParams r{1.f, 1.f, 1.f, 0.f};
    if (i < a.N) {
        if (!is_ik(MODEL)) r.tau = ldp(a.mp, RP_P_TAU, i, b);
        if (ModelTraits<MODEL>::SPIKING) r.tau_s = ldp(a.mp, RP_P_TAU_S, i, b);
        if (MODEL == RP_QIF_SFA) { r.tau_x = ldp(a.mp, RP_P_TAU_X, i, b); r.alpha = ldp(a.mp, RP_P_ALPHA, i, b); }
    }
    return r;
}

// "post" of reverse step t for one element, pure arithmetic: (av, as, ax) = adjoint at t+1 in, adjoint at t out.
// (v, s, x) = y_t.  Z = ((kW)^T g_t)[b][i].  Returns dI = dL/d(input current of step t).
template <int MODEL, class Acc>
__device__ __forceinline__ float adj_post_math(const AdjArgs& a, const AdjRowParams& rp_, const Acc& acc, int i, int b, float Z,
                                               float v, float s, float x, float& av, float& as, float& ax, float urec = 0.f,
                                               float w = 0.f, float* aw = nullptr) {     // (w, aw): fourth state plane and its adjoint
    constexpr bool SPK = ModelTraits<MODEL>::SPIKING;
    const float dt = a.dt, tau = rp_.tau, tau_s = rp_.tau_s, tau_x = rp_.tau_x, alpha = rp_.alpha;
    // readout / record gradient flowing into y_t[out]
    float ro = 0.f, yout = 0.f;
    if (a.e_t) {
        if (a.out_var == RP_VAR_V) yout = v; else if (a.out_var == RP_VAR_S) yout = s;
        else if (a.out_var == RP_VAR_X) yout = x;
        else { if constexpr (!SPK) yout = rate_act<MODEL>(a.mp, i, v, b); }
        if (a.out_mode == RP_OUT_DENSE) {
            ro = __ldg(a.e_t + (size_t)b * a.N + i) * a.e_scale;
        } else {
#pragma unroll
            for (int q = 0; q < RP_MAX_OUT; ++q) {
                if (q < a.k) {
                    const float e = __ldg(a.e_t + (size_t)b * a.k + q) * a.e_scale;
                    ro = fmaf(__ldg(a.W_out + (size_t)q * a.N + i), e, ro);
                    acc.add(RP_NUM_PARAMS + RP_MAX_IN + q, e * yout);
                }
            }
        }
    }
    float dI = 0.f;
    float nav, nas = 0.f, nax = 0.f;
    if constexpr (!SPK) {
        const float rg = rate_act_grad<MODEL>(a.mp, i, v, b);
        nav = av * (1.0f - dt / tau) + rg * Z;
        if (a.out_var == RP_VAR_V) nav += ro; else if (a.out_var == RP_VAR_R) nav += rg * ro;
        dI = dt * av;
        acc.add(RP_P_ETA, dt * av);
        acc.add(RP_P_TAU, dt * av * v / (tau * tau));
    } else {
        const bool p = v >= a.theta;
        const float gv = p ? 0.f : av;
        const float d = 1.0f + a.slope * fabsf(v - a.theta);
        const float sg = 1.0f / (d * d);                 // Spike.backward          nodes.py:478-481
        if constexpr (is_mean_field(MODEL)) {
            // ik_biexp_op: iku_op whose spike enters the rise variable x (plane 3: w, aw) instead of s:  s' = -s/tau_d + x,
            // x' = -x/tau_r + spike  ->  the surrogate term carries aw, and  aw <- aw (1 - dt/tau_r) + dt as
            // iku_op: as ik_op, but u' couples to the population means: d u_i'/d v_j = b_i / (N tau_u_i) and, through the surrogate,
            // d u_i'/d v_j = kappa_i sg_j / N for every j  ->  the local terms ax*dt*b/tau_u and kappa*ax become trial means (asum)
            const float C = ldp(a.mp, RP_P_C, i, b), kq = ldp(a.mp, RP_P_K, i, b), vr = ldp(a.mp, RP_P_VR, i, b), vth = ldp(a.mp, RP_P_VTH, i, b);
            const float Er = ldp(a.mp, RP_P_ER, i, b), bb = ldp(a.mp, RP_P_B, i, b), tau_u = ldp(a.mp, RP_P_TAU_U, i, b);
            const float eta = ldp(a.mp, RP_P_ETA, i, b);
            const float2 mfb = a.mf_t[b], sums = a.asum[b];
            float a_spk = as;                                   // adjoint of the variable that receives the spike
            if constexpr (MODEL == RP_IK_BIEXP) a_spk = aw ? *aw : 0.f;
            nav = gv * (1.0f + dt * (kq * (2.0f * v - vr - vth) - urec) / C) + sums.x + sg * (a_spk + sums.y);
            nas = as * (1.0f - dt / tau_s) + Z;
            if constexpr (MODEL == RP_IK_BIEXP) {
                const float tau_r = ldp(a.mp, RP_P_TAU_X, i, b);
                acc.add(RP_P_TAU_X, a_spk * w * dt / (tau_r * tau_r));
                if (aw) *aw = a_spk * (1.0f - dt / tau_r) + dt * as;
            }
            nax = ax * (1.0f - dt / tau_u) - gv * dt / C;
            dI = dt / C * gv;
            const float Iin = a.dparams[RP_P_C] ? input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i) : 0.f;
            const float q = kq * (v - vr) * (v - vth) - x + Iin + eta + urec * (Er - v);
            acc.add(RP_P_ETA, dI);
            acc.add(RP_P_C, -dt * gv * q / (C * C));
            acc.add(RP_P_K, dt * gv * (v - vr) * (v - vth) / C);
            acc.add(RP_P_VR, -dt * gv * kq * (v - vth) / C - ax * dt * bb / tau_u);
            acc.add(RP_P_VTH, -dt * gv * kq * (v - vr) / C);
            acc.add(RP_P_ER, dt * gv * urec / C);
            acc.add(RP_P_B, ax * dt * (mfb.x - vr) / tau_u);
            acc.add(RP_P_TAU_U, -ax * dt * (bb * (mfb.x - vr) - x) / (tau_u * tau_u));
            acc.add(RP_P_KAPPA, ax * mfb.y);
            acc.add(RP_P_TAU_S, as * s * dt / (tau_s * tau_s));
        } else if constexpr (MODEL == RP_IK) {
            // x == u (recovery variable), ax == its adjoint; urec = g*(W s_t) of the forward step
            const float C = ldp(a.mp, RP_P_C, i, b), kq = ldp(a.mp, RP_P_K, i, b), vr = ldp(a.mp, RP_P_VR, i, b), vth = ldp(a.mp, RP_P_VTH, i, b);
            const float Er = ldp(a.mp, RP_P_ER, i, b), bb = ldp(a.mp, RP_P_B, i, b), tau_u = ldp(a.mp, RP_P_TAU_U, i, b), kappa = ldp(a.mp, RP_P_KAPPA, i, b);
            const float eta = ldp(a.mp, RP_P_ETA, i, b);
            const float pf = p ? 1.0f : 0.0f;
            nav = gv * (1.0f + dt * (kq * (2.0f * v - vr - vth) - urec) / C) + ax * dt * bb / tau_u + sg * (as + kappa * ax);
            nas = as * (1.0f - dt / tau_s) + Z;
            nax = ax * (1.0f - dt / tau_u) - gv * dt / C;
            dI = dt / C * gv;
            const float Iin = a.dparams[RP_P_C] ? input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i) : 0.f;
            const float q = kq * (v - vr) * (v - vth) - x + Iin + eta + urec * (Er - v);
            acc.add(RP_P_ETA, dI);
            acc.add(RP_P_C, -dt * gv * q / (C * C));
            acc.add(RP_P_K, dt * gv * (v - vr) * (v - vth) / C);
            acc.add(RP_P_VR, -dt * gv * kq * (v - vth) / C - ax * dt * bb / tau_u);
            acc.add(RP_P_VTH, -dt * gv * kq * (v - vr) / C);
            acc.add(RP_P_ER, dt * gv * urec / C);
            acc.add(RP_P_B, ax * dt * (v - vr) / tau_u);
            acc.add(RP_P_TAU_U, -ax * dt * (bb * (v - vr) - x) / (tau_u * tau_u));
            acc.add(RP_P_KAPPA, ax * pf);
            acc.add(RP_P_TAU_S, as * s * dt / (tau_s * tau_s));
        } else if constexpr (MODEL == RP_LIF) {
            nav = gv * (1.0f - dt / tau) + sg * as;
            nas = as * (1.0f - dt / tau_s) + Z;
            dI = a.in_target == 0 ? dt * gv : dt * as;
            acc.add(RP_P_ETA, dt * gv);
            acc.add(RP_P_TAU, dt * gv * v / (tau * tau));
            acc.add(RP_P_TAU_S, as * s * dt / (tau_s * tau_s));
        } else {
            const float Iin = a.dparams[RP_P_TAU] ? input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i) : 0.f;
            const float eta = ldp(a.mp, RP_P_ETA, i, b);
            nav = gv * (1.0f + 2.0f * dt * v / tau) + sg * (as + alpha * ax);
            nas = as * (1.0f - dt / tau_s) + Z;
            dI = dt / tau * gv;
            acc.add(RP_P_ETA, dI);
            acc.add(RP_P_TAU, -dt * gv * (v * v + eta - x + Iin) / (tau * tau));
            acc.add(RP_P_TAU_S, as * s * dt / (tau_s * tau_s));
            if constexpr (MODEL == RP_QIF_SFA) {
                nax = ax * (1.0f - dt / tau_x) - dI;
                acc.add(RP_P_TAU_X, ax * x * dt / (tau_x * tau_x));
                acc.add(RP_P_ALPHA, ax * (p ? 1.0f : 0.0f));
            }
        }
        if (a.out_var == RP_VAR_V) nav += ro; else if (a.out_var == RP_VAR_S) nas += ro; else if (a.out_var == RP_VAR_X) nax += ro;
    }
    if (a.in_mode == RP_IN_PROJ && a.dW_in) {
#pragma unroll
        for (int j = 0; j < RP_MAX_IN; ++j)
            if (j < a.m) acc.add(RP_NUM_PARAMS + j, dI * __ldg(a.x_t + (size_t)b * a.m + j));
    }
    av = nav; as = nas; ax = nax;
    if (a.zero_after_post) { av = 0.f; as = 0.f; ax = 0.f; if (aw) *aw = 0.f; }
    return dI;
}

// "pre" of reverse step t-1: g_{t-1} = dt * gate_{t-1} * a_t  and the source value r_{t-1};  (vm, sm) = (v, s) of y_{t-1}
template <int MODEL>
__device__ __forceinline__ void adj_pre_math(const AdjArgs& a, int i, float av, float vm, float sm, float& g, float& srcv, int b = 0) {
    float gate = 1.0f;
    if constexpr (ModelTraits<MODEL>::SPIKING) { gate = (vm >= a.theta) ? 0.f : 1.0f; srcv = sm; }
    else srcv = rate_act<MODEL>(a.mp, i, vm, b);
    g = a.dt * gate * av;
    if constexpr (is_ik(MODEL)) g *= (ldp(a.mp, RP_P_ER, i, b) - vm) / ldp(a.mp, RP_P_C, i, b);    // d v' / d(g W s) = dt (E_r - v) / C
}

// scalar driver: loads, math, stores for element (neuron c.i, trial b)
template <int MODEL>
__device__ __forceinline__ void adj_element(const AdjArgs& a, AdjCtx<MODEL>& c, int b, float Z, float& g_out, float& src_out) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    const int i = c.i;
    const size_t plane = (size_t)a.B * a.N;
    const size_t idx = (size_t)b * a.N + i;
    float av = a.adj[idx];
    float as = NSV > 1 ? a.adj[plane + idx] : 0.f;
    float ax = NSV > 2 ? a.adj[2 * plane + idx] : 0.f;
    float aw = NSV > 3 ? a.adj[3 * plane + idx] : 0.f;
    if (a.do_post) {
        const float v = __ldg(a.y_t + idx);
        const float s = NSV > 1 ? __ldg(a.y_t + plane + idx) : 0.f;
        const float x = NSV > 2 ? __ldg(a.y_t + 2 * plane + idx) : 0.f;
        const float w = NSV > 3 ? __ldg(a.y_t + 3 * plane + idx) : 0.f;
        const AdjRowParams rp_ = a.per_trial ? adj_row_params<MODEL>(a, i, b) : AdjRowParams{c.tau, c.tau_s, c.tau_x, c.alpha};
        const RegAcc acc{c.acc};
        const float urec = (is_ik(MODEL) && a.urec_t) ? __ldg(a.urec_t + idx) : 0.f;
        const float dI = adj_post_math<MODEL>(a, rp_, acc, i, b, Z, v, s, x, av, as, ax, urec, w, &aw);
        if (a.g_x_t) a.g_x_t[idx] = dI;
        a.adj[idx] = av;
        if (NSV > 1) a.adj[plane + idx] = as;
        if (NSV > 2) a.adj[2 * plane + idx] = ax;
        if (NSV > 3) a.adj[3 * plane + idx] = aw;
    }
    g_out = 0.f; src_out = 0.f;
    if (a.do_pre) {
        const float vm = __ldg(a.y_tm1 + idx);
        const float sm = NSV > 1 ? __ldg(a.y_tm1 + plane + idx) : 0.f;
        float g, srcv;
        adj_pre_math<MODEL>(a, i, av, vm, sm, g, srcv, b);
        if (a.g) a.g[idx] = g;
        if (a.src) a.src[idx] = srcv;
        if (a.g_hi) {
            float hi, lo;
            split_tf32(g, hi, lo);
            a.g_hi[(size_t)b * a.ld_g + i] = hi;
            a.g_lo[(size_t)b * a.ld_g + i] = lo;
        }
        g_out = g; src_out = srcv;
    }
}

// destination of accumulator q (nullptr: not requested); uniform across the grid
__device__ __forceinline__ float* adj_acc_dst(const AdjArgs& a, int q, int i, size_t& off) {
    float* dst = nullptr;
    if (q < RP_NUM_PARAMS) { dst = a.dparams[q]; off = i; }
    else if (q < RP_NUM_PARAMS + RP_MAX_IN) {
        if (a.dW_in && (q - RP_NUM_PARAMS) < a.m) dst = a.dW_in;
        off = (size_t)i * a.m + (q - RP_NUM_PARAMS);
    } else {
        if (a.dW_out && (q - RP_NUM_PARAMS - RP_MAX_IN) < a.k && a.out_mode == RP_OUT_READOUT) dst = a.dW_out;
        off = (size_t)(q - RP_NUM_PARAMS - RP_MAX_IN) * a.N + i;
    }
    return dst;
}

template <int MODEL>
__global__ void __launch_bounds__(ADJ_TX * ADJ_TY) k_adj_step(AdjArgs a) {
    __shared__ float red[ADJ_TY][ADJ_TX + 1];
    // block tile: 32 neurons x 64 trials.  The weight-gradient operands are wanted trial-major ([neuron][trial]); they are
    // staged here and written with the trial index fastest so that both layouts are stored with full 128-byte rows.
    __shared__ float tg[ADJ_TY * ADJ_BPT][ADJ_TX + 1];
    __shared__ float ts[ADJ_TY * ADJ_BPT][ADJ_TX + 1];
    const int i = blockIdx.x * ADJ_TX + threadIdx.x;
    const int bblk = blockIdx.y * (ADJ_TY * ADJ_BPT);
    const bool valid_i = i < a.N;
    const bool transposed = a.do_pre && a.gT_hi != nullptr;
    AdjCtx<MODEL> c;
    adj_ctx_init<MODEL>(a, c, i);

    // Loads of a batch of 4 trials are issued before any of its stores (the adjoint is updated in place, so the compiler
    // cannot hoist them itself): ~30 independent 128-byte row requests in flight per warp.
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr int BATCH = 4;
    const size_t plane = (size_t)a.B * a.N;
    const AdjRowParams rp0{c.tau, c.tau_s, c.tau_x, c.alpha};
    const RegAcc racc{c.acc};
    float gmax = 0.f;
    for (int l0 = 0; l0 < ADJ_BPT; l0 += BATCH) {
        float av[BATCH], as[BATCH], ax[BATCH], v[BATCH], s[BATCH], x[BATCH], vm[BATCH], sm[BATCH], Z[BATCH], ur[BATCH];
        float aw[BATCH], w[BATCH];      // fourth state plane (ik_biexp_op)
        bool ok[BATCH];
#pragma unroll
        for (int l = 0; l < BATCH; ++l) {
            const int b = bblk + (l0 + l) * ADJ_TY + threadIdx.y;
            ok[l] = valid_i && b < a.B;
            const size_t idx = (size_t)b * a.N + i;
            av[l] = as[l] = ax[l] = v[l] = s[l] = x[l] = vm[l] = sm[l] = Z[l] = ur[l] = aw[l] = w[l] = 0.f;
            if (ok[l]) {
                av[l] = a.adj[idx];
                if (NSV > 1) as[l] = a.adj[plane + idx];
                if (NSV > 2) ax[l] = a.adj[2 * plane + idx];
                if (NSV > 3) aw[l] = a.adj[3 * plane + idx];
                if (a.do_post) {
                    v[l] = __ldg(a.y_t + idx);
                    if (NSV > 1) s[l] = __ldg(a.y_t + plane + idx);
                    if (NSV > 2) x[l] = __ldg(a.y_t + 2 * plane + idx);
                    if (NSV > 3) w[l] = __ldg(a.y_t + 3 * plane + idx);
                    Z[l] = a.Z[(size_t)b * a.ldz + i];
                    if (is_ik(MODEL) && a.urec_t) ur[l] = __ldg(a.urec_t + idx);
                }
                if (a.do_pre) {
                    vm[l] = __ldg(a.y_tm1 + idx);
                    if (NSV > 1) sm[l] = __ldg(a.y_tm1 + plane + idx);
                }
            }
        }
#pragma unroll
        for (int l = 0; l < BATCH; ++l) {
            const int bl = (l0 + l) * ADJ_TY + threadIdx.y;
            const int b = bblk + bl;
            const size_t idx = (size_t)b * a.N + i;
            float g = 0.f, srcv = 0.f;
            if (ok[l]) {
                if (a.do_post) {
                    const AdjRowParams rp_ = a.per_trial ? adj_row_params<MODEL>(a, i, b) : rp0;
                    const float dI = adj_post_math<MODEL>(a, rp_, racc, i, b, Z[l], v[l], s[l], x[l], av[l], as[l], ax[l], ur[l], w[l], &aw[l]);
                    if (a.g_x_t) a.g_x_t[idx] = dI;
                    a.adj[idx] = av[l];
                    if (NSV > 1) a.adj[plane + idx] = as[l];
                    if (NSV > 2) a.adj[2 * plane + idx] = ax[l];
                    if (NSV > 3) a.adj[3 * plane + idx] = aw[l];
                }
                if (a.do_pre) {
                    adj_pre_math<MODEL>(a, i, av[l], vm[l], sm[l], g, srcv, b);
                    gmax = fmaxf(gmax, fabsf(g));
                    if (a.g) a.g[idx] = g;
                    if (a.src) a.src[idx] = srcv;
                    if (a.g_hi) {
                        float hi, lo;
                        split_tf32(g, hi, lo);
                        a.g_hi[(size_t)b * a.ld_g + i] = hi;
                        a.g_lo[(size_t)b * a.ld_g + i] = lo;
                    }
                }
            }
            if (transposed) { tg[bl][threadIdx.x] = g; ts[bl][threadIdx.x] = srcv; }
        }
    }
    if (a.g_amax && a.do_pre) {
        gmax = warp_max(gmax);
        if (threadIdx.x == 0 && gmax > 0.f) atomic_max_nonneg(a.g_amax, gmax);
    }
    if (transposed) {
        __syncthreads();
        // thread (tx, ty) -> trial bblk + tx (+32), neurons ty, ty+8, ... : 32 consecutive trials per warp store
        for (int h = 0; h < (ADJ_TY * ADJ_BPT) / 32; ++h) {
            const int bl = h * 32 + threadIdx.x;
            const int b = bblk + bl;
            for (int r = threadIdx.y; r < ADJ_TX; r += ADJ_TY) {
                const int ii = blockIdx.x * ADJ_TX + r;
                if (ii < a.N && b < a.B) {
                    float hi, lo;
                    const size_t off = (size_t)ii * a.ld_t + a.t_col0 + b;
                    split_tf32(tg[bl][r], hi, lo);
                    a.gT_hi[off] = hi; a.gT_lo[off] = lo;
                    split_tf32(ts[bl][r], hi, lo);
                    a.srcT_hi[off] = hi; a.srcT_lo[off] = lo;
                }
            }
        }
    }

    if (a.do_post && a.any_param_grad) {
        // reduce the per-thread partial sums over the 8 trial lanes, then one atomic per (neuron, quantity)
#pragma unroll
        for (int q = 0; q < ADJ_NACC; ++q) {
            size_t off = 0;
            float* dst = adj_acc_dst(a, q, i, off);
            if (dst == nullptr) continue;     // uniform across the block
            __syncthreads();
            red[threadIdx.y][threadIdx.x] = c.acc[q];
            __syncthreads();
            if (threadIdx.y == 0 && valid_i) {
                float t = 0.f;
#pragma unroll
                for (int r = 0; r < ADJ_TY; ++r) t += red[r][threadIdx.x];
                atomicAdd(dst + off, t);
            }
        }
    }
}

// Vectorised variant for the tensor-core path when no per-neuron parameter sums are requested (edge gradients come from the
// contractions / k_readout_grad): a thread owns 4 consecutive neurons -> every global access is a 16-byte vector, loads of two
// trials are in flight before the first store, and the trial-major weight-gradient operands are transposed through shared memory.
// Block = 32 x 8 threads, tile = 128 neurons x 32 trials.  Requires N % 128 == 0, B % 32 == 0, 16-byte aligned rows.
constexpr int ADJ4_TB = 32;      // trials per block
__device__ __forceinline__ float f4at(const float4& v, int r) { return r == 0 ? v.x : (r == 1 ? v.y : (r == 2 ? v.z : v.w)); }

// TR: stage the trial-major weight-gradient operands (tf32 path); the binary16 path converts in k_adj_convert_f16 and needs no smem.
// NTY: warps per block (block = 32 x NTY threads, tile = 128 neurons x 4*NTY trials).  The binary16 path runs 4-warp blocks of 80
// registers so that one block fits beside a resident weight-gradient GEMM CTA (168 regs x 320 threads) on the same SM.
template <int MODEL, bool TR, int NTY>
__global__ void __launch_bounds__(32 * NTY, NTY == 4 ? 6 : 3) k_adj_step_v4(AdjArgs a) {      // <= 80 registers: 24 warps per SM (113 registers left it at 16)
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr int TB = 4 * NTY;
    static_assert(!TR || TB == ADJ4_TB, "the transposing variant stages 32 trials");
    // (no shared memory at all when !TR: a resident GEMM CTA leaves under 2 KB of the SM's carveout unused)
    __shared__ float tg_[TR ? ADJ4_TB * (128 + 4) : 0 + !TR];
    __shared__ float ts_[TR ? ADJ4_TB * (128 + 4) : 0 + !TR];
    float (*tg)[128 + 4] = reinterpret_cast<float (*)[128 + 4]>(tg_);
    float (*ts)[128 + 4] = reinterpret_cast<float (*)[128 + 4]>(ts_);
    const int tx = threadIdx.x, ty = threadIdx.y;
    const int i0 = blockIdx.x * 128 + 4 * tx;
    const int bblk = blockIdx.y * TB;
    const size_t plane = (size_t)a.B * a.N;
    const bool transposed = TR && a.do_pre && a.gT_hi != nullptr;
    pdl_launch_dependents();
    pdl_wait();                      // Z of this step, the adjoint state and the previous conversion come from earlier kernels
    TraceRec* trace = (tx == 0 && ty == 0) ? trace_begin(TR_ADJ_STEP) : nullptr;
    const NoAcc nacc;
    AdjRowParams rp_[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) rp_[rr] = adj_row_params<MODEL>(a, i0 + rr);
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float gmax = 0.f;

    for (int l0 = 0; l0 < 4; l0 += 2) {
        float4 av[2], as[2], ax[2], v[2], s[2], x[2], vm[2], sm[2], Z[2], ur[2];
        float4 aw[2], w[2];             // fourth state plane (ik_biexp_op)
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int b = bblk + (l0 + l) * NTY + ty;
            const size_t idx = (size_t)b * a.N + i0;
            av[l] = *reinterpret_cast<const float4*>(a.adj + idx);
            as[l] = NSV > 1 ? *reinterpret_cast<const float4*>(a.adj + plane + idx) : zero;
            ax[l] = NSV > 2 ? *reinterpret_cast<const float4*>(a.adj + 2 * plane + idx) : zero;
            aw[l] = NSV > 3 ? *reinterpret_cast<const float4*>(a.adj + 3 * plane + idx) : zero;
            v[l] = s[l] = x[l] = vm[l] = sm[l] = Z[l] = ur[l] = w[l] = zero;
            if (a.do_post) {
                v[l] = __ldg(reinterpret_cast<const float4*>(a.y_t + idx));
                if (NSV > 1) s[l] = __ldg(reinterpret_cast<const float4*>(a.y_t + plane + idx));
                if (NSV > 2) x[l] = __ldg(reinterpret_cast<const float4*>(a.y_t + 2 * plane + idx));
                if (NSV > 3) w[l] = __ldg(reinterpret_cast<const float4*>(a.y_t + 3 * plane + idx));
                Z[l] = *reinterpret_cast<const float4*>(a.Z + (size_t)b * a.ldz + i0);
                if (is_ik(MODEL) && a.urec_t) ur[l] = __ldg(reinterpret_cast<const float4*>(a.urec_t + idx));
            }
            if (a.do_pre) {
                vm[l] = __ldg(reinterpret_cast<const float4*>(a.y_tm1 + idx));
                // s_{t-1} is only the weight-gradient source value: on the binary16 path (!TR) the convert kernel reads it itself
                if (NSV > 1 && (TR || a.src)) sm[l] = __ldg(reinterpret_cast<const float4*>(a.y_tm1 + plane + idx));
            }
        }
#pragma unroll
        for (int l = 0; l < 2; ++l) {
            const int bl = (l0 + l) * NTY + ty;
            const int b = bblk + bl;
            const size_t idx = (size_t)b * a.N + i0;
            float nav[4], nas[4], nax[4], naw[4], g[4], sv[4], gh[4], gl[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                nav[rr] = f4at(av[l], rr); nas[rr] = f4at(as[l], rr); nax[rr] = f4at(ax[l], rr); naw[rr] = f4at(aw[l], rr);
                if (a.do_post)
                    adj_post_math<MODEL>(a, rp_[rr], nacc, i0 + rr, b, f4at(Z[l], rr), f4at(v[l], rr), f4at(s[l], rr), f4at(x[l], rr),
                                         nav[rr], nas[rr], nax[rr], f4at(ur[l], rr), f4at(w[l], rr), &naw[rr]);
                g[rr] = 0.f; sv[rr] = 0.f;
                if (a.do_pre) adj_pre_math<MODEL>(a, i0 + rr, nav[rr], f4at(vm[l], rr), f4at(sm[l], rr), g[rr], sv[rr], b);
                gmax = fmaxf(gmax, fabsf(g[rr]));
                split_tf32(g[rr], gh[rr], gl[rr]);
            }
            if (a.do_post) {
                *reinterpret_cast<float4*>(a.adj + idx) = make_float4(nav[0], nav[1], nav[2], nav[3]);
                if (NSV > 1) *reinterpret_cast<float4*>(a.adj + plane + idx) = make_float4(nas[0], nas[1], nas[2], nas[3]);
                if (NSV > 2) *reinterpret_cast<float4*>(a.adj + 2 * plane + idx) = make_float4(nax[0], nax[1], nax[2], nax[3]);
                if (NSV > 3) *reinterpret_cast<float4*>(a.adj + 3 * plane + idx) = make_float4(naw[0], naw[1], naw[2], naw[3]);
            }
            if (a.do_pre) {
                if (a.g) *reinterpret_cast<float4*>(a.g + idx) = make_float4(g[0], g[1], g[2], g[3]);
                if (a.src) *reinterpret_cast<float4*>(a.src + idx) = make_float4(sv[0], sv[1], sv[2], sv[3]);
                if (a.g_hi) {
                    *reinterpret_cast<float4*>(a.g_hi + (size_t)b * a.ld_g + i0) = make_float4(gh[0], gh[1], gh[2], gh[3]);
                    *reinterpret_cast<float4*>(a.g_lo + (size_t)b * a.ld_g + i0) = make_float4(gl[0], gl[1], gl[2], gl[3]);
                }
                if constexpr (TR) {
                    if (transposed) {
                        *reinterpret_cast<float4*>(&tg[bl][4 * tx]) = make_float4(g[0], g[1], g[2], g[3]);
                        *reinterpret_cast<float4*>(&ts[bl][4 * tx]) = make_float4(sv[0], sv[1], sv[2], sv[3]);
                    }
                }
            }
        }
    }
    if (a.g_amax && a.do_pre) {
        gmax = warp_max(gmax);
        if (tx == 0 && gmax > 0.f) atomic_max_nonneg(a.g_amax, gmax);
    }
    trace_end(trace);
    if constexpr (TR) if (transposed) {
        __syncthreads();
        // lane tx -> trial bblk + tx; warp ty walks neurons ty, ty+8, ...: every store is 32 consecutive trials of one neuron
        for (int r = ty; r < 128; r += 8) {
            float hi, lo;
            const size_t off = (size_t)(blockIdx.x * 128 + r) * a.ld_t + a.t_col0 + bblk + tx;
            split_tf32(tg[tx][r], hi, lo);
            a.gT_hi[off] = hi; a.gT_lo[off] = lo;
            split_tf32(ts[tx][r], hi, lo);
            a.srcT_hi[off] = hi; a.srcT_lo[off] = lo;
        }
    }
}

// Rolling-pipeline variant for the binary16 path (no transposed staging, no parameter sums): a warp owns one 128-neuron tile
// (lane -> 4 neurons, per-neuron constants loaded once) and walks over trials with a stride; the loads of the NEXT trial are
// issued before the current trial is computed and stored, so every warp keeps one trial's worth of 16-byte loads in flight all
// the time instead of exposing the load latency once per batch.  Grid = 3 blocks of 8 warps per SM (<= 80 registers).
template <int MODEL>
struct AdjLoads { float4 av, as, ax, v, vm, Z, ur; };

template <int MODEL>
__device__ __forceinline__ void adj_v5_load(const AdjArgs& a, AdjLoads<MODEL>& L, int b, int i0, size_t plane) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    const size_t idx = (size_t)b * a.N + i0;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    L.av = *reinterpret_cast<const float4*>(a.adj + idx);
    L.as = NSV > 1 ? *reinterpret_cast<const float4*>(a.adj + plane + idx) : zero;
    L.ax = NSV > 2 ? *reinterpret_cast<const float4*>(a.adj + 2 * plane + idx) : zero;
    L.v = L.vm = L.Z = L.ur = zero;
    if (a.do_post) {
        L.v = __ldg(reinterpret_cast<const float4*>(a.y_t + idx));
        L.Z = *reinterpret_cast<const float4*>(a.Z + (size_t)b * a.ldz + i0);
        if (is_ik(MODEL) && a.urec_t) L.ur = __ldg(reinterpret_cast<const float4*>(a.urec_t + idx));
    }
    if (a.do_pre) L.vm = __ldg(reinterpret_cast<const float4*>(a.y_tm1 + idx));
}

template <int MODEL>
__global__ void __launch_bounds__(256, 3) k_adj_step_v5(AdjArgs a) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    static_assert(!is_mean_field(MODEL), "iku_op / ik_biexp_op use the generic adjoint kernels (trial means)");
    const int lane = threadIdx.x & 31;
    const int gw = blockIdx.x * 8 + (threadIdx.x >> 5), nw = gridDim.x * 8;
    const int NT = a.N / 128;
    const int stride = nw / NT;                     // warps per neuron tile = trial stride (host guarantees nw >= NT)
    const int tile = gw % NT, b0 = gw / NT;
    const size_t plane = (size_t)a.B * a.N;
    pdl_launch_dependents();
    pdl_wait();
    TraceRec* trace = threadIdx.x == 0 ? trace_begin(TR_ADJ_STEP) : nullptr;
    float gmax = 0.f;
    if (b0 < stride && b0 < a.B) {
        const int i0 = tile * 128 + 4 * lane;
        const NoAcc nacc;
        AdjRowParams rp_[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) rp_[rr] = adj_row_params<MODEL>(a, i0 + rr);
        AdjLoads<MODEL> cur, nxt;
        adj_v5_load<MODEL>(a, cur, b0, i0, plane);
        for (int b = b0; b < a.B; b += stride) {
            const bool more = b + stride < a.B;
            if (more) adj_v5_load<MODEL>(a, nxt, b + stride, i0, plane);
            const size_t idx = (size_t)b * a.N + i0;
            float nav[4], nas[4], nax[4], g[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                nav[rr] = f4at(cur.av, rr); nas[rr] = f4at(cur.as, rr); nax[rr] = f4at(cur.ax, rr);
                if (a.do_post)
                    adj_post_math<MODEL>(a, rp_[rr], nacc, i0 + rr, b, f4at(cur.Z, rr), f4at(cur.v, rr), 0.f, 0.f,
                                         nav[rr], nas[rr], nax[rr], f4at(cur.ur, rr));
                g[rr] = 0.f;
                float sv_unused = 0.f;
                if (a.do_pre) adj_pre_math<MODEL>(a, i0 + rr, nav[rr], f4at(cur.vm, rr), 0.f, g[rr], sv_unused, b);
                gmax = fmaxf(gmax, fabsf(g[rr]));
            }
            if (a.do_post) {
                *reinterpret_cast<float4*>(a.adj + idx) = make_float4(nav[0], nav[1], nav[2], nav[3]);
                if (NSV > 1) *reinterpret_cast<float4*>(a.adj + plane + idx) = make_float4(nas[0], nas[1], nas[2], nas[3]);
                if (NSV > 2) *reinterpret_cast<float4*>(a.adj + 2 * plane + idx) = make_float4(nax[0], nax[1], nax[2], nax[3]);
            }
            if (a.do_pre && a.g) *reinterpret_cast<float4*>(a.g + idx) = make_float4(g[0], g[1], g[2], g[3]);
            if (more) cur = nxt;
        }
    }
    if (a.g_amax && a.do_pre) {
        gmax = warp_max(gmax);
        if (lane == 0 && gmax > 0.f) atomic_max_nonneg(a.g_amax, gmax);
    }
    trace_end(trace);
}

// ------------------------------------------------------------------------------------------------------
// binary16 operands of the reverse sweep.  The adjoint kernel leaves g_{t-1} in fp32 together with its exact maximum; this
// kernel turns it (and the source values r_{t-1}) into the split binary16 operands of the next two contractions:
//   K-major    g_hi/lo  [B][ld_g]           scale 2^e(max|g_{t-1}|)      -> Z_{t-1} = (kW)^T g_{t-1}
//   trial-major gT, srcT [N][ld_t]          one scale per weight-gradient K chunk (several steps share one accumulator)
// The chunk scale is fixed by the first step of the chunk with a non-zero gradient, with 2^11 headroom for growth inside the
// chunk; growth beyond that is absorbed by moving up to 2^3 into the source operand, and anything beyond raises a flag that
// rp_plan_status reports (the caller should then use RP_PREC_3XTF32).
// ------------------------------------------------------------------------------------------------------
constexpr int CV_HG = 14;        // exact-maximum operands: max -> [2^13, 2^14)
constexpr int CV_HCHUNK = 4;     // weight-gradient chunk (up to 16 steps): first maximum -> [2^3, 2^4), i.e. 2^11 of growth before the 2^15 limit
constexpr int CV_HSRC = 12;      // source operand bound -> [2^11, 2^12)
struct ConvArgs {
    int N, B;
    const float* g32;            // [B][N]
    const float* src32;          // [B][N] source values of the same step (spiking: checkpoint plane s_{t-1}; rate: act(v_{t-1}))
    void *g_hi, *g_lo; int ld_g;
    void *gT_hi, *gT_lo, *srcT_hi, *srcT_lo; int ld_t, t_col0;     // gT_hi == nullptr: no weight gradient wanted
    const float* g_amax;         // exact max |g32| (complete: written by the previous kernel)
    float* g_amax_clear;         // the other parity slot, zeroed here for the next step
    const float* chunk_ref_in;   // reference maximum of the open chunk (previous step's chunk_ref_out)
    float* chunk_ref_out;
    int chunk_first;             // 1: this step opens a new chunk
    ScaleRef sc_src;             // bound of |src| over the whole sweep
    int* flags;                  // bit 0: binary16 range exceeded
};

constexpr int CV_TN = 64, CV_TB = 32;     // block tile: 64 neurons x 32 trials, 17 KB of shared memory (fits beside a GEMM CTA)
static __global__ void __launch_bounds__(256) k_adj_convert_f16(ConvArgs a) {
    __shared__ float tg[CV_TB][CV_TN + 4];
    __shared__ float ts[CV_TB][CV_TN + 4];
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;        // tx -> 4 neurons, ty -> trials ty, ty + 16
    const int i0 = blockIdx.x * CV_TN + 4 * tx;
    const int bblk = blockIdx.y * CV_TB;
    const bool first = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
    pdl_launch_dependents();
    pdl_wait();                      // g32 and its maximum come from the adjoint kernel just before
    TraceRec* trace = threadIdx.x == 0 ? trace_begin(TR_ADJ_CONVERT) : nullptr;
    const float gmax = *a.g_amax;
    const int eg = expo_for(gmax, CV_HG);
    const float sg = exp2i(eg);
    const bool wg = a.gT_hi != nullptr;
    float sgT = 1.f, ssT = 1.f;
    if (wg) {
        const float ref_in = *a.chunk_ref_in;
        const float ref = (a.chunk_first || !(ref_in > 0.f)) ? gmax : ref_in;
        const int ec = expo_for(ref, CV_HCHUNK);
        const int elim = expo_for(gmax, 15);                 // largest exponent that keeps this step's g below 2^15 (binary16 max: 65504)
        const int eu = gmax > 0.f ? min(ec, elim) : ec;     // an all-zero step (e.g. right after a truncation cut) constrains nothing
        const int comp = ec - eu;                            // moved into the source operand so that the product scale stays 2^(ec+es)
        const int es = scale_expo(a.sc_src);
        sgT = exp2i(eu);
        ssT = exp2i(es + min(comp, 3));
        if (first) {
            *a.chunk_ref_out = ref;
            if (comp > 3) atomicOr(a.flags, 1);
        }
    }
    if (first) *a.g_amax_clear = 0.f;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int bl = l * 16 + ty;
        const int b = bblk + bl;
        const size_t idx = (size_t)b * a.N + i0;
        const float4 g4 = *reinterpret_cast<const float4*>(a.g32 + idx);
        const float g[4] = {g4.x, g4.y, g4.z, g4.w};
        store_split4_f16(a.g_hi, a.g_lo, (size_t)b * a.ld_g + i0, g, sg);
        if (wg) {
            const float4 s4 = __ldg(reinterpret_cast<const float4*>(a.src32 + idx));
            *reinterpret_cast<float4*>(&tg[bl][4 * tx]) = g4;
            *reinterpret_cast<float4*>(&ts[bl][4 * tx]) = s4;
        }
    }
    if (wg) {
        __syncthreads();
        // lane -> trial bblk + lane; warp w walks neurons w, w+8, ...: every store is 32 consecutive trials of one neuron
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        for (int r = w; r < CV_TN; r += 8) {
            const size_t off = (size_t)(blockIdx.x * CV_TN + r) * a.ld_t + a.t_col0 + bblk + lane;
            store_split1_f16(a.gT_hi, a.gT_lo, off, tg[lane][r], sgT);
            store_split1_f16(a.srcT_hi, a.srcT_lo, off, ts[lane][r], ssT);
        }
    }
    trace_end(trace);
}

// ------------------------------------------------------------------------------------------------------
// Fused reverse element-wise step of the binary16 path (round 2): ONE kernel per reverse step does what k_adj_step_v5 +
// k_adj_convert_f16 (+ k_readout_grad after the sweep) did in two kernels and an fp32 round trip of g:
//   post  adjoint of step t from Z_t = (kW)^T g_t                                   
//   pre   g_{t-1} = dt * gate_{t-1} * a_t, written straight as split binary16 operands: K-major (next adjoint product) and
//         trial-major together with s_{t-1} (weight-gradient chunk; transposed through shared memory)
//   dW_out contribution of step t-1 (readout from s): per-trial-block partial sums, plain read-modify-write -> deterministic
// The K-major scale no longer needs the exact max |g_{t-1}| (which is only known when the kernel has finished): for the
// spiking templates a_{t-1}[v] is a function of (a_t, v_{t-1}) alone -- it does not involve Z_{t-1} -- so THIS kernel
// evaluates |a_{t-1}[v]| for every element it holds and leaves  nb = dt * max|a_{t-1}[v]| >= max|g_{t-2}|  for the next
// kernel (three rotating device scalars: read / accumulate / clear).  The bound is tight up to the spike gate of step t-2,
// it is mapped to [2^12, 2^13) (CV_HGB), and a range guard flags any element that would leave the binary16 range anyway.
// ------------------------------------------------------------------------------------------------------
constexpr int CV_HGB = 13;
struct FusedAdjArgs {
    void *g_hi, *g_lo; int ld_g;
    void *gT_hi, *gT_lo, *srcT_hi, *srcT_lo; int ld_t, t_col0;     // gT_hi == nullptr: no weight gradient wanted
    const float* nb_in;          // bound of max |g_{t-1}| (complete: written by the previous kernel of the sweep)
    float* nb_out;               // receives the bound of max |g_{t-2}| (atomicMax, zero on entry)
    float* nb_clear;             // the slot the NEXT kernel accumulates into, zeroed here
    const float* chunk_ref_in;   // reference maximum of the open weight-gradient chunk (as in ConvArgs)
    float* chunk_ref_out;
    int chunk_first;
    ScaleRef sc_src;
    int* flags;                  // bit 0: weight-gradient chunk range exceeded; bit 1: K-major operand range exceeded
    const float* e_tm1;          // dL/d out_rec of the record window that contains step t-1 ([B][k] | [B][N]) or nullptr
    float e_scale_tm1;
    float* dwout_part;           // [B / FA_TB][k][N] partial sums of dW_out (readout from s), or nullptr
};

#ifndef RP_FUSED_OCC
#define RP_FUSED_OCC 3
#endif
constexpr int FA_TN = 64, FA_TB = 32, FA_LD = FA_TB + 1;
__device__ __forceinline__ float rcp_approx(float x) { float r; asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(x)); return r; }
// per-neuron constants of the reverse step, reciprocals hoisted (the generic adj_post_math divides per element)
struct FaRow { float fa, ds, dx, alpha, c1; };
template <int MODEL>
__device__ __forceinline__ FaRow fa_row(const AdjArgs& a, int i) {
    FaRow r;
    const float dt_tau = a.dt * rcp_approx(ldp(a.mp, RP_P_TAU, i));         // (an IEEE division costs ~15 instructions and a slow path)
    r.fa = MODEL == RP_LIF ? 1.0f - dt_tau : 2.0f * dt_tau;                 // d v_{t+1} / d v_t = fa (lif) | 1 + fa v_t (qif)
    r.ds = 1.0f - a.dt * rcp_approx(ldp(a.mp, RP_P_TAU_S, i));
    r.dx = 1.f; r.alpha = 0.f; r.c1 = dt_tau;
    if (MODEL == RP_QIF_SFA) { r.dx = 1.0f - a.dt * rcp_approx(ldp(a.mp, RP_P_TAU_X, i)); r.alpha = ldp(a.mp, RP_P_ALPHA, i); }
    return r;
}
// FULL: both halves of the step run (every launch of a sweep but its first and last) -- the flags become compile-time constants
template <int MODEL, bool FULL>
__global__ void __launch_bounds__(256, RP_FUSED_OCC) k_adj_fused_f16(AdjArgs a, FusedAdjArgs f) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr bool SFA = MODEL == RP_QIF_SFA;
    static_assert(MODEL == RP_QIF || MODEL == RP_QIF_SFA || MODEL == RP_LIF, "templates whose a_{t-1}[v] does not involve Z_{t-1}");
    __shared__ float tgT[FA_TN][FA_LD];        // g_{t-1}  [neuron][trial]  (row stride 33: conflict-free both ways)
    __shared__ float tsT[FA_TN][FA_LD];        // s_{t-1}
    __shared__ __align__(16) float se0[FA_TB][RP_MAX_OUT];    // e_t     of the tile's trials, divided by the window length
    __shared__ __align__(16) float se1[FA_TB][RP_MAX_OUT];    // e_{t-1}
    __shared__ __align__(16) float swo[RP_MAX_OUT][FA_TN];    // W_out columns of the tile's neurons
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;        // tx -> 4 neurons, ty -> trials ty, ty + 16
    const int i0 = blockIdx.x * FA_TN + 4 * tx;
    const int bblk = blockIdx.y * FA_TB;
    const bool first = blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0;
    const int N = a.N;
    const int plane = a.B * N;                  // (host guarantees batch * n < 2^31)
    const bool do_post = FULL || a.do_post, do_pre = FULL || a.do_pre;
    pdl_launch_dependents();
    pdl_wait();
    TraceRec* trace = threadIdx.x == 0 ? trace_begin(TR_ADJ_STEP) : nullptr;
    const bool wg = f.gT_hi != nullptr && do_pre;
    const bool have_e0 = a.e_t != nullptr && do_post, have_e1 = f.e_tm1 != nullptr && do_pre;
    const bool ro_acc = have_e1 && f.dwout_part != nullptr;
    __shared__ float s_scale[3];               // operand scales of this step: one thread derives them, everyone reads after the barrier
    __shared__ __align__(16) float s_row[FA_TN][SFA ? 8 : 2];      // per-neuron constants of the tile
    if (threadIdx.x == 0) {
        const float gb = *f.nb_in;
        float sgT = 1.f, ssT = 1.f;
        if (wg) {
            const float ref_in = *f.chunk_ref_in;
            const float ref = (f.chunk_first || !(ref_in > 0.f)) ? gb : ref_in;
            const int ec = expo_for(ref, CV_HCHUNK);
            const int elim = expo_for(gb, 15);
            const int eu = gb > 0.f ? min(ec, elim) : ec;
            const int comp = ec - eu;
            const int es = scale_expo(f.sc_src);
            sgT = exp2i(eu);
            ssT = exp2i(es + min(comp, 3));
            if (first) {
                *f.chunk_ref_out = ref;
                if (comp > 3) atomicOr(f.flags, 1);
            }
        }
        s_scale[0] = exp2i(expo_for(gb, CV_HGB)); s_scale[1] = sgT; s_scale[2] = ssT;
        if (first) *f.nb_clear = 0.f;
    }
    if (threadIdx.x >= 64 && threadIdx.x < 64 + FA_TN) {        // warps 2-3: the tile's per-neuron constants
        const int ii = threadIdx.x - 64;
        const FaRow r = fa_row<MODEL>(a, blockIdx.x * FA_TN + ii);
        s_row[ii][0] = r.fa; s_row[ii][1] = r.ds;
        if (SFA) { s_row[ii][2] = r.dx; s_row[ii][3] = r.alpha; s_row[ii][4] = r.c1; }
    }
    {   // readout-gradient operands of the tile (READOUT mode only; the host keeps dense-output problems off this kernel)
        const int bl = threadIdx.x / RP_MAX_OUT, q = threadIdx.x % RP_MAX_OUT;        // 32 x 8 == blockDim
        se0[bl][q] = (have_e0 && q < a.k) ? __ldg(a.e_t + (size_t)(bblk + bl) * a.k + q) * a.e_scale : 0.f;
        se1[bl][q] = (have_e1 && q < a.k) ? __ldg(f.e_tm1 + (size_t)(bblk + bl) * a.k + q) * f.e_scale_tm1 : 0.f;
        if (have_e0 || have_e1) {
            for (int idx = threadIdx.x; idx < RP_MAX_OUT * FA_TN; idx += 256) {
                const int qq = idx / FA_TN, ii = idx % FA_TN;
                swo[qq][ii] = qq < a.k ? __ldg(a.W_out + (size_t)qq * N + blockIdx.x * FA_TN + ii) : 0.f;
            }
        }
    }
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    float4 av[2], as[2], ax[2], v[2], Z[2], vm[2], sm[2];
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int b = bblk + l * 16 + ty;
        const int idx = b * N + i0;
        av[l] = *reinterpret_cast<const float4*>(a.adj + idx);
        as[l] = *reinterpret_cast<const float4*>(a.adj + plane + idx);
        ax[l] = NSV > 2 ? *reinterpret_cast<const float4*>(a.adj + 2 * (size_t)plane + idx) : zero;
        v[l] = Z[l] = vm[l] = sm[l] = zero;
        if (do_post) {
            v[l] = __ldg(reinterpret_cast<const float4*>(a.y_t + idx));
            Z[l] = *reinterpret_cast<const float4*>(a.Z + (size_t)b * a.ldz + i0);
        }
        if (do_pre) {
            vm[l] = __ldg(reinterpret_cast<const float4*>(a.y_tm1 + idx));
            if (wg || ro_acc) sm[l] = __ldg(reinterpret_cast<const float4*>(a.y_tm1 + plane + idx));
        }
    }
    __syncthreads();                       // scales / row constants / se0 / se1 / swo are staged (the loads above stay in flight)
    const float sg = s_scale[0], sgT = s_scale[1], ssT = s_scale[2];
    FaRow row[4];
#pragma unroll
    for (int rr = 0; rr < 4; ++rr) {
        row[rr].fa = s_row[4 * tx + rr][0]; row[rr].ds = s_row[4 * tx + rr][1];
        row[rr].dx = SFA ? s_row[4 * tx + rr][2] : 1.f; row[rr].alpha = SFA ? s_row[4 * tx + rr][3] : 0.f; row[rr].c1 = SFA ? s_row[4 * tx + rr][4] : 0.f;
    }
    const int ov = a.out_var;
    const float theta = a.theta, slope = a.slope, dt = a.dt;
    const bool cut = a.zero_after_post != 0;
    float nbmax = 0.f, gabs = 0.f;
#pragma unroll
    for (int l = 0; l < 2; ++l) {
        const int bl = l * 16 + ty;
        const int b = bblk + bl;
        const int idx = b * N + i0;
        // readout terms of the 4 neurons: W_out^T e_t (enters a_t) and W_out^T e_{t-1} (enters the bound when the output is v)
        float ro0[4] = {0.f, 0.f, 0.f, 0.f}, ro1[4] = {0.f, 0.f, 0.f, 0.f};
        if (have_e0) {
#pragma unroll 1
            for (int q = 0; q < a.k; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(&swo[q][4 * tx]);
                const float e0 = se0[bl][q];
                ro0[0] = fmaf(w4.x, e0, ro0[0]); ro0[1] = fmaf(w4.y, e0, ro0[1]); ro0[2] = fmaf(w4.z, e0, ro0[2]); ro0[3] = fmaf(w4.w, e0, ro0[3]);
            }
        }
        if (have_e1 && ov == RP_VAR_V) {
#pragma unroll 1
            for (int q = 0; q < a.k; ++q) {
                const float4 w4 = *reinterpret_cast<const float4*>(&swo[q][4 * tx]);
                const float e1 = se1[bl][q];
                ro1[0] = fmaf(w4.x, e1, ro1[0]); ro1[1] = fmaf(w4.y, e1, ro1[1]); ro1[2] = fmaf(w4.z, e1, ro1[2]); ro1[3] = fmaf(w4.w, e1, ro1[3]);
            }
        }
        float nav[4], nas[4], nax[4], g[4];
#pragma unroll
        for (int rr = 0; rr < 4; ++rr) {
            nav[rr] = f4at(av[l], rr); nas[rr] = f4at(as[l], rr); nax[rr] = f4at(ax[l], rr);
            if (do_post) {
                // adjoint of step t: spike gate on the v row, surrogate coupling of (s, x) into v
                const float vr = f4at(v[l], rr);
                const float gv = vr >= theta ? 0.f : nav[rr];
                const float dq = fmaf(slope, fabsf(vr - theta), 1.0f);
                const float sgr = rcp_approx(dq * dq);                              // Spike.backward, nodes.py:478-481
                const float F = MODEL == RP_LIF ? row[rr].fa : fmaf(row[rr].fa, vr, 1.0f);
                const float coup = SFA ? fmaf(row[rr].alpha, nax[rr], nas[rr]) : nas[rr];
                float n_v = fmaf(gv, F, sgr * coup);
                float n_s = fmaf(nas[rr], row[rr].ds, f4at(Z[l], rr));
                float n_x = SFA ? fmaf(nax[rr], row[rr].dx, -row[rr].c1 * gv) : 0.f;
                if (ov == RP_VAR_V) n_v += ro0[rr]; else if (ov == RP_VAR_S) n_s += ro0[rr]; else if (ov == RP_VAR_X) n_x += ro0[rr];
                if (cut) { n_v = 0.f; n_s = 0.f; n_x = 0.f; }
                nav[rr] = n_v; nas[rr] = n_s; nax[rr] = n_x;
            }
            g[rr] = 0.f;
            if (do_pre) {
                const float vmr = f4at(vm[l], rr);
                const float gn = vmr >= theta ? 0.f : nav[rr];                      // gate_{t-1} a_t[v]
                g[rr] = dt * gn;
                // |a_{t-1}[v]| of this element: the next kernel evaluates the same expression from the same stored values
                const float dq = fmaf(slope, fabsf(vmr - theta), 1.0f);
                const float sgr = rcp_approx(dq * dq);
                const float F = MODEL == RP_LIF ? row[rr].fa : fmaf(row[rr].fa, vmr, 1.0f);
                const float coup = SFA ? fmaf(row[rr].alpha, nax[rr], nas[rr]) : nas[rr];
                float nv = fmaf(gn, F, sgr * coup);
                if (ov == RP_VAR_V) nv += ro1[rr];
                nbmax = fmaxf(nbmax, fabsf(nv));
                gabs = fmaxf(gabs, fabsf(g[rr]));
            }
        }
        if (do_post) {
            *reinterpret_cast<float4*>(a.adj + idx) = make_float4(nav[0], nav[1], nav[2], nav[3]);
            *reinterpret_cast<float4*>(a.adj + plane + idx) = make_float4(nas[0], nas[1], nas[2], nas[3]);
            if (NSV > 2) *reinterpret_cast<float4*>(a.adj + 2 * (size_t)plane + idx) = make_float4(nax[0], nax[1], nax[2], nax[3]);
        }
        if (do_pre) {
            store_split4_f16(f.g_hi, f.g_lo, (size_t)b * f.ld_g + i0, g, sg);
            if (wg || ro_acc) {
#pragma unroll
                for (int rr = 0; rr < 4; ++rr) { tgT[4 * tx + rr][bl] = g[rr]; tsT[4 * tx + rr][bl] = f4at(sm[l], rr); }
            }
        }
    }
    if (do_pre) {
        nbmax = warp_max(nbmax);
        if ((threadIdx.x & 31) == 0 && nbmax > 0.f) atomic_max_nonneg(f.nb_out, dt * nbmax * 1.001f);
        if (!(gabs * sg < 60000.0f)) atomicOr(f.flags, 2);
    }
    trace_end(trace);
    if (wg || ro_acc) {
        __syncthreads();
        const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
        if (wg) {
            // half-warp -> one neuron row, lane -> two consecutive trials: every store instruction writes 2 x 64 contiguous bytes
            const int c = 2 * (lane & 15);
            __half* gth = reinterpret_cast<__half*>(f.gT_hi); __half* gtl = reinterpret_cast<__half*>(f.gT_lo);
            __half* sth = reinterpret_cast<__half*>(f.srcT_hi); __half* stl = reinterpret_cast<__half*>(f.srcT_lo);
#pragma unroll
            for (int j = 0; j < FA_TN / 16; ++j) {
                const int r = 16 * j + 2 * w + (lane >> 4);
                const size_t off = (size_t)(blockIdx.x * FA_TN + r) * f.ld_t + f.t_col0 + bblk + c;
                __half h0, l0, h1, l1;
                split_f16(tgT[r][c] * sgT, h0, l0); split_f16(tgT[r][c + 1] * sgT, h1, l1);
                *reinterpret_cast<__half2*>(gth + off) = __halves2half2(h0, h1);
                *reinterpret_cast<__half2*>(gtl + off) = __halves2half2(l0, l1);
                split_f16(tsT[r][c] * ssT, h0, l0); split_f16(tsT[r][c + 1] * ssT, h1, l1);
                *reinterpret_cast<__half2*>(sth + off) = __halves2half2(h0, h1);
                *reinterpret_cast<__half2*>(stl + off) = __halves2half2(l0, l1);
            }
        }
        if (ro_acc) {
            // dW_out[q][i] += sum_b e_{t-1}[b][q] s_{t-1}[b][i] over the tile's 32 trials: thread (i, trial octet) -> partial sums,
            // summed over the four octets in a fixed order, then one plain read-modify-write per (q, i) of this block's own slice
            const int i = threadIdx.x & 63, grp = threadIdx.x >> 6;
            float acc[RP_MAX_OUT];
#pragma unroll
            for (int q = 0; q < RP_MAX_OUT; ++q) acc[q] = 0.f;
#pragma unroll
            for (int u = 0; u < 8; ++u) {
                const float sv = tsT[i][8 * grp + u];
                const float4 ea = *reinterpret_cast<const float4*>(&se1[8 * grp + u][0]);
                acc[0] = fmaf(ea.x, sv, acc[0]); acc[1] = fmaf(ea.y, sv, acc[1]); acc[2] = fmaf(ea.z, sv, acc[2]); acc[3] = fmaf(ea.w, sv, acc[3]);
                if (a.k > 4) {
                    const float4 eb = *reinterpret_cast<const float4*>(&se1[8 * grp + u][4]);
                    acc[4] = fmaf(eb.x, sv, acc[4]); acc[5] = fmaf(eb.y, sv, acc[5]); acc[6] = fmaf(eb.z, sv, acc[6]); acc[7] = fmaf(eb.w, sv, acc[7]);
                }
            }
            __syncthreads();                       // everyone is done reading tgT
            float* red = &tgT[0][0];               // [4][RP_MAX_OUT][64] = 2048 floats <= 64 * 33
#pragma unroll
            for (int q = 0; q < RP_MAX_OUT; ++q) red[(grp * RP_MAX_OUT + q) * FA_TN + i] = acc[q];
            __syncthreads();
            for (int idx = threadIdx.x; idx < a.k * FA_TN; idx += 256) {
                const int q = idx / FA_TN, ii = idx % FA_TN;
                const float tot = (red[(0 * RP_MAX_OUT + q) * FA_TN + ii] + red[(1 * RP_MAX_OUT + q) * FA_TN + ii]) +
                                  (red[(2 * RP_MAX_OUT + q) * FA_TN + ii] + red[(3 * RP_MAX_OUT + q) * FA_TN + ii]);
                float* dst = f.dwout_part + ((size_t)blockIdx.y * a.k + q) * N + blockIdx.x * FA_TN + ii;
                *dst += tot;
            }
        }
    }
}

// dst[j] = sum over parts (fixed order) of part[p][j]
static __global__ void __launch_bounds__(256) k_sum_parts(const float* __restrict__ part, int nparts, size_t n, float* dst) {
    for (size_t j = (size_t)blockIdx.x * blockDim.x + threadIdx.x; j < n; j += (size_t)gridDim.x * blockDim.x) {
        float acc = 0.f;
        for (int p = 0; p < nparts; ++p) acc += part[(size_t)p * n + j];
        dst[j] = acc;
    }
}
// *p *= v  (one thread)
static __global__ void k_scale_scalar(float* p, float v) { if (threadIdx.x == 0 && blockIdx.x == 0) *p *= v; }

// max over a short device array (the per-step source maxima of the last forward pass) -> *dst
static __global__ void k_max_of_array(const float* src, int n, float* dst) {
    float m = 0.f;
    for (int i = threadIdx.x; i < n; i += 32) m = fmaxf(m, src[i]);
    m = warp_max(m);
    if (threadIdx.x == 0) *dst = m;
}

// ------------------------------------------------------------------------------------------------------
// weight preparation / finalisation
// ------------------------------------------------------------------------------------------------------
// Wk[i][j] = k_i * W[i][j]  and its transpose; optionally also the tf32 hi/lo splits (padded leading dims)
static __global__ void __launch_bounds__(256) k_prepare_weights(int N, const float* __restrict__ W, const float* __restrict__ kp,
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

// binary16 variant: hi/lo of k_i W[i][j] * 2^e and/or of its transpose; e from the tracked maximum of |kW|
static __global__ void __launch_bounds__(256) k_prepare_weights_f16(int N, const float* __restrict__ W, const float* __restrict__ kp, int k_stride, int ldw,
                                                              void* Wk_hi, void* Wk_lo, void* WkT_hi, void* WkT_lo, ScaleRef sc) {
    __shared__ float tile[32][33];
    const float scale = exp2i(scale_expo(sc));
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int r = ty; r < 32; r += 8) {
        const int i = by + r, j = bx + tx;
        float w = 0.f;
        if (i < N && j < N) w = W[(size_t)i * N + j] * __ldg(kp + (size_t)i * k_stride);
        tile[r][tx] = w;
        if (i < N && j < N && Wk_hi) store_split1_f16(Wk_hi, Wk_lo, (size_t)i * ldw + j, w, scale);
    }
    __syncthreads();
    if (WkT_hi) {
        for (int r = ty; r < 32; r += 8) {
            const int j = bx + r, i = by + tx;
            if (i < N && j < N) store_split1_f16(WkT_hi, WkT_lo, (size_t)j * ldw + i, tile[tx][r], scale);
        }
    }
}

// dW_out[q][i] = sum_{t,b} dL/do_t[b][q] * y_t[out][b][i]  over the stored checkpoints (one pass over the history, after the
// reverse sweep; keeps the readout gradient out of the per-step adjoint kernel).
// grid = (ceil(N/128), T, ceil(B/RG_TB)): one step and RG_TB trials per block, 8 independent row loads in flight per thread
constexpr int RG_TB = 128;
template <int MODEL>
__global__ void __launch_bounds__(128) k_readout_grad(int N, int B, int T, int S, int cutoff, int k, int out_var, const float* __restrict__ hist,
                                                       const float* __restrict__ g_out_rec, ModelParams mp, float* dW_out, int t_offset, int T_total) {
    const int i = blockIdx.x * 128 + threadIdx.x;
    const int t = blockIdx.y;
    const int b0 = blockIdx.z * RG_TB, b1 = min(B, b0 + RG_TB);
    const size_t plane = (size_t)B * N, slot = (size_t)HistPlanes<MODEL>::N * plane;
    const PWindow w = pwindow_of(t_offset + t, T_total, S, cutoff);
    if (w.j < 0 || i >= N) return;
    const float sc = 1.0f / (float)w.len;
    const float* yt = hist + (size_t)t * slot + (out_var == RP_VAR_R ? 0 : (size_t)out_var * plane) + i;
    const float* e = g_out_rec + (size_t)w.j * B * k;
    float acc[RP_MAX_OUT];
#pragma unroll
    for (int q = 0; q < RP_MAX_OUT; ++q) acc[q] = 0.f;
    for (int bb = b0; bb < b1; bb += 8) {
        float y[8];
#pragma unroll
        for (int u = 0; u < 8; ++u) y[u] = (bb + u < b1) ? __ldg(yt + (size_t)(bb + u) * N) : 0.f;
#pragma unroll
        for (int u = 0; u < 8; ++u) {
            const int b = bb + u;
            if (b < b1) {
                float yv = y[u];
                if (out_var == RP_VAR_R) { if constexpr (!ModelTraits<MODEL>::SPIKING) yv = rate_act<MODEL>(mp, i, yv, b); }
                for (int q = 0; q < k; ++q) acc[q] = fmaf(__ldg(e + (size_t)b * k + q), yv, acc[q]);
            }
        }
    }
    for (int q = 0; q < k; ++q) atomicAdd(dW_out + (size_t)q * N + i, acc[q] * sc);
}

// dW[i][j] = k_i * dWraw[i][j] ;  dk[i] = sum_j dWraw[i][j] * W[i][j]        (one block per row; 16-byte accesses when N % 4 == 0)
static __global__ void __launch_bounds__(256) k_finish_wgrad(int N, const float* __restrict__ dWraw, int ldr, const float* __restrict__ W,
                                                       const float* __restrict__ kp, int k_stride, float* dW, float* dk, int n_slices) {
    __shared__ float red[9];
    const int i = blockIdx.x;
    const float kv = __ldg(kp + (size_t)i * k_stride);
    float acc = 0.f;
    const bool vec = (N % 4 == 0) && (ldr % 4 == 0) && ((reinterpret_cast<uintptr_t>(dWraw) | reinterpret_cast<uintptr_t>(W) | reinterpret_cast<uintptr_t>(dW)) & 15u) == 0;
    if (vec) {
        for (int j = 4 * threadIdx.x; j < N; j += 4 * blockDim.x) {
            float4 r = __ldcs(reinterpret_cast<const float4*>(dWraw + (size_t)i * ldr + j));
            for (int z = 1; z < n_slices; ++z) {
                const float4 q = __ldcs(reinterpret_cast<const float4*>(dWraw + (size_t)z * N * ldr + (size_t)i * ldr + j));
                r.x += q.x; r.y += q.y; r.z += q.z; r.w += q.w;
            }
            const float4 w = __ldg(reinterpret_cast<const float4*>(W + (size_t)i * N + j));
            if (dW) *reinterpret_cast<float4*>(dW + (size_t)i * N + j) = make_float4(kv * r.x, kv * r.y, kv * r.z, kv * r.w);
            acc = fmaf(r.x, w.x, acc); acc = fmaf(r.y, w.y, acc); acc = fmaf(r.z, w.z, acc); acc = fmaf(r.w, w.w, acc);
        }
    } else {
        for (int j = threadIdx.x; j < N; j += blockDim.x) {
            float r = dWraw[(size_t)i * ldr + j];
            for (int z = 1; z < n_slices; ++z) r += dWraw[(size_t)z * N * ldr + (size_t)i * ldr + j];     // split-K slices
            if (dW) dW[(size_t)i * N + j] = kv * r;
            acc = fmaf(r, W[(size_t)i * N + j], acc);
        }
    }
    if (dk) {
        const float tot = block_sum(acc, red);
        if (threadIdx.x == 0) dk[i] = tot;
    }
}

static __global__ void __launch_bounds__(256) k_fill(float* p, size_t n, float v) {
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < n; idx += (size_t)gridDim.x * blockDim.x) p[idx] = v;
}

}  // namespace rp
