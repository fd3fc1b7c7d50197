// Persistent whole-horizon kernels for few trials (B <= 4): the reference's own shape (B = 1, nodes.py:90).
//
// One cooperative launch integrates all T steps.  Each CTA owns a contiguous block of neurons: their rows of k*W stay
// resident in shared memory for the whole horizon (N=1000: 4 MB / 125 CTAs = 32 KB each) or are streamed from the 126 MB
// L2 when they do not fit (N=4096: 64 MiB); the state of an owned neuron lives in registers of one thread for all T steps.
// Per step: load the source vector r_t (B*N floats, written by all CTAs in the previous step) into shared memory with
// L1-bypassing loads, one warp per row does the dot product with a shuffle reduction, the owning thread applies the vector
// field / threshold / reset, accumulates the Observer window, and publishes r_{t+1} and the checkpoint.
// The per-step all-to-all exchange of the source vector carries its own synchronisation: every element travels as an
// 8-byte {value, step tag} word written with one 64-bit store (the flag-in-data protocol NCCL's LL transport uses), double
// buffered by step parity, and consumers simply re-read an element until its tag is the step they are waiting for.  There
// is no grid barrier, no fence and no atomic on the critical path; a step costs one L2 write + one L2 read (~1.5 us) instead
// of fence + atomic + poll + read (~3.5 us measured with a counter barrier) or ~6 us per kernel launch.  Double buffering
// is sufficient: nobody can publish step t+2 into the slot of step t before every CTA has consumed step t, because
// publishing t+2 needs all of step t+1, which every CTA only produces after it has read step t.
//
// The reverse pass has the same structure with the roles swapped: a CTA owns rows j of (kW)^T and of dW^T, so that both the
// adjoint product Z_j = sum_i kW[i][j] g_i and the rank-1 weight-gradient update dW[i][j] += g_i r_j need only the shared
// vector g_t; dW accumulates on-chip (or in L2) across the whole horizon and is written once.
#pragma once
#include <cuda_runtime.h>
#include "rp_kernels.cuh"

namespace rp {

constexpr int PS_THREADS = 256;
constexpr int PS_MAX_B = 4;
constexpr int PS_MAX_ROWS = 64;      // owned neurons per CTA (rows * B <= 256 threads)

__device__ __forceinline__ unsigned int ld_acquire_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// all CTAs have arrived `target` times in total
__device__ __forceinline__ void grid_wait(const unsigned int* bar, unsigned int target) {
    if (threadIdx.x == 0) {
        while (ld_acquire_u32(bar) < target) { }
    }
    __syncthreads();
}
__device__ __forceinline__ void grid_arrive(unsigned int* bar) {
    __syncthreads();
    if (threadIdx.x == 0) {
        __threadfence();
        atomicAdd(bar, 1u);
    }
}

// ---- flag-in-data exchange ---------------------------------------------------------------------------------------------
__device__ __forceinline__ void ll_store(uint2* p, float v, unsigned int tag) {
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(p), "r"(__float_as_uint(v)), "r"(tag) : "memory");
}
__device__ __forceinline__ uint4 ll_load2(const uint4* p) {
    uint4 q;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(q.x), "=r"(q.y), "=r"(q.z), "=r"(q.w) : "l"(p) : "memory");
    return q;
}
// gather the [B][Npad] vector tagged `tag` from global {value, tag} pairs into shared memory (columns >= N are padding)
__device__ __forceinline__ void ll_gather(const uint2* gsrc, float* s_dst, int B, int N, int Npad, unsigned int tag) {
    const uint4* g4 = reinterpret_cast<const uint4*>(gsrc);
    const int pairs = (B * Npad) >> 1;
    for (int idx = threadIdx.x; idx < pairs; idx += PS_THREADS) {
        int col = 2 * idx;                                  // column of the first element of the pair: (2 idx) mod Npad without a division
        while (col >= Npad) col -= Npad;                    // at most B - 1 <= 3 subtractions
        const bool need0 = col < N, need1 = col + 1 < N;
        uint4 q;
        do { q = ll_load2(g4 + idx); } while ((need0 && q.y != tag) || (need1 && q.w != tag));
        s_dst[2 * idx] = __uint_as_float(q.x);
        s_dst[2 * idx + 1] = __uint_as_float(q.z);
    }
}

struct PersistFwdArgs {
    int N, B, T, m, k, in_mode, in_target, out_mode, out_var, S, cutoff, t_offset, T_total;
    float dt, theta, v_reset;
    const float* Wk; int ldw;       // [N][ldw]  k_i * W
    int rows_per_cta, w_resident;
    const float* x;                 // [T][B][m] | [T][B][N]
    const float* W_in;              // [N][m]
    const float* W_out;             // [k][N]
    ModelParams mp;
    const float* y0;                // [nsv][B][N]
    float* yT;                      // [nsv][B][N]
    float* history;                 // [(T+1)][nsv][B][N] (slot 0 written by the host) or nullptr
    uint2* srcbuf;                  // [2][B][Npad] {value, tag} pairs, tags zeroed by the host before the launch
    int Npad;
    float* out_rec;                 // READOUT: [n_rec][B][k] zero-initialised (atomics) ; DENSE: [n_rec][B][N]
    int n_rec_vars;
    int rec_var[RP_MAX_REC];
    int rec_reduce[RP_MAX_REC];
    float* rec_buf[RP_MAX_REC];     // reduce: [n_rec][B] zero-initialised (atomics); else [n_rec][B][N]
    int rec_post;
    unsigned int* barrier;          // zero-initialised
};

template <int MODEL>
__global__ void __launch_bounds__(PS_THREADS, 1) k_persist_fwd(PersistFwdArgs a) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr bool SPK = ModelTraits<MODEL>::SPIKING;
    extern __shared__ __align__(16) float psm[];
    const int N = a.N, B = a.B, Npad = a.Npad;
    float* s_src = psm;                                  // [B][Npad]
    float* s_u = s_src + B * Npad;                       // [rows][B]
    float* s_red = s_u + PS_MAX_ROWS * PS_MAX_B;         // [B*k + RP_MAX_REC*B] block partials
    float* s_W = s_red + 64;                             // [rows][ldw] when resident
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int R = max(0, min(a.rows_per_cta, N - r0));
    const size_t plane = (size_t)B * N;

    if (a.w_resident) {
        for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) s_W[idx] = a.Wk[(size_t)r0 * a.ldw + idx];
    }
    if (tid < 64) s_red[tid] = 0.f;

    // element ownership: thread e -> (row r, trial b)
    const bool own = tid < R * B;
    const int r = own ? tid % R : 0, b = own ? tid / R : 0;
    const int i = r0 + r;
    float v = 0.f, s = 0.f, x = 0.f, win_sum = 0.f;
    float w_in[RP_MAX_IN];
    FwdStepArgs fa;      // only the fields fwd_elem reads
    fa.dt = a.dt; fa.theta = a.theta; fa.v_reset = a.v_reset; fa.in_target = a.in_target; fa.mp = a.mp;
    if (own) {
        const size_t idx = (size_t)b * N + i;
        v = a.y0[idx];
        if (NSV > 1) s = a.y0[plane + idx];
        if (NSV > 2) x = a.y0[2 * plane + idx];
#pragma unroll
        for (int j = 0; j < RP_MAX_IN; ++j) w_in[j] = (a.in_mode == RP_IN_PROJ && j < a.m) ? a.W_in[(size_t)i * a.m + j] : 0.f;
    }
    // per-neuron constants of the owned neuron, loaded once for the whole horizon (templates other than ik: reciprocal form of the step,
    // the same arithmetic as the fused tensor-core epilogue, <= 1 ulp per term away from the divisions of fwd_elem)
    FwdRow frow{1.f, 0.f, 1.f, 1.f, 0.f, 0.f, 0.f};
    FwdStepArgs fa_fast = fa;
    fa_fast.in_mode = RP_IN_DENSE;                       // the input current is computed above the exchange and passed in as a dense value
    fa_fast.per_trial = 0;
    const bool fast_elem = !is_ik(MODEL) && a.mp.bstride[RP_P_TAU] == 0 && a.mp.bstride[RP_P_ETA] == 0 && a.mp.bstride[RP_P_TAU_S] == 0 &&
                           a.mp.bstride[RP_P_TAU_X] == 0 && a.mp.bstride[RP_P_ALPHA] == 0;
    if constexpr (!is_ik(MODEL)) { if (own && fast_elem) frow = fwd_row<MODEL>(fa_fast, i); }
    // record-window bookkeeping without a division per step: [w_start, w_rec] is the window that contains (or follows) the current step
    const int S_ = max(a.S, 1);
    const int w_r0 = ((a.cutoff + S_ - 1) / S_) * S_;
    int w_j, w_start, w_rec;
    if (a.t_offset <= w_r0) { w_j = 0; w_start = a.cutoff; w_rec = w_r0; }
    else { w_j = (a.t_offset - w_r0 + S_ - 1) / S_; w_rec = w_r0 + w_j * S_; w_start = w_rec - S_ + 1; }
    bool any_reduce = false;
    for (int q = 0; q < a.n_rec_vars; ++q) any_reduce = any_reduce || (a.rec_reduce[q] != 0);
    const int nvec = N >> 2;
    if (own && a.T > 0) {      // publish r_0 (tag 1) into slot 0
        float src0;
        if constexpr (SPK) src0 = s; else src0 = rate_act<MODEL>(a.mp, i, v, b);
        ll_store(a.srcbuf + (size_t)b * Npad + i, src0, 1u);
    }
    __syncthreads();

    for (int t = 0; t < a.T; ++t) {
        // 0) input current of this step: independent of the exchange, so its global-memory latency is paid while the gather spins
        float Iin = 0.f;
        if (own) {
            if (a.in_mode == RP_IN_DENSE) Iin = __ldg(a.x + (size_t)t * plane + (size_t)b * N + i);
            else if (a.in_mode == RP_IN_PROJ) {
                const float* xt = a.x + ((size_t)t * B + b) * a.m;
#pragma unroll
                for (int j = 0; j < RP_MAX_IN; ++j) if (j < a.m) Iin = fmaf(w_in[j], __ldg(xt + j), Iin);
            }
        }
        // 1) source vector of this step (tag t+1) -> shared memory; spins per element until every producer has published
        ll_gather(a.srcbuf + (size_t)(t & 1) * B * Npad, s_src, B, N, Npad, (unsigned int)(t + 1));
        __syncthreads();
        // 2) recurrent drive of the owned rows: one warp per row, lanes stride the columns
        for (int rr = warp; rr < R; rr += PS_THREADS / 32) {
            const float* wrow = a.w_resident ? s_W + (size_t)rr * a.ldw : a.Wk + (size_t)(r0 + rr) * a.ldw;
            float acc[PS_MAX_B];
#pragma unroll
            for (int q = 0; q < PS_MAX_B; ++q) acc[q] = 0.f;
            for (int kv = lane; kv < nvec; kv += 32) {
                const float4 w4 = a.w_resident ? reinterpret_cast<const float4*>(wrow)[kv] : __ldg(reinterpret_cast<const float4*>(wrow) + kv);
#pragma unroll
                for (int q = 0; q < PS_MAX_B; ++q) {
                    if (q < B) {
                        const float4 s4 = reinterpret_cast<const float4*>(s_src + q * Npad)[kv];
                        acc[q] = fmaf(w4.x, s4.x, acc[q]); acc[q] = fmaf(w4.y, s4.y, acc[q]);
                        acc[q] = fmaf(w4.z, s4.z, acc[q]); acc[q] = fmaf(w4.w, s4.w, acc[q]);
                    }
                }
            }
            const int kt = (nvec << 2) + lane;              // ragged tail (N % 4 columns)
            if (kt < N) {
#pragma unroll
                for (int q = 0; q < PS_MAX_B; ++q) if (q < B) acc[q] = fmaf(wrow[kt], s_src[q * Npad + kt], acc[q]);
            }
#pragma unroll
            for (int q = 0; q < PS_MAX_B; ++q) {
                if (q < B) {
                    float tot = acc[q];
                    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    if (lane == 0) s_u[rr * PS_MAX_B + q] = tot;
                }
            }
        }
        __syncthreads();
        // 3) vector field, threshold/reset, Observer -- the owning thread keeps the neuron's state in registers
        const int tg = a.t_offset + t;
        if (tg > w_rec) { ++w_j; w_start = w_rec + 1; w_rec += S_; }             // steps advance by one: at most one window per step
        PWindow w{-1, 0, 0, 0};
        if (tg >= a.cutoff && w_rec < a.T_total) { w.j = w_j; w.first = (tg == w_start); w.close = (tg == w_rec); w.len = w_rec - w_start + 1; }
        if (own) {
            const size_t idx = (size_t)b * N + i;
            float v1, s1, x1;
            const float urec = s_u[r * PS_MAX_B + b];
            bool done = false;
            if constexpr (!is_ik(MODEL)) {
                if (fast_elem) { fwd_elem_fast<MODEL>(fa_fast, frow, i, b, urec, 0.f, 0.f, Iin, v, s, x, v1, s1, x1); done = true; }
            }
            if (!done) fwd_elem<MODEL>(fa, i, urec, Iin, v, s, x, v1, s1, x1, b);
            float src1;
            if constexpr (SPK) src1 = s1; else src1 = rate_act<MODEL>(a.mp, i, v1, b);
            if (t + 1 < a.T) ll_store(a.srcbuf + (size_t)((t + 1) & 1) * B * Npad + (size_t)b * Npad + i, src1, (unsigned int)(t + 2));
            if (a.history) {
                constexpr int NH = HistPlanes<MODEL>::N;
                float* h = a.history + (size_t)(t + 1) * NH * plane + idx;
                h[0] = v1;
                if (NSV > 1) h[plane] = s1;
                if (NSV > 2) h[2 * plane] = x1;
                if (NH > NSV) a.history[(size_t)t * NH * plane + (size_t)NSV * plane + idx] = urec;   // drive of step t, slot t
            }
            if (w.j >= 0) {
                if (a.out_rec) {
                    float yout;
                    if (a.out_var == RP_VAR_V) yout = v; else if (a.out_var == RP_VAR_S) yout = s;
                    else if (a.out_var == RP_VAR_X) yout = x;
                    else { if constexpr (!SPK) yout = rate_act<MODEL>(a.mp, i, v, b); else yout = 0.f; }
                    if (a.out_mode == RP_OUT_DENSE) {
                        win_sum = w.first ? yout : win_sum + yout;
                        if (w.close) a.out_rec[((size_t)w.j * B + b) * N + i] = win_sum / (float)w.len;
                    } else {
                        for (int q = 0; q < a.k; ++q) atomicAdd(&s_red[b * a.k + q], __ldg(a.W_out + (size_t)q * N + i) * yout);
                    }
                }
                if (w.close) {
                    for (int q = 0; q < a.n_rec_vars; ++q) {
                        const int var = a.rec_var[q];
                        const float val = a.rec_post ? (var == 0 ? v1 : (var == 1 ? s1 : x1)) : (var == 0 ? v : (var == 1 ? s : x));
                        if (a.rec_reduce[q]) atomicAdd(&s_red[32 + q * PS_MAX_B + b], val);
                        else a.rec_buf[q][((size_t)w.j * B + b) * N + i] = val;
                    }
                }
            }
            v = v1; s = s1; x = x1;
        }
        // block-level reduction of this step's readout / neuron-mean contributions: only on steps that produced any (uniform condition)
        if (w.j >= 0 && ((a.out_rec && a.out_mode == RP_OUT_READOUT) || (w.close && any_reduce))) {
            __syncthreads();
            if (a.out_rec && a.out_mode == RP_OUT_READOUT && tid < B * a.k) {
                atomicAdd(a.out_rec + (size_t)w.j * B * a.k + tid, s_red[tid] / (float)w.len);
                s_red[tid] = 0.f;
            }
            if (w.close && tid >= 32 && tid < 32 + RP_MAX_REC * PS_MAX_B) {
                const int q = (tid - 32) / PS_MAX_B, bb = (tid - 32) % PS_MAX_B;
                if (q < a.n_rec_vars && a.rec_reduce[q] && bb < B) {
                    atomicAdd(a.rec_buf[q] + (size_t)w.j * B + bb, s_red[tid] / (float)N);
                    s_red[tid] = 0.f;
                }
            }
        }
        __syncthreads();          // s_src / s_u / s_red are rewritten by the next step
    }
    if (own) {
        const size_t idx = (size_t)b * N + i;
        a.yT[idx] = v;
        if (NSV > 1) a.yT[plane + idx] = s;
        if (NSV > 2) a.yT[2 * plane + idx] = x;
    }
}

// ---------------------------------------------------------------------------------------------------------------
// persistent reverse pass
// ---------------------------------------------------------------------------------------------------------------
struct PersistBwdArgs {
    int N, B, T, m, k, in_mode, in_target, out_mode, out_var, S, cutoff, truncate, t_offset, T_total;
    float dt, theta, slope;
    const float* WkT; int ldw;      // [N][ldw]  (k_i W_ij)^T : row j holds column j of kW
    int rows_per_cta, w_resident, dw_resident, need_dW;
    const float* x; const float* W_in; const float* W_out;
    ModelParams mp;
    const float* history;           // [(T+1)][nsv][B][N]
    const float* g_out_rec;         // [n_rec][B][k] | [n_rec][B][N] or nullptr
    const float* g_yT;              // [nsv][B][N] or nullptr
    uint2* gbuf;                    // [2][B][Npad] {value, tag} pairs, tags zeroed by the host before the launch
    int Npad;
    float* dWrawT;                  // [N][ldw]  dWraw^T (row j, col i), zero-initialised when not resident; written at the end
    float* dparams[RP_NUM_PARAMS];  // [N] each (plain stores: one owner thread per neuron and trial -> atomics only across trials)
    float* dW_in;                   // [N][m]
    float* dW_out;                  // [k][N]
    float* g_y0;                    // [nsv][B][N] or nullptr
    float* g_x;                     // [T][B][N] or nullptr
    unsigned int* barrier;
};

template <int MODEL>
__global__ void __launch_bounds__(PS_THREADS, 1) k_persist_bwd(PersistBwdArgs a) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr bool SPK = ModelTraits<MODEL>::SPIKING;
    extern __shared__ __align__(16) float psm[];
    const int N = a.N, B = a.B, Npad = a.Npad;
    float* s_g = psm;                                    // [B][Npad]  g_t of all neurons
    float* s_z = s_g + B * Npad;                         // [rows][B]
    float* s_src = s_z + PS_MAX_ROWS * PS_MAX_B;         // [rows][B]  r_t of the owned neurons (for the rank-1 update)
    float* s_W = s_src + PS_MAX_ROWS * PS_MAX_B;         // [rows][ldw] when resident
    float* s_dW = s_W + (a.w_resident ? (size_t)a.rows_per_cta * a.ldw : 0);   // [rows][ldw] when resident
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = blockIdx.x * a.rows_per_cta;
    const int R = max(0, min(a.rows_per_cta, N - r0));
    const size_t plane = (size_t)B * N, slot = (size_t)HistPlanes<MODEL>::N * plane;

    if (a.w_resident) for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) s_W[idx] = a.WkT[(size_t)r0 * a.ldw + idx];
    if (a.need_dW && a.dw_resident) for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) s_dW[idx] = 0.f;

    const bool own = tid < R * B;
    const int r = own ? tid % R : 0, b = own ? tid / R : 0;
    const int i = r0 + r;                                 // owned neuron (called j in the header comment)
    float av = 0.f, as = 0.f, ax = 0.f;                   // adjoint of (v, s, x) at t+1
    float acc[ADJ_NACC];                                  // parameter / edge gradient sums of the owned neuron
#pragma unroll
    for (int q = 0; q < ADJ_NACC; ++q) acc[q] = 0.f;
    // the shared adjoint arithmetic (rp_kernels.cuh) reads its constants from an AdjArgs record
    AdjArgs aa;
    aa.N = N; aa.B = B; aa.m = a.m; aa.k = a.k; aa.in_mode = a.in_mode; aa.in_target = a.in_target;
    aa.out_mode = a.out_mode; aa.out_var = a.out_var; aa.dt = a.dt; aa.theta = a.theta; aa.slope = a.slope;
    aa.W_in = a.W_in; aa.W_out = a.W_out; aa.mp = a.mp; aa.dW_in = a.dW_in; aa.dW_out = a.dW_out;
    for (int q = 0; q < RP_NUM_PARAMS; ++q) aa.dparams[q] = a.dparams[q];
    aa.x_t = nullptr; aa.e_t = nullptr; aa.e_scale = 0.f; aa.zero_after_post = 0; aa.per_trial = 0;
    AdjRowParams rowp{1.f, 1.f, 1.f, 0.f};
    if (own) {
        const size_t idx = (size_t)b * N + i;
        if (a.g_yT) { av = a.g_yT[idx]; if (NSV > 1) as = a.g_yT[plane + idx]; if (NSV > 2) ax = a.g_yT[2 * plane + idx]; }
        rowp = adj_row_params<MODEL>(aa, i, b);
    }
    const int nvec = N >> 2;
    const bool truncating = a.truncate > 0 && a.truncate < a.T_total;

    // g_{T-1} = dt * gate_{T-1} * a_T  ("pre" of the first reverse step)
    auto make_g = [&](int tm1) -> float {
        const size_t idx = (size_t)b * N + i;
        const float vm = a.history[(size_t)tm1 * slot + idx];
        float g, srcv;
        adj_pre_math<MODEL>(aa, i, av, vm, 0.f, g, srcv, b);
        return g;
    };
    // reverse step t consumes g_t carrying tag T - t (1, 2, ... as t runs down)
    if (own && a.T > 0) ll_store(a.gbuf + (size_t)((a.T - 1) & 1) * B * Npad + (size_t)b * Npad + i, make_g(a.T - 1), 1u);
    __syncthreads();

    // record-window bookkeeping without a division per step (steps run downwards): [w_start, w_rec] contains or follows the current step
    const int S_ = max(a.S, 1);
    const int w_r0 = ((a.cutoff + S_ - 1) / S_) * S_;
    int w_j, w_start, w_rec;
    {
        const int tg_last = a.t_offset + a.T - 1;
        if (tg_last <= w_r0) { w_j = 0; w_start = a.cutoff; w_rec = w_r0; }
        else { w_j = (tg_last - w_r0 + S_ - 1) / S_; w_rec = w_r0 + w_j * S_; w_start = w_rec - S_ + 1; }
    }
    for (int t = a.T - 1; t >= 0; --t) {
        // checkpoint y_t of the owned neuron: independent of the exchange, loaded while the gather spins
        float v = 0.f, s = 0.f, x = 0.f, urec = 0.f, vm_prev = 0.f;
        if (own) {
            const size_t idx = (size_t)b * N + i;
            const float* yt = a.history + (size_t)t * slot;
            v = __ldg(yt + idx);
            if (NSV > 1) s = __ldg(yt + plane + idx);
            if (NSV > 2) x = __ldg(yt + 2 * plane + idx);
            if (HistPlanes<MODEL>::N > NSV) urec = __ldg(yt + (size_t)NSV * plane + idx);
            if (t > 0) vm_prev = __ldg(a.history + (size_t)(t - 1) * slot + idx);      // v_{t-1}: gate of g_{t-1}, needed right before the publish
        }
        ll_gather(a.gbuf + (size_t)(t & 1) * B * Npad, s_g, B, N, Npad, (unsigned int)(a.T - t));
        // source value r_t of the owned neurons (rank-1 update operand)
        if (own && a.need_dW) {
            float rv;
            if constexpr (SPK) rv = s; else rv = rate_act<MODEL>(a.mp, i, v, b);
            s_src[r * PS_MAX_B + b] = rv;
        }
        __syncthreads();
        // Z_j = sum_i (kW)^T[j][i] g_i   and   dWraw^T[j][i] += r_j * g_i
        for (int rr = warp; rr < R; rr += PS_THREADS / 32) {
            const float* wrow = a.w_resident ? s_W + (size_t)rr * a.ldw : a.WkT + (size_t)(r0 + rr) * a.ldw;
            float* drow = a.dw_resident ? s_dW + (size_t)rr * a.ldw : a.dWrawT + (size_t)(r0 + rr) * a.ldw;
            float acc[PS_MAX_B], rj[PS_MAX_B];
#pragma unroll
            for (int q = 0; q < PS_MAX_B; ++q) { acc[q] = 0.f; rj[q] = (a.need_dW && q < B) ? s_src[rr * PS_MAX_B + q] : 0.f; }
            for (int kv = lane; kv < nvec; kv += 32) {
                const float4 w4 = a.w_resident ? reinterpret_cast<const float4*>(wrow)[kv] : __ldg(reinterpret_cast<const float4*>(wrow) + kv);
                float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                for (int q = 0; q < PS_MAX_B; ++q) {
                    if (q < B) {
                        const float4 g4 = reinterpret_cast<const float4*>(s_g + q * Npad)[kv];
                        acc[q] = fmaf(w4.x, g4.x, acc[q]); acc[q] = fmaf(w4.y, g4.y, acc[q]);
                        acc[q] = fmaf(w4.z, g4.z, acc[q]); acc[q] = fmaf(w4.w, g4.w, acc[q]);
                        d4.x = fmaf(rj[q], g4.x, d4.x); d4.y = fmaf(rj[q], g4.y, d4.y);
                        d4.z = fmaf(rj[q], g4.z, d4.z); d4.w = fmaf(rj[q], g4.w, d4.w);
                    }
                }
                if (a.need_dW) {
                    float4 o = reinterpret_cast<float4*>(drow)[kv];
                    o.x += d4.x; o.y += d4.y; o.z += d4.z; o.w += d4.w;
                    reinterpret_cast<float4*>(drow)[kv] = o;
                }
            }
            const int kt = (nvec << 2) + lane;
            if (kt < N) {
                float d = 0.f;
#pragma unroll
                for (int q = 0; q < PS_MAX_B; ++q) if (q < B) { const float gq = s_g[q * Npad + kt]; acc[q] = fmaf(wrow[kt], gq, acc[q]); d = fmaf(rj[q], gq, d); }
                if (a.need_dW) drow[kt] += d;
            }
#pragma unroll
            for (int q = 0; q < PS_MAX_B; ++q) {
                if (q < B) {
                    float tot = acc[q];
                    for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                    if (lane == 0) s_z[rr * PS_MAX_B + q] = tot;
                }
            }
        }
        __syncthreads();
        // adjoint of step t for the owned neuron, then g_{t-1}
        if (own) {
            const size_t idx = (size_t)b * N + i;
            const int tg = a.t_offset + t;
            if (tg < w_start && w_j > 0) { --w_j; w_rec = w_start - 1; w_start = (w_j == 0) ? a.cutoff : w_rec - S_ + 1; }
            PWindow w{-1, 0, 0, 0};
            if (tg >= a.cutoff && w_rec < a.T_total) { w.j = w_j; w.first = (tg == w_start); w.close = (tg == w_rec); w.len = w_rec - w_start + 1; }
            const size_t estride = a.out_mode == RP_OUT_READOUT ? (size_t)B * a.k : plane;
            aa.e_t = (a.g_out_rec && w.j >= 0) ? a.g_out_rec + (size_t)w.j * estride : nullptr;
            aa.e_scale = w.j >= 0 ? 1.0f / (float)w.len : 0.f;
            aa.x_t = a.x ? a.x + (size_t)t * (a.in_mode == RP_IN_DENSE ? plane : (size_t)B * a.m) : nullptr;
            aa.zero_after_post = (truncating && (a.t_offset + t) > 0 && (a.t_offset + t) % a.truncate == 0) ? 1 : 0;
            const RegAcc racc{acc};
            const float dI = adj_post_math<MODEL>(aa, rowp, racc, i, b, s_z[r * PS_MAX_B + b], v, s, x, av, as, ax, urec);
            if (a.g_x) a.g_x[(size_t)t * plane + idx] = dI;
            if (t > 0) {
                float gm, srcv;
                adj_pre_math<MODEL>(aa, i, av, vm_prev, 0.f, gm, srcv, b);
                ll_store(a.gbuf + (size_t)((t - 1) & 1) * B * Npad + (size_t)b * Npad + i, gm, (unsigned int)(a.T - t + 1));
            }
        }
        __syncthreads();          // s_g / s_z / s_src are rewritten by the next step
    }

    // write back: adjoint of y0, parameter / edge gradients (one thread per neuron and trial -> atomics across trials)
    if (own) {
        const size_t idx = (size_t)b * N + i;
        if (a.g_y0) { a.g_y0[idx] = av; if (NSV > 1) a.g_y0[plane + idx] = as; if (NSV > 2) a.g_y0[2 * plane + idx] = ax; }
        for (int q = 0; q < ADJ_NACC; ++q) {
            size_t off = 0;
            float* dst = adj_acc_dst(aa, q, i, off);
            if (dst != nullptr) atomicAdd(dst + off, acc[q]);
        }
    }
    if (a.need_dW && a.dw_resident) {
        __syncthreads();
        for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) a.dWrawT[(size_t)r0 * a.ldw + idx] = s_dW[idx];
    }
}

// dWraw^T -> dW = diag(k) dWraw and dk  (persistent path keeps the weight gradient transposed: row = source neuron j)
__global__ void __launch_bounds__(256) k_finish_wgrad_T(int N, const float* __restrict__ dWrawT, int ldr, const float* __restrict__ W,
                                                        const float* __restrict__ kp, int k_stride, float* dW, float* dk) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx: i base, by: j base (rows of dWrawT)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) {
        const int j = by + rr, i = bx + tx;
        tile[rr][tx] = (i < N && j < N) ? dWrawT[(size_t)j * ldr + i] : 0.f;
    }
    __syncthreads();
    for (int rr = ty; rr < 32; rr += 8) {
        const int i = bx + rr, j = by + tx;
        float contrib = 0.f;
        if (i < N && j < N) {
            const float raw = tile[tx][rr];
            if (dW) dW[(size_t)i * N + j] = __ldg(kp + (size_t)i * k_stride) * raw;
            contrib = raw * W[(size_t)i * N + j];
        }
        if (dk) {
            for (int o = 16; o > 0; o >>= 1) contrib += __shfl_xor_sync(0xffffffffu, contrib, o);
            if (tx == 0 && i < N) atomicAdd(dk + i, contrib);
        }
    }
}

}  // namespace rp
