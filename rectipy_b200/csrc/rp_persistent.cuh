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

// Round 2: two-dimensional decomposition for more than a handful of trials.  CTA (rb, tb) owns a block of neuron rows AND a block
// of BL <= PS_MAX_B trials; it gathers only the source vectors of its own trials (the all-to-all volume per CTA shrinks by the
// number of trial blocks) while its rows of kW serve all of them.  With more than 4 trials per CTA the row x trial products run
// as register tiles (PS_RG rows x PS_BT trials per warp and pass over the columns: 12 shared-memory vector loads per 128 FMAs
// instead of 9 per 32), reduced across the lanes with a halving butterfly (31 shuffles for the 32 sums).
// The readout and neuron-mean records no longer use floating-point atomics: every CTA reduces its rows in a fixed order, keeps
// its own window sum, and stores one partial per record; a final pass adds the row blocks in order (bit-reproducible runs).
constexpr int PS_THREADS = 256;
constexpr int PS_MAX_B = 8;          // trials per CTA
constexpr int PS_MAX_ROWS = 64;      // owned neurons per CTA (rows * trials-per-CTA <= 256 threads)
constexpr int PS_RG = 8, PS_BT = 4;  // register tile of the blocked product

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
// All polling loads of a thread are issued before the first tag is examined (a mismatch re-polls that word only): one L2 round
// trip per batch instead of one per word -- with the words polled one after the other the gather cost 16 round trips per step at
// 8 trials per CTA (11.5 us per step, N = 1000, 32 trials) and 2 at the reference's own shape (C1).
__device__ __forceinline__ void ll_gather(const uint2* gsrc, float* s_dst, int B, int N, int Npad, unsigned int tag) {
    const uint4* g4 = reinterpret_cast<const uint4*>(gsrc);
    const int half = Npad >> 1;                             // pairs per trial row (Npad is a multiple of 4)
    if (B <= 2) {
        constexpr int U = 8;                                // column iterations in flight
        for (int bq = 0; bq < B; ++bq) {
            for (int cb = threadIdx.x; cb < half; cb += PS_THREADS * U) {
                uint4 q[U];
#pragma unroll
                for (int u = 0; u < U; ++u) { const int c2 = cb + u * PS_THREADS; if (c2 < half) q[u] = ll_load2(g4 + bq * half + c2); }
#pragma unroll
                for (int u = 0; u < U; ++u) {
                    const int c2 = cb + u * PS_THREADS;
                    if (c2 < half) {
                        const int col = 2 * c2, idx = bq * half + c2;
                        const bool need0 = col < N, need1 = col + 1 < N;
                        while ((need0 && q[u].y != tag) || (need1 && q[u].w != tag)) q[u] = ll_load2(g4 + idx);
                        s_dst[2 * idx] = __uint_as_float(q[u].x);
                        s_dst[2 * idx + 1] = __uint_as_float(q[u].z);
                    }
                }
            }
        }
    } else {
        for (int c2 = threadIdx.x; c2 < half; c2 += PS_THREADS) {
            const int col = 2 * c2;
            const bool need0 = col < N, need1 = col + 1 < N;
            uint4 q[PS_MAX_B];
#pragma unroll
            for (int bq = 0; bq < PS_MAX_B; ++bq) if (bq < B) q[bq] = ll_load2(g4 + bq * half + c2);
#pragma unroll
            for (int bq = 0; bq < PS_MAX_B; ++bq) {
                if (bq < B) {
                    const int idx = bq * half + c2;
                    while ((need0 && q[bq].y != tag) || (need1 && q[bq].w != tag)) q[bq] = ll_load2(g4 + idx);
                    s_dst[2 * idx] = __uint_as_float(q[bq].x);
                    s_dst[2 * idx + 1] = __uint_as_float(q[bq].z);
                }
            }
        }
    }
}

// sum of s_con[base + rr], rr in [0, R), R <= 64, over the lanes of one warp in a fixed order (same tree every run)
__device__ __forceinline__ float warp_ordered_sum(const float* s_con, int base, int R, int lane) {
    float v = (lane < R ? s_con[base + lane] : 0.f) + (lane + 32 < R ? s_con[base + lane + 32] : 0.f);
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// 32 per-lane partial sums v[0..31] -> lane L returns the total (over the 32 lanes) of element L: halving butterfly, 31 shuffles
__device__ __forceinline__ float lane_transpose_reduce(float (&v)[32], int lane) {
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int j = 0; j < off; ++j) {
            const float send = upper ? v[j] : v[j + off];
            const float keep = upper ? v[j + off] : v[j];
            v[j] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
    return v[0];
}

// out[rr][q] = sum_k W[row rbase+rr][k] * vec[trial bbase+q][k] for one PS_RG x PS_BT unit of a warp; lane L ends with element
// (rr, q) = (L / PS_BT, L % PS_BT).  rows >= R and trials >= Bl contribute zeros.
__device__ __forceinline__ float blocked_unit(const float* s_W, const float* gW, int ldw, bool w_resident, const float* s_vec, int Npad,
                                              int N, int rbase, int R, int bbase, int Bl, int lane) {
    static_assert(PS_RG * PS_BT == 32, "one reduced element per lane");
    float acc[32];
#pragma unroll
    for (int e = 0; e < 32; ++e) acc[e] = 0.f;
    const int nvec = N >> 2;
    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int kv = lane; kv < nvec; kv += 32) {
        float4 s4[PS_BT];
#pragma unroll
        for (int q = 0; q < PS_BT; ++q) s4[q] = (bbase + q < Bl) ? reinterpret_cast<const float4*>(s_vec + (bbase + q) * Npad)[kv] : zero;
#pragma unroll
        for (int rr = 0; rr < PS_RG; ++rr) {
            if (rbase + rr < R) {
                const float4 w4 = w_resident ? reinterpret_cast<const float4*>(s_W + (size_t)(rbase + rr) * ldw)[kv]
                                             : __ldg(reinterpret_cast<const float4*>(gW + (size_t)(rbase + rr) * ldw) + kv);
#pragma unroll
                for (int q = 0; q < PS_BT; ++q) {
                    float t = acc[rr * PS_BT + q];
                    t = fmaf(w4.x, s4[q].x, t); t = fmaf(w4.y, s4[q].y, t); t = fmaf(w4.z, s4[q].z, t); t = fmaf(w4.w, s4[q].w, t);
                    acc[rr * PS_BT + q] = t;
                }
            }
        }
    }
    const int kt = (nvec << 2) + lane;                      // ragged tail (N % 4 columns)
    if (kt < N) {
#pragma unroll
        for (int rr = 0; rr < PS_RG; ++rr) {
            if (rbase + rr < R) {
                const float wv = w_resident ? s_W[(size_t)(rbase + rr) * ldw + kt] : __ldg(gW + (size_t)(rbase + rr) * ldw + kt);
#pragma unroll
                for (int q = 0; q < PS_BT; ++q) if (bbase + q < Bl) acc[rr * PS_BT + q] = fmaf(wv, s_vec[(bbase + q) * Npad + kt], acc[rr * PS_BT + q]);
            }
        }
    }
    return lane_transpose_reduce(acc, lane);
}

struct PersistFwdArgs {
    int N, B, T, m, k, in_mode, in_target, out_mode, out_var, S, cutoff, t_offset, T_total;
    float dt, theta, v_reset;
    const float* Wk; int ldw;       // [N][ldw]  k_i * W
    int rows_per_cta, w_resident;
    int n_rb, BL;                   // grid = n_rb row blocks x ceil(B / BL) trial blocks; CTA = (blockIdx.x % n_rb, blockIdx.x / n_rb)
    const float* x;                 // [T][B][m] | [T][B][N]
    const float* W_in;              // [N][m]
    const float* W_out;             // [k][N]
    ModelParams mp;
    const float* y0;                // [nsv][B][N]
    float* yT;                      // [nsv][B][N]
    float* history;                 // [(T+1)][nsv][B][N] (slot 0 written by the host) or nullptr
    uint2* srcbuf;                  // [2][B][Npad] {value, tag} pairs, tags zeroed by the host before the launch
    int Npad;
    float* out_rec;                 // READOUT without out_part: [n_rec][B][k] zero-initialised (atomics) ; DENSE: [n_rec][B][N]
    float* out_part;                // READOUT: [n_rec][n_rb][B][k] per-row-block partial window means (plain stores), or nullptr
    int n_rec_vars;
    int rec_var[RP_MAX_REC];
    int rec_reduce[RP_MAX_REC];
    float* rec_buf[RP_MAX_REC];     // reduce without rec_part: [n_rec][B] zero-initialised (atomics); else [n_rec][B][N]
    float* rec_part[RP_MAX_REC];    // reduce: [n_rec][n_rb][B] per-row-block partial sums (plain stores), or nullptr
    int rec_post;
    unsigned int* barrier;          // zero-initialised
};

// floats of shared memory the forward kernel needs besides the resident rows of kW
__host__ __device__ inline size_t ps_fwd_base_floats(int BL, int Npad) {
    return (size_t)BL * Npad + PS_MAX_ROWS * PS_MAX_B + (RP_MAX_OUT + RP_MAX_REC) * PS_THREADS + PS_MAX_B * RP_MAX_OUT + 64;
}

template <int MODEL>
__global__ void __launch_bounds__(PS_THREADS, 1) k_persist_fwd(PersistFwdArgs a) {
    constexpr int NSV = ModelTraits<MODEL>::NSV;
    constexpr bool SPK = ModelTraits<MODEL>::SPIKING;
    extern __shared__ __align__(16) float psm[];
    const int N = a.N, B = a.B, Npad = a.Npad;
    const int rb = blockIdx.x % a.n_rb, tb = blockIdx.x / a.n_rb;
    const int b0 = tb * a.BL;
    const int Bl = max(0, min(a.BL, B - b0));            // trials of this CTA
    float* s_src = psm;                                  // [BL][Npad]
    float* s_u = s_src + a.BL * Npad;                    // [rows][PS_MAX_B]
    float* s_con = s_u + PS_MAX_ROWS * PS_MAX_B;         // [RP_MAX_OUT + RP_MAX_REC][PS_THREADS] per-element contributions of a step
    float* s_win = s_con + (RP_MAX_OUT + RP_MAX_REC) * PS_THREADS;   // [BL * k] open record window of the readout
    float* s_W = s_win + PS_MAX_B * RP_MAX_OUT + 64;     // [rows][ldw] when resident
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = rb * a.rows_per_cta;
    const int R = max(0, min(a.rows_per_cta, N - r0));
    const size_t plane = (size_t)B * N;

    if (a.w_resident) {
        for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) s_W[idx] = a.Wk[(size_t)r0 * a.ldw + idx];
    }
    if (tid < PS_MAX_B * RP_MAX_OUT) s_win[tid] = 0.f;

    // element ownership: thread e -> (row r, local trial bl); b = global trial
    const bool own = tid < R * Bl;
    const int r = own ? tid % R : 0, bl = own ? tid / R : 0;
    const int b = b0 + bl;
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
    // readout weights of the owned neuron
    float w_out[RP_MAX_OUT];
#pragma unroll
    for (int q = 0; q < RP_MAX_OUT; ++q) w_out[q] = (own && a.out_rec && a.out_mode == RP_OUT_READOUT && q < a.k) ? __ldg(a.W_out + (size_t)q * N + i) : 0.f;
    // record-window bookkeeping without a division per step: [w_start, w_rec] is the window that contains (or follows) the current step
    const int S_ = max(a.S, 1);
    const int w_r0 = ((a.cutoff + S_ - 1) / S_) * S_;
    int w_j, w_start, w_rec;
    if (a.t_offset <= w_r0) { w_j = 0; w_start = a.cutoff; w_rec = w_r0; }
    else { w_j = (a.t_offset - w_r0 + S_ - 1) / S_; w_rec = w_r0 + w_j * S_; w_start = w_rec - S_ + 1; }
    bool any_reduce = false;
    for (int q = 0; q < a.n_rec_vars; ++q) any_reduce = any_reduce || (a.rec_reduce[q] != 0);
    const bool readout = a.out_rec != nullptr && a.out_mode == RP_OUT_READOUT;
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
        // 1) source vectors of this CTA's trials (tag t+1) -> shared memory; spins per element until every producer has published
        ll_gather(a.srcbuf + (size_t)(t & 1) * B * Npad + (size_t)b0 * Npad, s_src, Bl, N, Npad, (unsigned int)(t + 1));
        __syncthreads();
        // 2) recurrent drive of the owned rows
        if (Bl <= 4) {
            // one warp per row, lanes stride the columns, all (<= 4) trials in registers
            for (int rr = warp; rr < R; rr += PS_THREADS / 32) {
                const float* wrow = a.w_resident ? s_W + (size_t)rr * a.ldw : a.Wk + (size_t)(r0 + rr) * a.ldw;
                float acc[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) acc[q] = 0.f;
                for (int kv = lane; kv < nvec; kv += 32) {
                    const float4 w4 = a.w_resident ? reinterpret_cast<const float4*>(wrow)[kv] : __ldg(reinterpret_cast<const float4*>(wrow) + kv);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q < Bl) {
                            const float4 s4 = reinterpret_cast<const float4*>(s_src + q * Npad)[kv];
                            acc[q] = fmaf(w4.x, s4.x, acc[q]); acc[q] = fmaf(w4.y, s4.y, acc[q]);
                            acc[q] = fmaf(w4.z, s4.z, acc[q]); acc[q] = fmaf(w4.w, s4.w, acc[q]);
                        }
                    }
                }
                const int kt = (nvec << 2) + lane;              // ragged tail (N % 4 columns)
                if (kt < N) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) if (q < Bl) acc[q] = fmaf(wrow[kt], s_src[q * Npad + kt], acc[q]);
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < Bl) {
                        float tot = acc[q];
                        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                        if (lane == 0) s_u[rr * PS_MAX_B + q] = tot;
                    }
                }
            }
        } else {
            // register tiles of PS_RG rows x PS_BT trials per warp
            const int n_bt = (Bl + PS_BT - 1) / PS_BT, n_units = ((R + PS_RG - 1) / PS_RG) * n_bt;
            for (int unit = warp; unit < n_units; unit += PS_THREADS / 32) {
                const int rbase = (unit / n_bt) * PS_RG, bbase = (unit % n_bt) * PS_BT;
                const float tot = blocked_unit(s_W, a.Wk + (size_t)r0 * a.ldw, a.ldw, a.w_resident != 0, s_src, Npad, N, rbase, R, bbase, Bl, lane);
                const int rr = rbase + lane / PS_BT, q = bbase + lane % PS_BT;
                if (rr < R && q < Bl) s_u[rr * PS_MAX_B + q] = tot;
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
            const float urec = s_u[r * PS_MAX_B + bl];
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
#pragma unroll
                        for (int q = 0; q < RP_MAX_OUT; ++q) if (q < a.k) s_con[q * PS_THREADS + tid] = w_out[q] * yout;
                    }
                }
                if (w.close) {
                    for (int q = 0; q < a.n_rec_vars; ++q) {
                        const int var = a.rec_var[q];
                        const float val = a.rec_post ? (var == 0 ? v1 : (var == 1 ? s1 : x1)) : (var == 0 ? v : (var == 1 ? s : x));
                        if (a.rec_reduce[q]) s_con[(RP_MAX_OUT + q) * PS_THREADS + tid] = val;
                        else a.rec_buf[q][((size_t)w.j * B + b) * N + i] = val;
                    }
                }
            }
            v = v1; s = s1; x = x1;
        }
        // block-level reduction of this step's readout / neuron-mean contributions in a FIXED order (rows ascending), only on steps
        // that produced any (uniform condition); the readout window accumulates in shared memory and leaves the CTA once per record
        if (w.j >= 0 && (readout || (w.close && any_reduce))) {
            __syncthreads();
            if (readout) {
                for (int pr = warp; pr < Bl * a.k; pr += PS_THREADS / 32) {           // one warp per (trial, readout channel)
                    const int bq = pr / a.k, q = pr - bq * a.k;
                    const float tot = warp_ordered_sum(s_con, q * PS_THREADS + bq * R, R, lane);
                    if (lane == 0) {
                        const float acc = w.first ? tot : s_win[pr] + tot;
                        if (w.close) {
                            const float mean = acc / (float)w.len;
                            if (a.out_part) a.out_part[(((size_t)w.j * a.n_rb + rb) * B + b0 + bq) * a.k + q] = mean;
                            else atomicAdd(a.out_rec + ((size_t)w.j * B + b0 + bq) * a.k + q, mean);
                        } else {
                            s_win[pr] = acc;
                        }
                    }
                }
            }
            if (w.close && any_reduce) {
                for (int pr = warp; pr < a.n_rec_vars * Bl; pr += PS_THREADS / 32) {  // one warp per (recorded variable, trial)
                    const int q = pr / Bl, bq = pr - q * Bl;
                    if (!a.rec_reduce[q]) continue;
                    const float tot = warp_ordered_sum(s_con, (RP_MAX_OUT + q) * PS_THREADS + bq * R, R, lane);
                    if (lane == 0) {
                        if (a.rec_part[q]) a.rec_part[q][((size_t)w.j * a.n_rb + rb) * B + b0 + bq] = tot;
                        else atomicAdd(a.rec_buf[q] + (size_t)w.j * B + b0 + bq, tot / (float)N);
                    }
                }
            }
        }
        __syncthreads();          // s_src / s_u / s_con are rewritten by the next step
    }
    if (own) {
        const size_t idx = (size_t)b * N + i;
        a.yT[idx] = v;
        if (NSV > 1) a.yT[plane + idx] = s;
        if (NSV > 2) a.yT[2 * plane + idx] = x;
    }
}

// out[j][e] = scale * sum_rb part[j][rb][e]   (fixed order over the row blocks)
static __global__ void __launch_bounds__(256) k_sum_row_blocks(const float* __restrict__ part, int n_rec, int n_rb, int width, float scale, float* out) {
    const size_t total = (size_t)n_rec * width;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const size_t j = idx / width, e = idx - j * width;
        float acc = 0.f;
        for (int rbk = 0; rbk < n_rb; ++rbk) acc += part[(j * n_rb + rbk) * width + e];
        out[idx] = acc * scale;
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
    int n_rb, BL;                   // grid = n_rb row blocks x ceil(B / BL) trial blocks (as in the forward kernel)
    const float* x; const float* W_in; const float* W_out;
    ModelParams mp;
    const float* history;           // [(T+1)][nsv][B][N]
    const float* g_out_rec;         // [n_rec][B][k] | [n_rec][B][N] or nullptr
    const float* g_yT;              // [nsv][B][N] or nullptr
    uint2* gbuf;                    // [2][B][Npad] {value, tag} pairs, tags zeroed by the host before the launch
    int Npad;
    float* dWrawT;                  // [n_tb][N][ldw]  dWraw^T (row j, col i) of every trial block, zero-initialised when not resident; written at the end
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
    const int rb = blockIdx.x % a.n_rb, tb = blockIdx.x / a.n_rb;
    const int b0 = tb * a.BL;
    const int Bl = max(0, min(a.BL, B - b0));            // trials of this CTA
    float* s_g = psm;                                    // [BL][Npad]  g_t of all neurons, this CTA's trials
    float* s_z = s_g + a.BL * Npad;                      // [rows][PS_MAX_B]
    float* s_src = s_z + PS_MAX_ROWS * PS_MAX_B;         // [rows][PS_MAX_B]  r_t of the owned neurons (for the rank-Bl update)
    float* s_W = s_src + PS_MAX_ROWS * PS_MAX_B;         // [rows][ldw] when resident
    float* s_dW = s_W + (a.w_resident ? (size_t)a.rows_per_cta * a.ldw : 0);   // [rows][ldw] when resident
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int r0 = rb * a.rows_per_cta;
    const int R = max(0, min(a.rows_per_cta, N - r0));
    const size_t plane = (size_t)B * N, slot = (size_t)HistPlanes<MODEL>::N * plane;
    float* dW_slice = a.dWrawT + (size_t)tb * N * a.ldw;  // this trial block's accumulation slice

    if (a.w_resident) for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) s_W[idx] = a.WkT[(size_t)r0 * a.ldw + idx];
    if (a.need_dW && a.dw_resident) for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) s_dW[idx] = 0.f;
    for (int idx = tid; idx < 2 * PS_MAX_ROWS * PS_MAX_B; idx += PS_THREADS) s_z[idx] = 0.f;      // s_z and s_src (unused trial slots stay 0)

    const bool own = tid < R * Bl;
    const int r = own ? tid % R : 0, bl = own ? tid / R : 0;
    const int b = b0 + bl;
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
        ll_gather(a.gbuf + (size_t)(t & 1) * B * Npad + (size_t)b0 * Npad, s_g, Bl, N, Npad, (unsigned int)(a.T - t));
        // source value r_t of the owned neurons (rank-Bl update operand)
        if (own && a.need_dW) {
            float rv;
            if constexpr (SPK) rv = s; else rv = rate_act<MODEL>(a.mp, i, v, b);
            s_src[r * PS_MAX_B + bl] = rv;
        }
        __syncthreads();
        // Z_j = sum_i (kW)^T[j][i] g_i   and   dWraw^T[j][i] += sum_trials r_j * g_i
        if (Bl <= 4) {
            for (int rr = warp; rr < R; rr += PS_THREADS / 32) {
                const float* wrow = a.w_resident ? s_W + (size_t)rr * a.ldw : a.WkT + (size_t)(r0 + rr) * a.ldw;
                float* drow = a.dw_resident ? s_dW + (size_t)rr * a.ldw : dW_slice + (size_t)(r0 + rr) * a.ldw;
                float acc[4], rj[4];
#pragma unroll
                for (int q = 0; q < 4; ++q) { acc[q] = 0.f; rj[q] = (a.need_dW && q < Bl) ? s_src[rr * PS_MAX_B + q] : 0.f; }
                for (int kv = lane; kv < nvec; kv += 32) {
                    const float4 w4 = a.w_resident ? reinterpret_cast<const float4*>(wrow)[kv] : __ldg(reinterpret_cast<const float4*>(wrow) + kv);
                    float4 d4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        if (q < Bl) {
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
                    for (int q = 0; q < 4; ++q) if (q < Bl) { const float gq = s_g[q * Npad + kt]; acc[q] = fmaf(wrow[kt], gq, acc[q]); d = fmaf(rj[q], gq, d); }
                    if (a.need_dW) drow[kt] += d;
                }
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    if (q < Bl) {
                        float tot = acc[q];
                        for (int o = 16; o > 0; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
                        if (lane == 0) s_z[rr * PS_MAX_B + q] = tot;
                    }
                }
            }
        } else {
            // Z: register tiles of PS_RG rows x PS_BT trials per warp
            const int n_bt = (Bl + PS_BT - 1) / PS_BT, n_units = ((R + PS_RG - 1) / PS_RG) * n_bt;
            for (int unit = warp; unit < n_units; unit += PS_THREADS / 32) {
                const int rbase = (unit / n_bt) * PS_RG, bbase = (unit % n_bt) * PS_BT;
                const float tot = blocked_unit(s_W, a.WkT + (size_t)r0 * a.ldw, a.ldw, a.w_resident != 0, s_g, Npad, N, rbase, R, bbase, Bl, lane);
                const int rr = rbase + lane / PS_BT, q = bbase + lane % PS_BT;
                if (rr < R && q < Bl) s_z[rr * PS_MAX_B + q] = tot;
            }
            // dW: every thread owns columns (no reduction), PS_RG rows at a time, the trials in tiles of 4
            if (a.need_dW) {
                const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                for (int rbase = 0; rbase < R; rbase += PS_RG) {
                    for (int kv = tid; kv < nvec; kv += PS_THREADS) {
                        float4 d4[PS_RG];
#pragma unroll
                        for (int rr = 0; rr < PS_RG; ++rr) d4[rr] = zero;
                        for (int qt = 0; qt < Bl; qt += 4) {
                            float4 g4[4];
#pragma unroll
                            for (int q = 0; q < 4; ++q) g4[q] = (qt + q < Bl) ? reinterpret_cast<const float4*>(s_g + (qt + q) * Npad)[kv] : zero;
#pragma unroll
                            for (int rr = 0; rr < PS_RG; ++rr) {
                                const float4 rj = *reinterpret_cast<const float4*>(s_src + (rbase + rr) * PS_MAX_B + qt);     // rows >= R / trials >= Bl hold 0
                                d4[rr].x = fmaf(rj.x, g4[0].x, d4[rr].x); d4[rr].y = fmaf(rj.x, g4[0].y, d4[rr].y); d4[rr].z = fmaf(rj.x, g4[0].z, d4[rr].z); d4[rr].w = fmaf(rj.x, g4[0].w, d4[rr].w);
                                d4[rr].x = fmaf(rj.y, g4[1].x, d4[rr].x); d4[rr].y = fmaf(rj.y, g4[1].y, d4[rr].y); d4[rr].z = fmaf(rj.y, g4[1].z, d4[rr].z); d4[rr].w = fmaf(rj.y, g4[1].w, d4[rr].w);
                                d4[rr].x = fmaf(rj.z, g4[2].x, d4[rr].x); d4[rr].y = fmaf(rj.z, g4[2].y, d4[rr].y); d4[rr].z = fmaf(rj.z, g4[2].z, d4[rr].z); d4[rr].w = fmaf(rj.z, g4[2].w, d4[rr].w);
                                d4[rr].x = fmaf(rj.w, g4[3].x, d4[rr].x); d4[rr].y = fmaf(rj.w, g4[3].y, d4[rr].y); d4[rr].z = fmaf(rj.w, g4[3].z, d4[rr].z); d4[rr].w = fmaf(rj.w, g4[3].w, d4[rr].w);
                            }
                        }
#pragma unroll
                        for (int rr = 0; rr < PS_RG; ++rr) {
                            if (rbase + rr < R) {
                                float* drow = a.dw_resident ? s_dW + (size_t)(rbase + rr) * a.ldw : dW_slice + (size_t)(r0 + rbase + rr) * a.ldw;
                                float4 o = reinterpret_cast<float4*>(drow)[kv];
                                o.x += d4[rr].x; o.y += d4[rr].y; o.z += d4[rr].z; o.w += d4[rr].w;
                                reinterpret_cast<float4*>(drow)[kv] = o;
                            }
                        }
                    }
                    const int kt = (nvec << 2) + tid;               // ragged tail (N % 4 columns)
                    if (kt < N) {
                        for (int rr = 0; rr < PS_RG && rbase + rr < R; ++rr) {
                            float d = 0.f;
                            for (int q = 0; q < Bl; ++q) d = fmaf(s_src[(rbase + rr) * PS_MAX_B + q], s_g[q * Npad + kt], d);
                            float* drow = a.dw_resident ? s_dW + (size_t)(rbase + rr) * a.ldw : dW_slice + (size_t)(r0 + rbase + rr) * a.ldw;
                            drow[kt] += d;
                        }
                    }
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
            const float dI = adj_post_math<MODEL>(aa, rowp, racc, i, b, s_z[r * PS_MAX_B + bl], v, s, x, av, as, ax, urec);
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
        for (int idx = tid; idx < R * a.ldw; idx += PS_THREADS) dW_slice[(size_t)r0 * a.ldw + idx] = s_dW[idx];
    }
}

// dWraw^T -> dW = diag(k) dWraw and dk  (persistent path keeps the weight gradient transposed: row = source neuron j)
static __global__ void __launch_bounds__(256) k_finish_wgrad_T(int N, const float* __restrict__ dWrawT, int ldr, const float* __restrict__ W,
                                                        const float* __restrict__ kp, int k_stride, float* dW, float* dk, int n_slices) {
    __shared__ float tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx: i base, by: j base (rows of dWrawT)
    const int tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
    for (int rr = ty; rr < 32; rr += 8) {
        const int j = by + rr, i = bx + tx;
        float acc = 0.f;
        if (i < N && j < N) for (int z = 0; z < n_slices; ++z) acc += dWrawT[(size_t)z * N * ldr + (size_t)j * ldr + i];     // trial blocks, fixed order
        tile[rr][tx] = acc;
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
