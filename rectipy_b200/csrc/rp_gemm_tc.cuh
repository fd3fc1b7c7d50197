// tcgen05 (5th-gen tensor core) contraction with error-compensated split accumulation, sm_100a only.
//
//   C[q*ldc + p] (+)= sum_k A[p][k] * B[q][k]        A, B K-major, each pre-split into (hi, lo) with 11-bit significands
//   A.B ~= A_hi.B_hi + A_lo.B_hi + A_hi.B_lo         three MMAs per logical product, fp32 accumulate in TMEM
//
// Two element formats carry the split (template parameter F16):
//   tf32 (kind::tf32): hi, lo stored as fp32 words, fp32 exponent range, K = 8 per MMA
//   fp16 (kind::f16) : hi, lo stored as binary16 -- the SAME 11-bit significand as tf32, so hi + lo represents x to
//                      2^-22 |x| exactly like the tf32 split, but the tensor core issues it at twice the rate (K = 16 per
//                      MMA) and the operands take half the bytes.  The 5-bit exponent is handled by exact power-of-two
//                      operand scales derived on the device from tracked maxima (ScaleRef, rp_kernels.cuh); the epilogue
//                      undoes them.  The product of two binary16 numbers is exact in fp32, so nothing else changes.
//
// This is what replaces `weights @ x` (edges.py:49 / the generated field, nodes.py:169) and autograd's W^T.grad and
// grad (x) r products when many trials are batched.  1e-5 parity (BASELINE.json) rules out single-pass TF32 / FP16.
//
// Kernel shape: one CTA per 128(p) x BQ(q) tile, 6 warps:
//   warp 0   TMA producer   (cp.async.bulk.tensor 2D; K blocks of 16 fp32 = 64-byte swizzled rows, 4-stage ring at BQ=256.
//                            Measured: 2 stages of 32-wide blocks left the tensor pipe 75 % busy, 4 stages of 16 -> 124 vs 148 us)
//   warp 1   TMEM allocator + single-thread tcgen05.mma issuer (6 MMAs of K=8 per 16-wide K block: lo.hi, hi.lo, hi.hi)
//   warps 2-9 epilogue      (tcgen05.ld 32x32b -> fp32 register accumulators -> coalesced global stores, p fastest)
// smem ring of full/empty mbarriers between producer and issuer; two TMEM accumulator buffers ping-pong between the
// issuer and the epilogue (tcgen05.commit signals "chunk done", the epilogue's mbarrier.arrive signals "buffer drained").
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <cudaTypedefs.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <type_traits>
#include "rp_kernels.cuh"

namespace rp {

inline char* tc_err_buf() { static thread_local char buf[384] = ""; return buf; }
inline const char* tc_last_error() { return tc_err_buf(); }
#define RP_TC_FAIL(...) do { snprintf(rp::tc_err_buf(), 384, __VA_ARGS__); return 1; } while (0)

constexpr int TC_BP = 128;          // tile rows (UMMA M, TMEM lanes)
#ifndef RP_TC_ROW_BYTES
#define RP_TC_ROW_BYTES 64
#endif
constexpr int TC_ROW_BYTES = RP_TC_ROW_BYTES;   // bytes of one operand row per K block = one swizzle row
static_assert(TC_ROW_BYTES == 128 || TC_ROW_BYTES == 64, "K block must be one 128B or one 64B swizzle row");
#ifndef RP_TC_CHUNK_K
#define RP_TC_CHUNK_K 128
#endif
#ifndef RP_TC_CHUNK_K_F16
#define RP_TC_CHUNK_K_F16 128
#endif
// element format of the split operands
template <bool F16> struct TcElt {
    static constexpr int ESIZE = F16 ? 2 : 4;
    static constexpr int BK = TC_ROW_BYTES / ESIZE;                            // elements per K block (64 B rows: 16 tf32 / 32 fp16)
    static constexpr int KSTEPS = TC_ROW_BYTES / 32;                           // MMAs of 32 operand bytes (K = 8 tf32 / 16 fp16) per K block
    static constexpr int KC = (F16 ? RP_TC_CHUNK_K_F16 : RP_TC_CHUNK_K) / BK;  // K blocks per TMEM accumulation chunk
    static_assert(KC >= 1, "accumulation chunk shorter than one K block");
};

enum { TC_FWD = 0, TC_DGRAD = 1, TC_WGRAD = 2 };
constexpr int TC_WGRAD_SPLITS = 2;

// ---------------------------------------------------------------------------------------------------------
// device helpers (inline PTX)
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "WAIT_LOOP:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra WAIT_LOOP;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, uint64_t* bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void tma_prefetch_desc(const CUtensorMap* map) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(map)) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n"
        ".reg .b32 rx;\n"
        ".reg .pred px;\n"
        "elect.sync rx|px, %1;\n"
        "@px mov.s32 %0, 1;\n"
        "}\n" : "+r"(pred) : "r"(0xffffffffu));
    return pred != 0;
}
// K-major, SWIZZLE_128B shared-memory matrix descriptor (cute::UMMA::SmemDescriptor bit layout):
//   [0,14) start>>4 | [16,30) LBO>>4 (unused for swizzled K-major) | [32,46) SBO>>4 = 1024>>4 | [46,48) version=1 | [61,64) layout=2
__device__ __forceinline__ uint64_t make_sw128_kmajor_desc(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);
    d |= (uint64_t)((8u * TC_ROW_BYTES) >> 4) << 32;        // SBO: 8 rows of one swizzle row each
    d |= (uint64_t)1 << 46;
    d |= (uint64_t)(TC_ROW_BYTES == 128 ? 2 : 4) << 61;     // SWIZZLE_128B = 2, SWIZZLE_64B = 4
    return d;
}
// instruction descriptor (cute::UMMA::InstrDescriptor): c_format F32 [4,6)=1, a/b format [7,10)/[10,13) (kind::tf32: TF32 = 2,
// kind::f16: F16 = 0), a/b K-major, N>>3 at [17,23), M>>4 at [24,29)
__host__ __device__ constexpr uint32_t make_idesc(bool f16, int M, int N) {
    return (1u << 4) | ((f16 ? 0u : 2u) << 7) | ((f16 ? 0u : 2u) << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
template <bool F16>
__device__ __forceinline__ void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    if constexpr (F16) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    } else {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "setp.ne.b32 p, %4, 0;\n"
            "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n"
            "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate) : "memory");
    }
}
__device__ __forceinline__ void umma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr));
}

__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr));
}

// The tensor core adds into its fp32 accumulator with truncation (measured on B200: all-positive operands give a
// systematic -2.3e-5 relative bias at K=2048, growing linearly with K).  To stay at fp32 accuracy the K loop is cut into
// chunks of TC_KC K-blocks; each chunk accumulates into one of two TMEM buffers from zero and the epilogue warps add the
// finished chunk into fp32 registers with round-to-nearest while the next chunk is being multiplied.
constexpr int TC_EPI_WARPS = 8;     // two warps per TMEM lane quadrant, each owning half of the tile's columns
constexpr int TC_THREADS = 64 + 32 * TC_EPI_WARPS;

template <int BQ> struct TcCfg {
    static constexpr int A_BYTES = TC_BP * TC_ROW_BYTES;
    static constexpr int B_BYTES = BQ * TC_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * B_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;      // 128 B rows: 2 (BQ=256) / 3 ; 64 B rows: 4 / 6
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 /*align slack*/ + 256 /*barriers*/;
    static constexpr int TMEM_COLS = 2 * BQ;
    static constexpr int COLS_PER_THREAD = BQ / 2;
};

__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}

// ---- epilogues -----------------------------------------------------------------------------------------------------------
// While the K loop runs, epilogue thread (warp, lane) owns TMEM lane = tile row and CPT columns (that is how tcgen05.ld hands
// out the accumulator).  Fused epilogues then park the fp32 tile in the idle shared-memory stages as tile[col][row], meet at a
// named barrier and RE-PARTITION the work: thread (w, l) takes the 4 consecutive neurons 4l..4l+3 and a contiguous block of
// BQ/8 trials.  Every global access of the element-wise step is then a 16-byte vector over neurons (512 B per warp and trial),
// loads of a batch are issued before any store, and the weight-gradient operands are written as 16-byte vectors over trials.
struct EpiStore {                       // C[q*ldc + p] (+)= acc ; split-K slices (blockIdx.z) go to C + z*split_stride
    static constexpr bool kStage = false;
    static constexpr unsigned kTag = TR_GEMM_STORE;
    float* C; int ldc; int accumulate; size_t split_stride;
    ScaleRef sa, sb;                    // binary16 operands: scales of the A and B operand (undone here)
    __device__ __forceinline__ float2 unscale(int /*p0*/) const { return make_float2(exp2i(-scale_expo(sa)), exp2i(-scale_expo(sb))); }
};

__device__ __forceinline__ float4 ldg4(const float* p) { return __ldg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ void st4(float* p, float a, float b, float c, float d) { *reinterpret_cast<float4*>(p) = make_float4(a, b, c, d); }
__device__ __forceinline__ float f4get(const float4& v, int r) { return r == 0 ? v.x : (r == 1 ? v.y : (r == 2 ? v.z : v.w)); }

// forward step: u = (kW . r_t)[b][i] never touches global memory; tile rows >= N hold W_out (readout o_t = W_out . s_t)
// GEN: general element path (per-element parameter loads: ik_op, per-trial parameter sweeps) instead of hoisted row constants
template <int MODEL, bool GEN = is_ik(MODEL)>
struct EpiFwd {
    static constexpr bool kStage = true;
    static constexpr unsigned kTag = TR_GEMM_FWD;
    FwdStepArgs a;
    float* out_rec_j; int k; int win_first, win_close; float inv_len;
    ScaleRef sA, sA_ro, sB;             // binary16 operands: scales of kW, of the appended W_out rows and of src_t
    __device__ __forceinline__ float2 unscale(int p0) const {
        return make_float2(exp2i(-scale_expo(p0 < a.N ? sA : sA_ro)), exp2i(-scale_expo(sB)));
    }

    template <int BQ, bool F16>
    __device__ __forceinline__ void run_tile(int p0, int q0, const float* tile, int et, float* extra) const {
        constexpr int NSV = ModelTraits<MODEL>::NSV;
        constexpr int NCOL = BQ / 8;
        const int w = et >> 5, l = et & 31;
        const int cbase = w * NCOL;
        const size_t plane = (size_t)a.B * a.N;
        if (p0 < a.N) {
            const int i0 = p0 + 4 * l;
            FwdRow row[4];
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) row[rr] = fwd_row<MODEL>(a, i0 + rr);
            const float so = F16 ? exp2i(scale_expo(a.sc_out)) : 1.f;
            float smax = 0.f;
            // Round 2, lean loop for the headline templates (qif_op / lif_op, projected input of width <= 2, binary16 operands).
            // ncu source page of the round-1 loop: 587 instructions per batch of 16 elements, of which ~240 integer / control
            // (64-bit index arithmetic per access, the per-element in_mode / channel-loop branches, constant reloads) and IPC per
            // warp ~0.15 -- with 8 warps per SM the loop is bound by the length of each warp's instruction stream, not by memory
            // (a double-buffered cp.async pipeline for the state loads was measured: 31.35 vs 31.34 ms per pass, no gain).  Here:
            // running pointers, input mode resolved once, packed binary16 conversions; the arithmetic is expression for expression that of
            // fwd_elem_fast (a cheaper s' = fma(s, 1 - dt/tau_s, spike) moved a lif spike by a step in the parity tests).
            if constexpr (!GEN && F16 && (MODEL == RP_QIF || MODEL == RP_LIF)) {
                if (!a.no_lean && a.in_target == 0 && (a.in_mode == RP_IN_NONE || (a.in_mode == RP_IN_PROJ && a.m <= 2))) {
                    constexpr int NB = 4;
                    const int N = a.N;
                    const bool proj = a.in_mode == RP_IN_PROJ;
                    const int xm = proj ? a.m : 0;
                    const float dt = a.dt, theta = a.theta, v_reset = a.v_reset;
                    float its[4], eta[4], itau[4], wi0[4], wi1[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        its[rr] = row[rr].inv_tau_s; eta[rr] = row[rr].eta; itau[rr] = row[rr].inv_tau;
                        wi0[rr] = row[rr].wi0; wi1[rr] = (proj && a.m > 1) ? row[rr].wi1 : 0.f;
                    }
                    const size_t o0 = (size_t)(q0 + cbase) * N + i0;
                    const float* yv = a.y_cur + o0;
                    const float* ys = yv + plane;
                    float* nv = a.y_next + o0;
                    float* ns = nv + plane;
                    __half* ph = reinterpret_cast<__half*>(a.src_hi) + (size_t)(q0 + cbase) * a.ld_src + i0;
                    __half* pl = reinterpret_cast<__half*>(a.src_lo) + (size_t)(q0 + cbase) * a.ld_src + i0;
                    const float* xp = proj ? a.x_t + (size_t)(q0 + cbase) * xm : nullptr;
                    const float* ut = tile + (size_t)cbase * TC_BP + 4 * l;
                    const int x1off = xm > 1 ? 1 : 0;
                    for (int c = 0; c < NCOL; c += NB) {
                        float4 u4[NB], v4[NB], s4[NB];
                        float x0[NB], x1[NB];
#pragma unroll
                        for (int cc = 0; cc < NB; ++cc) {
                            u4[cc] = *reinterpret_cast<const float4*>(ut + cc * TC_BP);
                            v4[cc] = ldg4(yv + (size_t)cc * N);
                            s4[cc] = ldg4(ys + (size_t)cc * N);
                            x0[cc] = proj ? __ldg(xp + cc * xm) : 0.f;
                            x1[cc] = proj ? __ldg(xp + cc * xm + x1off) : 0.f;
                        }
#pragma unroll
                        for (int cc = 0; cc < NB; ++cc) {
                            float v1[4], s1[4];
#pragma unroll
                            for (int rr = 0; rr < 4; ++rr) {
                                const float v = f4get(v4[cc], rr), sv = f4get(s4[cc], rr), u = f4get(u4[cc], rr);
                                const float Iin = fmaf(wi1[rr], x1[cc], wi0[rr] * x0[cc]);
                                const bool spike = v >= theta;                                   // heaviside(v - theta, 1.0)  nodes.py:383,476
                                float vt;
                                if constexpr (MODEL == RP_LIF) vt = fmaf(dt, fmaf(-v, itau[rr], u) + Iin + eta[rr], v);      // lif.yaml:10-15
                                else vt = fmaf(dt, fmaf(fmaf(v, v, eta[rr]) + Iin, itau[rr], u), v);                           // qif.yaml:10-12
                                s1[rr] = sv + dt * (-sv * its[rr]) + (spike ? 1.0f : 0.0f);      // same expression as every other path; dt * (spike / dt) == 1  nodes.py:385
                                v1[rr] = spike ? v_reset : vt;                                   // reset blend  nodes.py:390
                                smax = fmaxf(smax, fabsf(s1[rr]));
                            }
                            st4(nv + (size_t)cc * N, v1[0], v1[1], v1[2], v1[3]);
                            st4(ns + (size_t)cc * N, s1[0], s1[1], s1[2], s1[3]);
                            // binary16 split of the next source operand, two elements per conversion
                            const __half2 h01 = __floats2half2_rn(s1[0] * so, s1[1] * so), h23 = __floats2half2_rn(s1[2] * so, s1[3] * so);
                            const float2 f01 = __half22float2(h01), f23 = __half22float2(h23);
                            const __half2 l01 = __floats2half2_rn(fmaf(s1[0], so, -f01.x), fmaf(s1[1], so, -f01.y));
                            const __half2 l23 = __floats2half2_rn(fmaf(s1[2], so, -f23.x), fmaf(s1[3], so, -f23.y));
                            uint2 wh, wl;
                            wh.x = *reinterpret_cast<const uint32_t*>(&h01); wh.y = *reinterpret_cast<const uint32_t*>(&h23);
                            wl.x = *reinterpret_cast<const uint32_t*>(&l01); wl.y = *reinterpret_cast<const uint32_t*>(&l23);
                            *reinterpret_cast<uint2*>(ph + (size_t)cc * a.ld_src) = wh;
                            *reinterpret_cast<uint2*>(pl + (size_t)cc * a.ld_src) = wl;
                        }
                        yv += (size_t)NB * N; ys += (size_t)NB * N; nv += (size_t)NB * N; ns += (size_t)NB * N;
                        ph += (size_t)NB * a.ld_src; pl += (size_t)NB * a.ld_src; ut += NB * TC_BP;
                        if (proj) xp += NB * xm;
                    }
                    if (a.amax_out) {
                        smax = warp_max(smax);
                        if (l == 0 && smax > 0.f) atomic_max_nonneg(a.amax_out, smax);
                    }
                    return;
                }
            }
            // NB trials per batch: all their loads are issued before the first store (8 warps per SM have to cover the HBM
            // latency alone here: measured epilogue 35 us at 2 trials per batch, 25 us at 4, 43 us at 8; issuing batch k+1 before
            // batch k is computed -- 2+2 or 4+4 trials in flight -- measured 27 and 31 us: no better than plain batches of 4)
            constexpr int NB = 4;
            static_assert(NCOL % NB == 0, "trial batch must divide the per-warp column block");
            for (int c = 0; c < NCOL; c += NB) {
                float4 u4[NB], v4[NB], s4[NB], x4[NB], xd4[NB];
                float4 w4[NSV > 3 ? NB : 1];       // fourth state plane (ik_biexp_op)
                float xin0[NB], xin1[NB];
#pragma unroll
                for (int cc = 0; cc < NB; ++cc) {
                    const int b = q0 + cbase + c + cc;
                    const size_t idx = (size_t)b * a.N + i0;
                    const float4 zero = make_float4(0.f, 0.f, 0.f, 0.f);
                    u4[cc] = *reinterpret_cast<const float4*>(tile + (size_t)(cbase + c + cc) * TC_BP + 4 * l);
                    v4[cc] = ldg4(a.y_cur + idx);
                    s4[cc] = NSV > 1 ? ldg4(a.y_cur + plane + idx) : zero;
                    x4[cc] = NSV > 2 ? ldg4(a.y_cur + 2 * plane + idx) : zero;
                    if constexpr (NSV > 3) w4[cc] = ldg4(a.y_cur + 3 * plane + idx);
                    xd4[cc] = a.in_mode == RP_IN_DENSE ? ldg4(a.x_t + idx) : zero;
                    xin0[cc] = a.in_mode == RP_IN_PROJ ? __ldg(a.x_t + (size_t)b * a.m) : 0.f;
                    xin1[cc] = (a.in_mode == RP_IN_PROJ && a.m > 1) ? __ldg(a.x_t + (size_t)b * a.m + 1) : 0.f;
                }
#pragma unroll
                for (int cc = 0; cc < NB; ++cc) {
                    const int b = q0 + cbase + c + cc;
                    const size_t idx = (size_t)b * a.N + i0;
                    float v1[4], s1[4], x1[4], w1[4], hi[4], lo[4], sr[4];
#pragma unroll
                    for (int rr = 0; rr < 4; ++rr) {
                        const int i = i0 + rr;
                        w1[rr] = 0.f;
                        if constexpr (GEN) {                  // general (per-element parameter loads) path: ik, parameter sweeps
                            const float Iin = input_current(a.in_mode, a.m, a.x_t, a.W_in, a.N, b, i);
                            const float wv = NSV > 3 ? f4get(w4[NSV > 3 ? cc : 0], rr) : 0.f;
                            fwd_elem<MODEL>(a, i, f4get(u4[cc], rr), Iin, f4get(v4[cc], rr), f4get(s4[cc], rr), f4get(x4[cc], rr), v1[rr], s1[rr], x1[rr], b,
                                            wv, &w1[rr]);
                        } else {
                            fwd_elem_fast<MODEL>(a, row[rr], i, b, f4get(u4[cc], rr), xin0[cc], xin1[cc], f4get(xd4[cc], rr),
                                                 f4get(v4[cc], rr), f4get(s4[cc], rr), f4get(x4[cc], rr), v1[rr], s1[rr], x1[rr]);
                        }
                        float src1;
                        if constexpr (ModelTraits<MODEL>::SPIKING) src1 = s1[rr]; else src1 = rate_act<MODEL>(a.mp, i, v1[rr], b);
                        sr[rr] = src1;
                        if constexpr (F16) smax = fmaxf(smax, fabsf(src1)); else split_tf32(src1, hi[rr], lo[rr]);
                    }
                    if (is_ik(MODEL) && a.urec_out) *reinterpret_cast<float4*>(a.urec_out + idx) = u4[cc];
                    st4(a.y_next + idx, v1[0], v1[1], v1[2], v1[3]);
                    if (NSV > 1) st4(a.y_next + plane + idx, s1[0], s1[1], s1[2], s1[3]);
                    if (NSV > 2) st4(a.y_next + 2 * plane + idx, x1[0], x1[1], x1[2], x1[3]);
                    if constexpr (NSV > 3) st4(a.y_next + 3 * plane + idx, w1[0], w1[1], w1[2], w1[3]);
                    if constexpr (F16) {
                        store_split4_f16(a.src_hi, a.src_lo, (size_t)b * a.ld_src + i0, sr, so);
                    } else {
                        st4(reinterpret_cast<float*>(a.src_hi) + (size_t)b * a.ld_src + i0, hi[0], hi[1], hi[2], hi[3]);
                        st4(reinterpret_cast<float*>(a.src_lo) + (size_t)b * a.ld_src + i0, lo[0], lo[1], lo[2], lo[3]);
                    }
                }
            }
            if constexpr (F16) {
                if (a.amax_out) {
                    smax = warp_max(smax);
                    if (l == 0 && smax > 0.f) atomic_max_nonneg(a.amax_out, smax);
                }
            }
        } else if (out_rec_j != nullptr && p0 == a.N) {      // (a padding tile further out -- CTA-pair grid -- has nothing to do)
            // readout rows: row kk of this tile is W_out[kk]; o_t[b][kk] accumulates into the open record window
#pragma unroll
            for (int rr = 0; rr < 4; ++rr) {
                const int kk = 4 * l + rr;
                if (kk < k) {
                    for (int c = 0; c < NCOL; ++c) {
                        float* dst = out_rec_j + (size_t)(q0 + cbase + c) * k + kk;
                        float val = tile[(size_t)(cbase + c) * TC_BP + kk];
                        if (!win_first) val += *dst;
                        if (win_close) val *= inv_len;
                        *dst = val;
                    }
                }
            }
        }
    }
};

// A launch may cover only the work items [lin0, lin0 + gridDim.x) of the full (gx, gy, gz) grid, x fastest (gx == 0: the
// launch grid is the full grid).  rp_backward uses this to spread one weight-gradient contraction over several reverse steps.
struct TcSlice { int lin0, gx, gy; };

// LEAN: register-lean variant (16-column TMEM drains, 16-deep read-modify-write batches) compiled to at most TC_LEAN_REGS
// registers.  An SM's register file is split over its four sub-partitions (16 K registers each) and the 10 warps of this
// kernel land 3/3/2/2 on them: at 168 registers the two full sub-partitions have 256 registers left and no other block can
// become resident; at 144 every sub-partition keeps >= 2560 free = one 80-register warp, so one 4-warp block of the HBM-bound
// adjoint kernels runs beside each weight-gradient CTA (rp_backward overlaps them on purpose).
constexpr int TC_LEAN_REGS = 144;

template <int BQ, class Epi, bool F16, bool LEAN>
__device__ __forceinline__ void gemm_split3_body(const CUtensorMap& tmA_hi, const CUtensorMap& tmA_lo, const CUtensorMap& tmB_hi, const CUtensorMap& tmB_lo,
                                                 int num_k_blocks, const Epi& epi, const TcSlice slice) {
    using Cfg = TcCfg<BQ>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CPT = Cfg::COLS_PER_THREAD;
    constexpr int TC_BK = TcElt<F16>::BK;
    constexpr int TC_KC = TcElt<F16>::KC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    if (slice.gx > 0) {
        const int lin = slice.lin0 + (int)blockIdx.x;
        bx = lin % slice.gx; by = (lin / slice.gx) % slice.gy; bz = lin / (slice.gx * slice.gy);
    }
    const int p0 = bx * TC_BP, q0 = by * BQ;
    const int num_chunks = (num_k_blocks + TC_KC - 1) / TC_KC;
    const int kb0 = bz * num_k_blocks;                      // split-K: this CTA contracts K blocks [kb0, kb0 + num_k_blocks)

    TraceRec** trace_slot = reinterpret_cast<TraceRec**>(tmem_slot + 2);     // shared: every warp stamps its own end time
    if (warp == 0 && lane == 0) {
        *trace_slot = trace_begin(slice.gx > 0 ? (unsigned)TR_GEMM_WGRAD : Epi::kTag);
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;
    // everything above (barriers, tensor-map prefetch, TMEM allocation) overlapped the tail of the previous kernel of the chain;
    // operands, scales and state are its products
    pdl_launch_dependents();
    pdl_wait();

    if (warp == 0) {
        // ===== TMA producer =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                tma_load_2d(sa, &tmA_hi, &full_bar[stage], (kb0 + kb) * TC_BK, p0);
                tma_load_2d(sa + Cfg::A_BYTES, &tmA_lo, &full_bar[stage], (kb0 + kb) * TC_BK, p0);
                tma_load_2d(sa + 2 * Cfg::A_BYTES, &tmB_hi, &full_bar[stage], (kb0 + kb) * TC_BK, q0);
                tma_load_2d(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES, &tmB_lo, &full_bar[stage], (kb0 + kb) * TC_BK, q0);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one elected thread) =====
        constexpr uint32_t idesc = make_idesc(F16, TC_BP, BQ);
        int stage = 0; uint32_t phase = 0;
        for (int kb = 0; kb < num_k_blocks; ++kb) {
            const int chunk = kb / TC_KC, kin = kb - chunk * TC_KC;
            const int buf = chunk & 1;
            if (kin == 0) {
                // the epilogue must have drained this buffer (chunk-2) before it is overwritten
                mbar_wait(&tmem_empty_bar[buf], ((chunk >> 1) & 1) ^ 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            }
            mbar_wait(&full_bar[stage], phase);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            if (elect_one()) {
                const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BQ);
                const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                const uint64_t dA_hi = make_sw128_kmajor_desc(sa);
                const uint64_t dA_lo = make_sw128_kmajor_desc(sa + Cfg::A_BYTES);
                const uint64_t dB_hi = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES);
                const uint64_t dB_lo = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                for (int k = 0; k < TcElt<F16>::KSTEPS; ++k) {
                    const uint64_t adv = (uint64_t)((k * 32) >> 4);                // 32 operand bytes per MMA, encoded >>4
                    // small cross terms first, leading term last
                    umma_ss<F16>(tmem_d, dA_lo + adv, dB_hi + adv, idesc, (kin > 0 || k > 0) ? 1u : 0u);
                    umma_ss<F16>(tmem_d, dA_hi + adv, dB_lo + adv, idesc, 1u);
                    umma_ss<F16>(tmem_d, dA_hi + adv, dB_hi + adv, idesc, 1u);
                }
                umma_commit(&empty_bar[stage]);                       // frees the smem stage when these MMAs retire
                if (kin == TC_KC - 1 || kb == num_k_blocks - 1) umma_commit(&tmem_full_bar[buf]);
            }
            __syncwarp();
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
    } else {
        // ===== epilogue: warps 2..9; warp w owns TMEM lanes 32*(w%4).., columns [half*CPT, (half+1)*CPT) =====
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;
        const int half = ew >> 2;
        const int p = p0 + lane_base + lane;
        float acc[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
        for (int chunk = 0; chunk < num_chunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(&tmem_full_bar[buf], (chunk >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * BQ + half * CPT);
            if constexpr (LEAN) {
#pragma unroll
                for (int c = 0; c < CPT / 16; ++c) {
                    uint32_t r[16];
                    tmem_ld16(taddr + (uint32_t)(c * 16), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 16; ++j) acc[c * 16 + j] += __uint_as_float(r[j]);
                }
            } else {
#pragma unroll
                for (int c = 0; c < CPT / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                }
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
        }
        float2 usc = make_float2(1.f, 1.f);
        if constexpr (F16) usc = epi.unscale(p0);
        if constexpr (!Epi::kStage) {
            float* cbase = epi.C + bz * epi.split_stride + (size_t)(q0 + half * CPT) * epi.ldc + p;
            if (epi.accumulate) {
                // read-modify-write of the accumulation slice: RB independent loads in flight before the first store
                constexpr int RB = LEAN ? 8 : 32;
#pragma unroll
                for (int j0 = 0; j0 < CPT; j0 += RB) {
                    float old[RB];
#pragma unroll
                    for (int j = 0; j < RB; ++j) old[j] = __ldcg(cbase + (size_t)(j0 + j) * epi.ldc);
#pragma unroll
                    for (int j = 0; j < RB; ++j) {
                        float v = acc[j0 + j];
                        if constexpr (F16) v = v * usc.x * usc.y;
                        cbase[(size_t)(j0 + j) * epi.ldc] = v + old[j];
                    }
                }
            } else {
#pragma unroll
                for (int j = 0; j < CPT; ++j) {
                    float v = acc[j];
                    if constexpr (F16) v = v * usc.x * usc.y;
                    cbase[(size_t)j * epi.ldc] = v;
                }
            }
        } else {
            // all MMAs have retired (last tmem_full barrier) and every TMA load was consumed: the stages are free
            float* tile = reinterpret_cast<float*>(smem);                      // tile[col][row], BQ x 128 fp32
            float* my = tile + (size_t)(half * CPT) * TC_BP + lane_base + lane;
#pragma unroll
            for (int j = 0; j < CPT; ++j) my[j * TC_BP] = F16 ? acc[j] * usc.x * usc.y : acc[j];
            asm volatile("bar.sync 1, 256;" ::: "memory");                      // the 8 epilogue warps only
            epi.template run_tile<BQ, F16>(p0, q0, tile, ew * 32 + lane, tile + (size_t)BQ * TC_BP);
        }
    }
    if (lane == 0 && *trace_slot != nullptr) atomicMax(&(*trace_slot)->t1, trace_now());
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

template <int BQ, class Epi, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_split3(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
              const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
              int num_k_blocks, const __grid_constant__ Epi epi, const TcSlice slice) {
    gemm_split3_body<BQ, Epi, F16, false>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, num_k_blocks, epi, slice);
}
// the co-residency variant (see TC_LEAN_REGS)
template <int BQ, class Epi, bool F16>
__global__ void __maxnreg__(TC_LEAN_REGS)
k_gemm_split3_lean(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,
                   int num_k_blocks, const __grid_constant__ Epi epi, const TcSlice slice) {
    gemm_split3_body<BQ, Epi, F16, true>(tmA_hi, tmA_lo, tmB_hi, tmB_lo, num_k_blocks, epi, slice);
}

// ---------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant (tcgen05 cta_group::2) of the plain contraction (adjoint product, weight gradient), round 2.
// Two CTAs of a cluster (neighbouring row tiles, the SAME 256 columns) run one M = 256 MMA: each CTA keeps its own 128 rows of A and
// its 128 x 256 accumulator, and only HALF of the B tile -- the tensor cores of the pair read both halves.  Per CTA and K block the
// shared-memory fill drops from 48 KB to 32 KB and every MMA reads 8 KB instead of 12 KB: the contraction is power-bound under the
// 1000 W cap, so bytes moved per flop is what is left to save.  Protocol (as in CUTLASS' 2-SM kernels): both producers load with
// cp.async.bulk.tensor...cta_group::2 onto the LEADER's full barrier (address with the peer bit cleared), the leader's elected
// thread issues tcgen05.mma.cta_group::2 and commits with a cluster multicast onto both CTAs' empty / accumulator-ready barriers,
// the epilogue warps of both CTAs arrive on the leader's accumulator-free barrier.
// ---------------------------------------------------------------------------------------------------------------------------
constexpr uint32_t TC_PEER_MASK = 0xFEFFFFFFu;      // shared::cluster address of the same offset in the even CTA of the pair
struct TcCfg2 {
    static constexpr int BQ = 256;
    static constexpr int A_BYTES = TC_BP * TC_ROW_BYTES;
    static constexpr int BH_BYTES = (BQ / 2) * TC_ROW_BYTES;
    static constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * BH_BYTES;
    static constexpr int STAGES = (192 * 1024) / STAGE_BYTES;
    static constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 1024 + 256;
    static constexpr int TMEM_COLS = 2 * BQ;
    static constexpr int COLS_PER_THREAD = BQ / 2;
};
__device__ __forceinline__ void tma_load_2d_cg2(void* dst, const CUtensorMap* map, uint64_t* leader_bar, int c0, int c1) {
    asm volatile(
        "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_u32(dst)), "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(leader_bar) & TC_PEER_MASK), "r"(c0), "r"(c1) : "memory");
}
__device__ __forceinline__ void umma_ss_cg2(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, {%5, %5, %5, %5, %5, %5, %5, %5}, p;\n"
        "}\n" ::"r"(tmem_d), "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate), "r"(0u) : "memory");
}
__device__ __forceinline__ void umma_commit_cg2_mc(uint64_t* bar) {     // arrives on the barrier at this offset in BOTH CTAs
    asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;"
                 ::"r"(smem_u32(bar)), "h"((uint16_t)3) : "memory");
}
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & TC_PEER_MASK) : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\nbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}

__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_split3_cg2(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                  const __grid_constant__ CUtensorMap tmB_hi, const __grid_constant__ CUtensorMap tmB_lo,      // boxes of 128 rows
                  int num_k_blocks, const __grid_constant__ EpiStore epi) {
    using Cfg = TcCfg2;
    constexpr int BQ = Cfg::BQ;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CPT = Cfg::COLS_PER_THREAD;
    constexpr int TC_BK = TcElt<true>::BK;
    constexpr int TC_KC = TcElt<true>::KC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);      // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] used in the leader only (both CTAs' epilogue warps arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int bx = blockIdx.x, by = blockIdx.y, bz = blockIdx.z;
    const int p0 = bx * TC_BP, q0 = by * BQ;
    const int qh = q0 + (int)rank * (BQ / 2);               // this CTA's half of the B tile
    const int num_chunks = (num_k_blocks + TC_KC - 1) / TC_KC;
    const int kb0 = bz * num_k_blocks;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB_hi); tma_prefetch_desc(&tmB_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * TC_EPI_WARPS); }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                      // the peer's barriers exist before anything arrives on them
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs): own rows of A, own half of B, transaction bytes onto the leader's barrier =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                mbar_wait(&empty_bar[stage], phase ^ 1);
                uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                tma_load_2d_cg2(sa, &tmA_hi, &full_bar[stage], (kb0 + kb) * TC_BK, p0);
                tma_load_2d_cg2(sa + Cfg::A_BYTES, &tmA_lo, &full_bar[stage], (kb0 + kb) * TC_BK, p0);
                tma_load_2d_cg2(sa + 2 * Cfg::A_BYTES, &tmB_hi, &full_bar[stage], (kb0 + kb) * TC_BK, qh);
                tma_load_2d_cg2(sa + 2 * Cfg::A_BYTES + Cfg::BH_BYTES, &tmB_lo, &full_bar[stage], (kb0 + kb) * TC_BK, qh);
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread of the LEADER CTA drives both tensor cores =====
        if (leader) {
            constexpr uint32_t idesc = make_idesc(true, 2 * TC_BP, BQ);
            int stage = 0; uint32_t phase = 0;
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int chunk = kb / TC_KC, kin = kb - chunk * TC_KC;
                const int buf = chunk & 1;
                if (kin == 0) {
                    mbar_wait(&tmem_empty_bar[buf], ((chunk >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&full_bar[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BQ);
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t dA_hi = make_sw128_kmajor_desc(sa);
                    const uint64_t dA_lo = make_sw128_kmajor_desc(sa + Cfg::A_BYTES);
                    const uint64_t dB_hi = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES);
                    const uint64_t dB_lo = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES + Cfg::BH_BYTES);
#pragma unroll
                    for (int k = 0; k < TcElt<true>::KSTEPS; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        umma_ss_cg2(tmem_d, dA_lo + adv, dB_hi + adv, idesc, (kin > 0 || k > 0) ? 1u : 0u);
                        umma_ss_cg2(tmem_d, dA_hi + adv, dB_lo + adv, idesc, 1u);
                        umma_ss_cg2(tmem_d, dA_hi + adv, dB_hi + adv, idesc, 1u);
                    }
                    umma_commit_cg2_mc(&empty_bar[stage]);
                    if (kin == TC_KC - 1 || kb == num_k_blocks - 1) umma_commit_cg2_mc(&tmem_full_bar[buf]);
                }
                __syncwarp();
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue (both CTAs): own 128 rows x 256 columns =====
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;
        const int half = ew >> 2;
        const int p = p0 + lane_base + lane;
        float acc[CPT];
#pragma unroll
        for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
        for (int chunk = 0; chunk < num_chunks; ++chunk) {
            const int buf = chunk & 1;
            mbar_wait(&tmem_full_bar[buf], (chunk >> 1) & 1);
            asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
            const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * BQ + half * CPT);
#pragma unroll
            for (int c = 0; c < CPT / 32; ++c) {
                uint32_t r[32];
                tmem_ld32(taddr + (uint32_t)(c * 32), r);
                asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
            }
            asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
            __syncwarp();
            if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
        }
        const float2 usc = epi.unscale(p0);
        float* cbase = epi.C + bz * epi.split_stride + (size_t)(q0 + half * CPT) * epi.ldc + p;
        if (epi.accumulate) {
            constexpr int RB = 32;
#pragma unroll
            for (int j0 = 0; j0 < CPT; j0 += RB) {
                float old[RB];
#pragma unroll
                for (int j = 0; j < RB; ++j) old[j] = __ldcg(cbase + (size_t)(j0 + j) * epi.ldc);
#pragma unroll
                for (int j = 0; j < RB; ++j) cbase[(size_t)(j0 + j) * epi.ldc] = acc[j0 + j] * usc.x * usc.y + old[j];
            }
        } else {
#pragma unroll
            for (int j = 0; j < CPT; ++j) cbase[(size_t)j * epi.ldc] = acc[j] * usc.x * usc.y;
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();                                      // nobody leaves while the peer may still read its shared memory / arrive on its barriers
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// Persistent multi-step forward kernel of the batched path (round 2): ONE cooperative launch integrates all T steps.
//
// CTA (bx, by) owns output tile (row tile bx, trial group by) for the whole horizon: barriers, TMEM and tensor maps are set up
// once, the smem ring and the TMEM ping-pong simply keep cycling through their phases, and the kernel boundary per step -- its
// launch gap, drain and ramp -- is replaced by a per-trial-group dependency: step t+1 of a tile needs src_{t+1} of ITS trials from
// all row tiles, nothing from the other trial groups.  Protocol per step:
//   epilogue warps   write y_{t+1}, src_{t+1} (buffer (t+1) & 1), fence, meet at the named barrier, one thread adds 1 to done[by]
//                    and to done[groups] (the all-groups counter);
//   TMA producer     before loading step t's operands: spin until done[by] == (#row tiles that write src) * t, then a
//                    generic->async proxy fence (the operands were written with ordinary stores by other SMs);
//   epilogue warps   before reading the scale of src_{t+1} (a maximum over ALL trials of step t-1's epilogues): spin until the
//                    all-groups counter reached (#writing CTAs) * t -- one main loop later than the increments, i.e. never in practice.
// The source operand is double buffered by step parity (with one launch per step a single buffer is enough only because the CTAs of
// a launch run in lock step).  A spin that lasts longer than ~2 s traps instead of hanging the device.
// ---------------------------------------------------------------------------------------------------------------------------
struct FwdPersist {
    int T, t_offset, T_total, S, cutoff;
    int n_row_tiles;               // tiles of kW rows (a further tile holds the readout rows when `readout`)
    int readout;                   // 1: grid.x = n_row_tiles + 1, the last row tile computes o_t = W_out . s_t
    const float* y_hist; size_t hslot;      // checkpoints [(T+1)][hslot] or nullptr
    float* y_pp; size_t slot;               // ping-pong state [2][slot] when no checkpoints are kept
    const float* x; size_t x_stride;
    float* out_rec; size_t out_stride;
    float* amax_src;               // binary16 + spiking: per-step maxima of the source operand (else nullptr)
    ScaleRef sc_static;            // binary16 + rate models: static bound of the activation
    void* src_hi[2]; void* src_lo[2];
    int nsv; size_t plane; int urec;        // ik: the recurrent drive goes into checkpoint plane nsv of slot t
    unsigned int* done;            // [groups + 1] zero-initialised
    int skew_ns;                   // trial group g starts g * skew_ns late: the HBM-bound epilogues of the groups then do not coincide
};

__device__ __forceinline__ unsigned int ld_acquire_gpu_u32(const unsigned int* p) {
    unsigned int v;
    asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ void spin_until(const unsigned int* p, unsigned int target) {
    if (ld_acquire_gpu_u32(p) >= target) return;
    const long long t0 = clock64();
    while (ld_acquire_gpu_u32(p) < target) {
        if (clock64() - t0 > 4000000000LL) __trap();           // ~2 s: a protocol error must not hang the device
    }
}

template <int BQ, class Epi, bool F16>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_fwd_persist(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                   const __grid_constant__ CUtensorMap tmB0_hi, const __grid_constant__ CUtensorMap tmB0_lo,
                   const __grid_constant__ CUtensorMap tmB1_hi, const __grid_constant__ CUtensorMap tmB1_lo,
                   int num_k_blocks, const __grid_constant__ Epi epi0, const __grid_constant__ FwdPersist ps) {
    using Cfg = TcCfg<BQ>;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CPT = Cfg::COLS_PER_THREAD;
    constexpr int TC_BK = TcElt<F16>::BK;
    constexpr int TC_KC = TcElt<F16>::KC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint64_t* tile_free_bar = tmem_empty_bar + 3;      // the epilogue has finished with the tile parked in the stage memory

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bx = blockIdx.x, by = blockIdx.y;
    const int p0 = bx * TC_BP, q0 = by * BQ;
    const int num_chunks = (num_k_blocks + TC_KC - 1) / TC_KC;
    // every CTA of a trial group takes part in the step counters -- also the readout tile, which writes no operand but READS the
    // source buffer that the others overwrite two steps later
    const unsigned int writers = gridDim.x;
    const unsigned int writers_all = writers * gridDim.y;
    const bool is_writer = true;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB0_hi); tma_prefetch_desc(&tmB0_lo);
        tma_prefetch_desc(&tmB1_hi); tma_prefetch_desc(&tmB1_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], TC_EPI_WARPS); }
        mbar_init(tile_free_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer: one elected thread, all steps =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            if (ps.skew_ns > 0 && by > 0) {
                const unsigned long long t_start = trace_now(), wait = (unsigned long long)by * (unsigned long long)ps.skew_ns;
                while (trace_now() - t_start < wait) __nanosleep(500);
            }
            for (int t = 0; t < ps.T; ++t) {
                const CUtensorMap* mb_hi = (t & 1) ? &tmB1_hi : &tmB0_hi;
                const CUtensorMap* mb_lo = (t & 1) ? &tmB1_lo : &tmB0_lo;
                bool ready = (t == 0);
                // the stages double as the parking space of the previous step's output tile: wait until this CTA's epilogue has left it
                if (t > 0) mbar_wait(tile_free_bar, (uint32_t)((t - 1) & 1));
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    mbar_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
                    // the weight tiles do not depend on the step: they are on their way while the source operand is awaited
                    tma_load_2d(sa, &tmA_hi, &full_bar[stage], kb * TC_BK, p0);
                    tma_load_2d(sa + Cfg::A_BYTES, &tmA_lo, &full_bar[stage], kb * TC_BK, p0);
                    if (!ready) {
                        // src_t of this trial group is complete when every CTA of the group has finished step t-1; the operand was
                        // written with ordinary stores by other SMs: generic -> async proxy fence before the TMA reads it
                        spin_until(ps.done + by, writers * (unsigned int)t);
                        asm volatile("fence.proxy.async;" ::: "memory");
                        ready = true;
                    }
                    tma_load_2d(sa + 2 * Cfg::A_BYTES, mb_hi, &full_bar[stage], kb * TC_BK, q0);
                    tma_load_2d(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES, mb_lo, &full_bar[stage], kb * TC_BK, q0);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: the chunk counter (TMEM ping-pong parity) runs on across the steps =====
        constexpr uint32_t idesc = make_idesc(F16, TC_BP, BQ);
        int stage = 0; uint32_t phase = 0;
        int chunk_g = 0;                                   // global chunk index
        for (int t = 0; t < ps.T; ++t) {
            for (int kb = 0; kb < num_k_blocks; ++kb) {
                const int kin = kb % TC_KC;
                const int buf = chunk_g & 1;
                if (kin == 0) {
                    mbar_wait(&tmem_empty_bar[buf], ((chunk_g >> 1) & 1) ^ 1);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                }
                mbar_wait(&full_bar[stage], phase);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                if (elect_one()) {
                    const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BQ);
                    const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                    const uint64_t dA_hi = make_sw128_kmajor_desc(sa);
                    const uint64_t dA_lo = make_sw128_kmajor_desc(sa + Cfg::A_BYTES);
                    const uint64_t dB_hi = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES);
                    const uint64_t dB_lo = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES + Cfg::B_BYTES);
#pragma unroll
                    for (int k = 0; k < TcElt<F16>::KSTEPS; ++k) {
                        const uint64_t adv = (uint64_t)((k * 32) >> 4);
                        umma_ss<F16>(tmem_d, dA_lo + adv, dB_hi + adv, idesc, (kin > 0 || k > 0) ? 1u : 0u);
                        umma_ss<F16>(tmem_d, dA_hi + adv, dB_lo + adv, idesc, 1u);
                        umma_ss<F16>(tmem_d, dA_hi + adv, dB_hi + adv, idesc, 1u);
                    }
                    umma_commit(&empty_bar[stage]);
                    if (kin == TC_KC - 1 || kb == num_k_blocks - 1) umma_commit(&tmem_full_bar[buf]);
                }
                __syncwarp();
                if (kin == TC_KC - 1 || kb == num_k_blocks - 1) ++chunk_g;
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        // ===== epilogue warps =====
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;
        const int half = ew >> 2;
        const bool f16_spk = F16 && ps.amax_src != nullptr;
        int chunk_g = 0;
        // record-window bookkeeping without a division per step (as in the persistent few-trial kernel)
        const int S_ = max(ps.S, 1);
        const int w_r0 = ((ps.cutoff + S_ - 1) / S_) * S_;
        int w_j, w_start, w_rec;
        if (ps.t_offset <= w_r0) { w_j = 0; w_start = ps.cutoff; w_rec = w_r0; }
        else { w_j = (ps.t_offset - w_r0 + S_ - 1) / S_; w_rec = w_r0 + w_j * S_; w_start = w_rec - S_ + 1; }
        for (int t = 0; t < ps.T; ++t) {
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
            for (int chunk = 0; chunk < num_chunks; ++chunk, ++chunk_g) {
                const int buf = chunk_g & 1;
                mbar_wait(&tmem_full_bar[buf], (chunk_g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * BQ + half * CPT);
#pragma unroll
                for (int c = 0; c < CPT / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive(&tmem_empty_bar[buf]);
            }
            // ---- per-step view of the epilogue arguments ----
            // A fresh copy per step, and only the operand scale before the accumulators are parked: the invariant fields then stay
            // operands in the constant bank and the per-step pointers / window state are not live while the 128 accumulators are
            // (with one copy outside the loop the kernel spilled 24..216 B depending on the build, and ran up to 12 % slower).
            Epi e = epi0;
            if (f16_spk) e.sB = t == 0 ? ScaleRef{ps.amax_src, 0.f, CV_HSRC} : ScaleRef{ps.amax_src + (t - 1), 1.f, CV_HSRC};
            else if (F16) e.sB = ps.sc_static;
            // the scale of src_{t+1} is a maximum over ALL trials of step t-1's epilogues
            if (t > 0 && f16_spk) { if (lane == 0) spin_until(ps.done + gridDim.y, writers_all * (unsigned int)t); __syncwarp(); }
            float2 usc = make_float2(1.f, 1.f);
            if constexpr (F16) usc = e.unscale(p0);
            float* tile = reinterpret_cast<float*>(smem);
            {
                float* my = tile + (size_t)(half * CPT) * TC_BP + lane_base + lane;
#pragma unroll
                for (int j = 0; j < CPT; ++j) my[j * TC_BP] = F16 ? acc[j] * usc.x * usc.y : acc[j];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float* cur = ps.y_hist ? ps.y_hist + (size_t)t * ps.hslot : ps.y_pp + (size_t)(t & 1) * ps.slot;
            float* nxt = ps.y_hist ? const_cast<float*>(ps.y_hist) + (size_t)(t + 1) * ps.hslot : ps.y_pp + (size_t)((t + 1) & 1) * ps.slot;
            e.a.y_cur = cur; e.a.y_next = nxt;
            e.a.x_t = ps.x ? ps.x + (size_t)t * ps.x_stride : nullptr;
            e.a.src_hi = ps.src_hi[(t + 1) & 1]; e.a.src_lo = ps.src_lo[(t + 1) & 1];
            e.a.urec_out = ps.urec ? const_cast<float*>(cur) + (size_t)ps.nsv * ps.plane : nullptr;
            if (f16_spk) {
                e.a.sc_out = ScaleRef{ps.amax_src + t, 1.f, CV_HSRC};
                e.a.amax_out = ps.amax_src + (t + 1);
            } else if (F16) {
                e.a.sc_out = ps.sc_static; e.a.amax_out = nullptr;
            }
            const int tg = ps.t_offset + t;
            if (tg > w_rec) { ++w_j; w_start = w_rec + 1; w_rec += S_; }
            const bool in_win = tg >= ps.cutoff && w_rec < ps.T_total;
            e.out_rec_j = (ps.readout && in_win && ps.out_rec) ? ps.out_rec + (size_t)w_j * ps.out_stride : nullptr;
            e.win_first = (tg == w_start); e.win_close = (tg == w_rec);
            e.inv_len = in_win ? 1.0f / (float)(w_rec - w_start + 1) : 0.f;
            e.template run_tile<BQ, F16>(p0, q0, tile, ew * 32 + lane, tile + (size_t)BQ * TC_BP);
            // publish: the named barrier orders every epilogue thread's stores before the one thread that fences (cumulativity, the
            // pattern of a cooperative-groups grid barrier) and bumps the counters
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ew == 0 && lane == 0) mbar_arrive(tile_free_bar);
            if (ew == 0 && lane == 0 && is_writer) {
                __threadfence();
                atomicAdd(ps.done + by, 1u);
                atomicAdd(ps.done + gridDim.y, 1u);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// split a dense fp32 matrix [rows][ld] into hi/lo copies [rows_out][ld_out] (zero padded): tf32-exact fp32 words, or
// binary16 with the power-of-two scale of `sc` (f16 = 1)
__global__ void __launch_bounds__(256) k_split_matrix(int rows, int cols, const float* __restrict__ src, int ld,
                                                       void* hi, void* lo, int ld_out, int rows_out, int f16, ScaleRef sc) {
    const size_t total = (size_t)rows_out * ld_out;
    const float scale = f16 ? exp2i(scale_expo(sc)) : 1.f;
    for (size_t idx = (size_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (size_t)gridDim.x * blockDim.x) {
        const int r = (int)(idx / ld_out), c = (int)(idx - (size_t)r * ld_out);
        const float x = (r < rows && c < cols) ? src[(size_t)r * ld + c] : 0.f;
        if (f16) {
            store_split1_f16(hi, lo, idx, x, scale);
        } else {
            float h, l;
            split_tf32(x, h, l);
            reinterpret_cast<float*>(hi)[idx] = h; reinterpret_cast<float*>(lo)[idx] = l;
        }
    }
}

// ---------------------------------------------------------------------------------------------------------------------------
// CTA-pair variant of the persistent forward kernel (the default; RP_NO_FWD_CG2=1 selects the 1-CTA kernel): the protocol of
// k_gemm_fwd_persist with the operand delivery of k_gemm_split3_cg2.  Two row tiles of the SAME trial group form a cluster and run
// M = 256 MMAs; each CTA loads its own 128 rows of kW and HALF of the source operand (32 instead of 48 KB of shared-memory fill per
// CTA and K block), keeps its own 128 x 256 accumulator and runs the unchanged epilogue on it.  grid.x is padded to an even number of
// row tiles (a tile past the readout rows loads zero rows and has nothing to step, but takes part in every counter).
// ---------------------------------------------------------------------------------------------------------------------------
template <class Epi>
__global__ void __launch_bounds__(TC_THREADS, 1)
k_gemm_fwd_persist_cg2(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                       const __grid_constant__ CUtensorMap tmB0_hi, const __grid_constant__ CUtensorMap tmB0_lo,      // 128-row boxes
                       const __grid_constant__ CUtensorMap tmB1_hi, const __grid_constant__ CUtensorMap tmB1_lo,
                       int num_k_blocks, const __grid_constant__ Epi epi0, const __grid_constant__ FwdPersist ps) {
    using Cfg = TcCfg2;
    constexpr int BQ = Cfg::BQ;
    constexpr int STAGES = Cfg::STAGES;
    constexpr int CPT = Cfg::COLS_PER_THREAD;
    constexpr int TC_BK = TcElt<true>::BK;
    constexpr int TC_KC = TcElt<true>::KC;
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * Cfg::STAGE_BYTES);      // used in the leader only
    uint64_t* empty_bar = full_bar + STAGES;
    uint64_t* tmem_full_bar = empty_bar + STAGES;      // [2]
    uint64_t* tmem_empty_bar = tmem_full_bar + 2;      // [2] used in the leader only (both CTAs' epilogue warps arrive)
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tmem_empty_bar + 2);
    uint64_t* tile_free_bar = tmem_empty_bar + 3;      // this CTA's epilogue has finished with the tile parked in ITS stage memory

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    uint32_t rank;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(rank));
    const bool leader = rank == 0;
    const int bx = blockIdx.x, by = blockIdx.y;
    const int p0 = bx * TC_BP, q0 = by * BQ;
    const int qh = q0 + (int)rank * (BQ / 2);
    const int num_chunks = (num_k_blocks + TC_KC - 1) / TC_KC;
    const unsigned int writers = gridDim.x;
    const unsigned int writers_all = writers * gridDim.y;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo); tma_prefetch_desc(&tmB0_hi); tma_prefetch_desc(&tmB0_lo);
        tma_prefetch_desc(&tmB1_hi); tma_prefetch_desc(&tmB1_lo);
        for (int s = 0; s < STAGES; ++s) { mbar_init(&full_bar[s], 1); mbar_init(&empty_bar[s], 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(&tmem_full_bar[s], 1); mbar_init(&tmem_empty_bar[s], 2 * TC_EPI_WARPS); }
        mbar_init(tile_free_bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_slot)), "n"(Cfg::TMEM_COLS) : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    const uint32_t tmem_base = *tmem_slot;

    if (warp == 0) {
        // ===== TMA producer (both CTAs), all steps =====
        if (elect_one()) {
            int stage = 0; uint32_t phase = 0;
            for (int t = 0; t < ps.T; ++t) {
                const CUtensorMap* mb_hi = (t & 1) ? &tmB1_hi : &tmB0_hi;
                const CUtensorMap* mb_lo = (t & 1) ? &tmB1_lo : &tmB0_lo;
                bool ready = (t == 0);
                if (t > 0) mbar_wait(tile_free_bar, (uint32_t)((t - 1) & 1));       // own stage memory no longer holds the parked tile
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    mbar_wait(&empty_bar[stage], phase ^ 1);
                    uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
                    if (leader) mbar_expect_tx(&full_bar[stage], 2 * Cfg::STAGE_BYTES);
                    tma_load_2d_cg2(sa, &tmA_hi, &full_bar[stage], kb * TC_BK, p0);
                    tma_load_2d_cg2(sa + Cfg::A_BYTES, &tmA_lo, &full_bar[stage], kb * TC_BK, p0);
                    if (!ready) {
                        spin_until(ps.done + by, writers * (unsigned int)t);
                        asm volatile("fence.proxy.async;" ::: "memory");
                        ready = true;
                    }
                    tma_load_2d_cg2(sa + 2 * Cfg::A_BYTES, mb_hi, &full_bar[stage], kb * TC_BK, qh);
                    tma_load_2d_cg2(sa + 2 * Cfg::A_BYTES + Cfg::BH_BYTES, mb_lo, &full_bar[stage], kb * TC_BK, qh);
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer: one elected thread of the LEADER CTA, the chunk counter runs on across the steps =====
        if (leader) {
            constexpr uint32_t idesc = make_idesc(true, 2 * TC_BP, BQ);
            int stage = 0; uint32_t phase = 0;
            int chunk_g = 0;
            for (int t = 0; t < ps.T; ++t) {
                for (int kb = 0; kb < num_k_blocks; ++kb) {
                    const int kin = kb % TC_KC;
                    const int buf = chunk_g & 1;
                    if (kin == 0) {
                        mbar_wait(&tmem_empty_bar[buf], ((chunk_g >> 1) & 1) ^ 1);
                        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    }
                    mbar_wait(&full_bar[stage], phase);
                    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                    if (elect_one()) {
                        const uint32_t tmem_d = tmem_base + (uint32_t)(buf * BQ);
                        const uint32_t sa = smem_u32(smem + stage * Cfg::STAGE_BYTES);
                        const uint64_t dA_hi = make_sw128_kmajor_desc(sa);
                        const uint64_t dA_lo = make_sw128_kmajor_desc(sa + Cfg::A_BYTES);
                        const uint64_t dB_hi = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES);
                        const uint64_t dB_lo = make_sw128_kmajor_desc(sa + 2 * Cfg::A_BYTES + Cfg::BH_BYTES);
#pragma unroll
                        for (int k = 0; k < TcElt<true>::KSTEPS; ++k) {
                            const uint64_t adv = (uint64_t)((k * 32) >> 4);
                            umma_ss_cg2(tmem_d, dA_lo + adv, dB_hi + adv, idesc, (kin > 0 || k > 0) ? 1u : 0u);
                            umma_ss_cg2(tmem_d, dA_hi + adv, dB_lo + adv, idesc, 1u);
                            umma_ss_cg2(tmem_d, dA_hi + adv, dB_hi + adv, idesc, 1u);
                        }
                        umma_commit_cg2_mc(&empty_bar[stage]);
                        if (kin == TC_KC - 1 || kb == num_k_blocks - 1) umma_commit_cg2_mc(&tmem_full_bar[buf]);
                    }
                    __syncwarp();
                    if (kin == TC_KC - 1 || kb == num_k_blocks - 1) ++chunk_g;
                    if (++stage == STAGES) { stage = 0; phase ^= 1; }
                }
            }
        }
    } else {
        // ===== epilogue warps (both CTAs): own 128 rows x 256 trials, exactly as in k_gemm_fwd_persist =====
        const int ew = warp - 2;
        const int lane_base = (warp & 3) * 32;
        const int half = ew >> 2;
        const bool f16_spk = ps.amax_src != nullptr;
        int chunk_g = 0;
        const int S_ = max(ps.S, 1);
        const int w_r0 = ((ps.cutoff + S_ - 1) / S_) * S_;
        int w_j, w_start, w_rec;
        if (ps.t_offset <= w_r0) { w_j = 0; w_start = ps.cutoff; w_rec = w_r0; }
        else { w_j = (ps.t_offset - w_r0 + S_ - 1) / S_; w_rec = w_r0 + w_j * S_; w_start = w_rec - S_ + 1; }
        for (int t = 0; t < ps.T; ++t) {
            float acc[CPT];
#pragma unroll
            for (int j = 0; j < CPT; ++j) acc[j] = 0.f;
            for (int chunk = 0; chunk < num_chunks; ++chunk, ++chunk_g) {
                const int buf = chunk_g & 1;
                mbar_wait(&tmem_full_bar[buf], (chunk_g >> 1) & 1);
                asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
                const uint32_t taddr = tmem_base + ((uint32_t)lane_base << 16) + (uint32_t)(buf * BQ + half * CPT);
#pragma unroll
                for (int c = 0; c < CPT / 32; ++c) {
                    uint32_t r[32];
                    tmem_ld32(taddr + (uint32_t)(c * 32), r);
                    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
                    for (int j = 0; j < 32; ++j) acc[c * 32 + j] += __uint_as_float(r[j]);
                }
                asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
                __syncwarp();
                if (lane == 0) mbar_arrive_leader(&tmem_empty_bar[buf]);
            }
            Epi e = epi0;
            if (f16_spk) e.sB = t == 0 ? ScaleRef{ps.amax_src, 0.f, CV_HSRC} : ScaleRef{ps.amax_src + (t - 1), 1.f, CV_HSRC};
            else e.sB = ps.sc_static;
            if (t > 0 && f16_spk) { if (lane == 0) spin_until(ps.done + gridDim.y, writers_all * (unsigned int)t); __syncwarp(); }
            const float2 usc = e.unscale(p0);
            float* tile = reinterpret_cast<float*>(smem);
            {
                float* my = tile + (size_t)(half * CPT) * TC_BP + lane_base + lane;
#pragma unroll
                for (int j = 0; j < CPT; ++j) my[j * TC_BP] = acc[j] * usc.x * usc.y;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            const float* cur = ps.y_hist ? ps.y_hist + (size_t)t * ps.hslot : ps.y_pp + (size_t)(t & 1) * ps.slot;
            float* nxt = ps.y_hist ? const_cast<float*>(ps.y_hist) + (size_t)(t + 1) * ps.hslot : ps.y_pp + (size_t)((t + 1) & 1) * ps.slot;
            e.a.y_cur = cur; e.a.y_next = nxt;
            e.a.x_t = ps.x ? ps.x + (size_t)t * ps.x_stride : nullptr;
            e.a.src_hi = ps.src_hi[(t + 1) & 1]; e.a.src_lo = ps.src_lo[(t + 1) & 1];
            e.a.urec_out = ps.urec ? const_cast<float*>(cur) + (size_t)ps.nsv * ps.plane : nullptr;
            if (f16_spk) { e.a.sc_out = ScaleRef{ps.amax_src + t, 1.f, CV_HSRC}; e.a.amax_out = ps.amax_src + (t + 1); }
            else { e.a.sc_out = ps.sc_static; e.a.amax_out = nullptr; }
            const int tg = ps.t_offset + t;
            if (tg > w_rec) { ++w_j; w_start = w_rec + 1; w_rec += S_; }
            const bool in_win = tg >= ps.cutoff && w_rec < ps.T_total;
            e.out_rec_j = (ps.readout && in_win && ps.out_rec) ? ps.out_rec + (size_t)w_j * ps.out_stride : nullptr;
            e.win_first = (tg == w_start); e.win_close = (tg == w_rec);
            e.inv_len = in_win ? 1.0f / (float)(w_rec - w_start + 1) : 0.f;
            e.template run_tile<BQ, true>(p0, q0, tile, ew * 32 + lane, tile + (size_t)BQ * TC_BP);
            asm volatile("bar.sync 1, 256;" ::: "memory");
            if (ew == 0 && lane == 0) {
                mbar_arrive(tile_free_bar);
                __threadfence();
                atomicAdd(ps.done + by, 1u);
                atomicAdd(ps.done + gridDim.y, 1u);
            }
        }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    __syncthreads();
    cluster_sync_all();
    if (warp == 1) {
        asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "n"(Cfg::TMEM_COLS) : "memory");
    }
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
inline PFN_cuTensorMapEncodeTiled_v12000 tc_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    if (!fn) {
        void* ptr = nullptr;
        cudaDriverEntryPointQueryResult qres;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess && qres == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(ptr);
    }
    return fn;
}

// 2D K-major operand map: global [rows][ld] elements (fp32 words or binary16), box = one swizzle row (K) x box_rows
inline int tc_make_map(CUtensorMap* map, const void* base, bool f16, int rows, int k_extent, int ld, int box_rows) {
    auto fn = tc_encode_fn();
    if (!fn) RP_TC_FAIL("cuTensorMapEncodeTiled entry point not available");
    const int esize = f16 ? 2 : 4;
    cuuint64_t gdim[2] = {(cuuint64_t)k_extent, (cuuint64_t)rows};
    cuuint64_t gstride[1] = {(cuuint64_t)ld * esize};
    cuuint32_t box[2] = {(cuuint32_t)(TC_ROW_BYTES / esize), (cuuint32_t)box_rows};
    cuuint32_t estr[2] = {1, 1};
    CUresult r = fn(map, f16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<void*>(base), gdim, gstride, box, estr,
                    CU_TENSOR_MAP_INTERLEAVE_NONE, TC_ROW_BYTES == 128 ? CU_TENSOR_MAP_SWIZZLE_128B : CU_TENSOR_MAP_SWIZZLE_64B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                    CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) RP_TC_FAIL("cuTensorMapEncodeTiled failed with CUresult %d (rows=%d k=%d ld=%d box_rows=%d f16=%d)", (int)r, rows, k_extent, ld, box_rows, (int)f16);
    return 0;
}

inline bool tc_supported(int N, int B) { return N % 128 == 0 && B % 128 == 0 && N >= 128 && B >= 128; }

// device-resident scalars of the binary16 path (one small float array per plan)
// TCM_CHUNK0 + 2*buffer + parity: reference maximum of the weight-gradient chunk being written into that operand buffer
// TCM_NB0..2: rotating bounds of max |g| of the fused reverse kernel (read / accumulate / clear, see k_adj_fused_f16)
enum { TCM_AMAX_W = 0, TCM_AMAX_WOUT, TCM_SRC_BOUND, TCM_G_AMAX0, TCM_G_AMAX1, TCM_CHUNK0, TCM_FLAGS = TCM_CHUNK0 + 4, TCM_NB0, TCM_COUNT = 16 };
static_assert(TCM_NB0 + 3 <= TCM_COUNT, "meta slots");

struct TcWorkspace {
    int N = 0, B = 0, ldk = 0, ldt = 0, wgrad_chunk = 0;
    bool f16 = false;
    void *W_hi = nullptr, *W_lo = nullptr, *WT_hi = nullptr, *WT_lo = nullptr;       // [N][ldk]
    void *src_hi = nullptr, *src_lo = nullptr, *g_hi = nullptr, *g_lo = nullptr;     // [B][ldk]
    // weight-gradient operands, trial-major [N][ldt]; two buffers so that the contraction of one K chunk (side stream) overlaps the
    // reverse steps that fill the other (binary16 path; the tf32 path uses buffer 0 only)
    void *gT_hi[2] = {nullptr, nullptr}, *gT_lo[2] = {nullptr, nullptr}, *srcT_hi[2] = {nullptr, nullptr}, *srcT_lo[2] = {nullptr, nullptr};
    cudaStream_t ws = nullptr;                                    // side stream of the overlapped weight-gradient slices
    cudaEvent_t ev_z = nullptr, ev_ops[2] = {nullptr, nullptr}, ev_done[2] = {nullptr, nullptr}, ev_join = nullptr;
    float* g32 = nullptr;        // binary16 path: g_{t-1} in fp32 [B][N] before the exact-maximum conversion
    float* src32 = nullptr;      // binary16 path, rate models: act(v_{t-1}) [B][N]
    float* meta = nullptr;       // [TCM_COUNT]
    float* amax_src = nullptr;   // [amax_cap] max |src_t| of the steps of the last forward call
    int amax_cap = 0;
    const void* fwd_history = nullptr; int fwd_T = -1;      // which checkpoints amax_src describes
    const void* wt_W = nullptr;                             // weights whose transposed split (WT_hi/lo) the last forward call left behind
    CUtensorMap m_g_h[2], m_gT_h[2][2];                     // 128-row boxes of the B operands (CTA-pair kernel)
    void *src2_hi = nullptr, *src2_lo = nullptr;            // second source-operand buffer of the persistent multi-step forward kernel
    CUtensorMap m_src2[2];
    CUtensorMap m_src_h[2], m_src2_h[2];                    // 128-row boxes of the two source buffers (CTA-pair persistent forward kernel)
    unsigned int* fwd_done = nullptr;                       // [B / bq_fwd + 1] step counters of that kernel
    int bq_fwd = 0, bq_wg = 0;
    CUtensorMap m_W[2], m_WT[2], m_src[2], m_g[2], m_gT[2][2], m_srcT[2][2];     // m_gT[buffer][hi/lo]
    int esize() const { return f16 ? 2 : 4; }
};

inline int tc_alloc(void** p, size_t bytes_wanted, size_t* bytes) {
    if (cudaMalloc(p, bytes_wanted) != cudaSuccess) RP_TC_FAIL("cudaMalloc of %zu bytes failed", bytes_wanted);
    const char* poison = getenv("RP_POISON_TC_ALLOC");        // debug only: these buffers are zero-initialised BY CONTRACT (scalars, counters, padding)
    if (cudaMemset(*p, poison ? atoi(poison) : 0, bytes_wanted) != cudaSuccess) RP_TC_FAIL("cudaMemset failed");
    *bytes += bytes_wanted;
    return 0;
}

template <class Epi, bool F16>
inline int tc_set_attrs() {
    static bool done = false;
    if (done) return 0;
    cudaError_t e = cudaFuncSetAttribute(k_gemm_split3<256, Epi, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_split3<128, Epi, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES);
    if constexpr (std::is_same<Epi, EpiStore>::value && F16) {
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_split3_lean<256, Epi, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_split3_lean<128, Epi, F16>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<128>::SMEM_BYTES);
        // the full 228 KB carveout (the default would be the smallest one that holds the CTA, leaving no room for a co-resident block)
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_split3_lean<256, Epi, F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(k_gemm_split3_lean<128, Epi, F16>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    }
    if (e != cudaSuccess) RP_TC_FAIL("cudaFuncSetAttribute(max dynamic smem) failed: %s", cudaGetErrorString(e));
    done = true;
    return 0;
}

inline void tc_workspace_destroy(TcWorkspace* w) {
    void* bufs[] = {w->W_hi, w->W_lo, w->WT_hi, w->WT_lo, w->src_hi, w->src_lo, w->g_hi, w->g_lo, w->gT_hi[0], w->gT_lo[0], w->srcT_hi[0], w->srcT_lo[0],
                    w->gT_hi[1], w->gT_lo[1], w->srcT_hi[1], w->srcT_lo[1], w->g32, w->src32, w->meta, w->amax_src, w->src2_hi, w->src2_lo, w->fwd_done};
    for (void* b : bufs) if (b) cudaFree(b);
    cudaEvent_t evs[] = {w->ev_z, w->ev_ops[0], w->ev_ops[1], w->ev_done[0], w->ev_done[1], w->ev_join};
    for (cudaEvent_t e : evs) if (e) cudaEventDestroy(e);
    if (w->ws) cudaStreamDestroy(w->ws);
    *w = TcWorkspace();
}

inline int tc_workspace_create(TcWorkspace* w, int N, int B, bool f16, bool rate_model, size_t* bytes) {
    w->N = N; w->B = B; w->ldk = N; w->f16 = f16;
    const size_t es = w->esize();
    // weight-gradient K chunk: several steps' (g, src) columns per GEMM so that the read-modify-write of dW amortises
    // (measured at N=4096, B=1024: 83 us per step at 4 steps per chunk, 72 at 8, 66 at 16)
    // round 2: 32 steps per chunk (K = 32768 at 1024 trials): 30.54 vs 31.34 ms per pass at 16 (same box, back to back)
    int chunk = 32768 / B; if (chunk < 1) chunk = 1; if (chunk > 32) chunk = 32;
    if (getenv("RP_WG_CHUNK") && atoi(getenv("RP_WG_CHUNK")) > 0) chunk = atoi(getenv("RP_WG_CHUNK"));     // tuning experiments
    w->wgrad_chunk = chunk;
    w->ldt = chunk * B;
    w->bq_fwd = (B % 256 == 0) ? 256 : 128;
    if (getenv("RP_TC_BQ_FWD") && atoi(getenv("RP_TC_BQ_FWD")) == 128) w->bq_fwd = 128;      // tuning experiments
    w->bq_wg = (N % 256 == 0) ? 256 : 128;
    const size_t nn = (size_t)N * w->ldk * es, bn = (size_t)B * w->ldk * es;
    const size_t nw = (size_t)(N + TC_BP) * w->ldk * es;        // + one row tile for the fused readout rows (W_out)
    if (tc_alloc(&w->W_hi, nw, bytes) || tc_alloc(&w->W_lo, nw, bytes) || tc_alloc(&w->WT_hi, nn, bytes) || tc_alloc(&w->WT_lo, nn, bytes)) return 1;
    if (tc_alloc(&w->src_hi, bn, bytes) || tc_alloc(&w->src_lo, bn, bytes) || tc_alloc(&w->g_hi, bn, bytes) || tc_alloc(&w->g_lo, bn, bytes)) return 1;
    if (f16) {
        if (tc_alloc(reinterpret_cast<void**>(&w->meta), TCM_COUNT * sizeof(float), bytes)) return 1;
        if (tc_alloc(reinterpret_cast<void**>(&w->g32), (size_t)B * N * sizeof(float), bytes)) return 1;
        if (rate_model && tc_alloc(reinterpret_cast<void**>(&w->src32), (size_t)B * N * sizeof(float), bytes)) return 1;
    }
    if (tc_make_map(&w->m_W[0], w->W_hi, f16, N + TC_BP, N, w->ldk, TC_BP) || tc_make_map(&w->m_W[1], w->W_lo, f16, N + TC_BP, N, w->ldk, TC_BP)) return 1;
    if (tc_make_map(&w->m_WT[0], w->WT_hi, f16, N, N, w->ldk, TC_BP) || tc_make_map(&w->m_WT[1], w->WT_lo, f16, N, N, w->ldk, TC_BP)) return 1;
    if (tc_make_map(&w->m_src[0], w->src_hi, f16, B, N, w->ldk, w->bq_fwd) || tc_make_map(&w->m_src[1], w->src_lo, f16, B, N, w->ldk, w->bq_fwd)) return 1;
    if (tc_make_map(&w->m_g[0], w->g_hi, f16, B, N, w->ldk, w->bq_fwd) || tc_make_map(&w->m_g[1], w->g_lo, f16, B, N, w->ldk, w->bq_fwd)) return 1;
    if (tc_make_map(&w->m_g_h[0], w->g_hi, f16, B, N, w->ldk, 128) || tc_make_map(&w->m_g_h[1], w->g_lo, f16, B, N, w->ldk, 128)) return 1;
    return 0;
}

// per-step source maxima of a forward call (binary16 path): grown on demand, zeroed by the caller
inline int tc_workspace_ensure_amax(TcWorkspace* w, int n, size_t* bytes) {
    if (n <= w->amax_cap) return 0;
    if (w->amax_src) cudaFree(w->amax_src);
    w->amax_src = nullptr; w->amax_cap = 0;
    const int cap = n + 1024;
    if (tc_alloc(reinterpret_cast<void**>(&w->amax_src), (size_t)cap * sizeof(float), bytes)) return 1;
    w->amax_cap = cap;
    return 0;
}

// the transposed (trial-major) operand buffers of the weight gradient are only needed by rp_backward
inline int tc_workspace_ensure_wgrad(TcWorkspace* w, size_t* bytes) {
    if (w->gT_hi[0]) return 0;
    const size_t nt = (size_t)w->N * w->ldt * w->esize();
    const bool f16 = w->f16;
    const bool side = f16 && (getenv("RP_WG_OVERLAP") || getenv("RP_WG_IDLE_SLICES"));      // the opt-in overlap variants
    const int nbuf = side ? 2 : 1;
    for (int c = 0; c < nbuf; ++c) {
        if (tc_alloc(&w->gT_hi[c], nt, bytes) || tc_alloc(&w->gT_lo[c], nt, bytes) || tc_alloc(&w->srcT_hi[c], nt, bytes) || tc_alloc(&w->srcT_lo[c], nt, bytes)) return 1;
        if (tc_make_map(&w->m_srcT[c][0], w->srcT_hi[c], f16, w->N, w->ldt, w->ldt, TC_BP) || tc_make_map(&w->m_srcT[c][1], w->srcT_lo[c], f16, w->N, w->ldt, w->ldt, TC_BP)) return 1;
        if (tc_make_map(&w->m_gT[c][0], w->gT_hi[c], f16, w->N, w->ldt, w->ldt, w->bq_wg) || tc_make_map(&w->m_gT[c][1], w->gT_lo[c], f16, w->N, w->ldt, w->ldt, w->bq_wg)) return 1;
        if (tc_make_map(&w->m_gT_h[c][0], w->gT_hi[c], f16, w->N, w->ldt, w->ldt, 128) || tc_make_map(&w->m_gT_h[c][1], w->gT_lo[c], f16, w->N, w->ldt, w->ldt, 128)) return 1;
    }
    if (side) {
        int lo = 0, hi = 0;
        cudaDeviceGetStreamPriorityRange(&lo, &hi);           // lo = least priority (the default), hi = greatest
        const int prio = getenv("RP_WS_HIGH_PRIORITY") ? hi : lo;
        if (cudaStreamCreateWithPriority(&w->ws, cudaStreamNonBlocking, prio) != cudaSuccess) RP_TC_FAIL("cudaStreamCreateWithPriority failed");
        cudaEvent_t* evs[] = {&w->ev_z, &w->ev_ops[0], &w->ev_ops[1], &w->ev_done[0], &w->ev_done[1], &w->ev_join};
        for (cudaEvent_t* e : evs) if (cudaEventCreateWithFlags(e, cudaEventDisableTiming) != cudaSuccess) RP_TC_FAIL("cudaEventCreate failed");
    }
    return 0;
}

// item0/items: launch only the work items [item0, item0 + items) of the full grid (items <= 0: all of it)
template <class Epi, bool F16>
inline int tc_launch_epi(int bq, int P, int Q, int K, const CUtensorMap* A, const CUtensorMap* Bm, const Epi& epi, cudaStream_t st, int k_splits = 1,
                         int item0 = 0, int items = 0) {
    constexpr int BK = TcElt<F16>::BK;
    if (P % TC_BP || Q % bq || K % (BK * k_splits)) RP_TC_FAIL("tc_launch: extents P=%d Q=%d K=%d do not match tile %dx%dx%d (x%d splits)", P, Q, K, TC_BP, bq, BK, k_splits);
    if (tc_set_attrs<Epi, F16>()) return 1;
    dim3 grid(P / TC_BP, Q / bq, k_splits);
    TcSlice slice{0, 0, 0};
    if (items > 0) {
        const int total = (int)(grid.x * grid.y * grid.z);
        if (item0 < 0 || item0 + items > total) RP_TC_FAIL("tc_launch: work items [%d, %d) outside the grid of %d", item0, item0 + items, total);
        slice = TcSlice{item0, (int)grid.x, (int)grid.y};
        grid = dim3(items, 1, 1);
    }
    const int kb = K / BK / k_splits;
    bool lean = false;
    if constexpr (std::is_same<Epi, EpiStore>::value && F16) lean = items > 0 && getenv("RP_WG_OVERLAP") && !getenv("RP_NO_LEAN_SLICES");
    if (lean) {
        if constexpr (std::is_same<Epi, EpiStore>::value && F16) {
            if (bq == 256) launch_pdl(k_gemm_split3_lean<256, Epi, F16>, grid, dim3(TC_THREADS), TcCfg<256>::SMEM_BYTES, st, A[0], A[1], Bm[0], Bm[1], kb, epi, slice);
            else           launch_pdl(k_gemm_split3_lean<128, Epi, F16>, grid, dim3(TC_THREADS), TcCfg<128>::SMEM_BYTES, st, A[0], A[1], Bm[0], Bm[1], kb, epi, slice);
        }
    } else if (bq == 256) launch_pdl(k_gemm_split3<256, Epi, F16>, grid, dim3(TC_THREADS), TcCfg<256>::SMEM_BYTES, st, A[0], A[1], Bm[0], Bm[1], kb, epi, slice);
    else                  launch_pdl(k_gemm_split3<128, Epi, F16>, grid, dim3(TC_THREADS), TcCfg<128>::SMEM_BYTES, st, A[0], A[1], Bm[0], Bm[1], kb, epi, slice);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) RP_TC_FAIL("tcgen05 GEMM launch failed: %s", cudaGetErrorString(e));
    return 0;
}
// CTA-pair launch of the plain contraction: Bh = maps of the B operand with 128-row boxes; cluster (2,1,1) along the row tiles
// default on (A/B on one box, alternating: pass 31.99 / 31.87 ms vs 32.40 / 32.76 ms; weight-gradient chunk alone 1.94 vs 2.05 ms, adjoint
// product alone 80.0 vs 81.3 us); RP_NO_TC_CG2=1 restores the single-CTA kernel
inline bool tc_cg2_enabled() { static const bool on = getenv("RP_NO_TC_CG2") == nullptr; return on; }
inline int tc_launch_cg2(int P, int Q, int K, const CUtensorMap* A, const CUtensorMap* Bh, const EpiStore& epi, cudaStream_t st, int k_splits) {
    constexpr int BK = TcElt<true>::BK;
    if (P % (2 * TC_BP) || Q % 256 || K % (BK * k_splits)) RP_TC_FAIL("tc_launch_cg2: extents P=%d Q=%d K=%d", P, Q, K);
    static bool attr_done = false;
    if (!attr_done) {
        if (cudaFuncSetAttribute(k_gemm_split3_cg2, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg2::SMEM_BYTES) != cudaSuccess) RP_TC_FAIL("cudaFuncSetAttribute failed");
        attr_done = true;
    }
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(P / TC_BP, Q / 256, k_splits); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TcCfg2::SMEM_BYTES; cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr; cfg.numAttrs = 1;
    const int kb = K / BK / k_splits;
    cudaError_t e = cudaLaunchKernelEx(&cfg, k_gemm_split3_cg2, A[0], A[1], Bh[0], Bh[1], kb, epi);
    if (e != cudaSuccess) RP_TC_FAIL("CTA-pair GEMM launch failed: %s", cudaGetErrorString(e));
    return 0;
}
inline int tc_launch(bool f16, int bq, int P, int Q, int K, const CUtensorMap* A, const CUtensorMap* Bm, float* C, int ldc, int accumulate, cudaStream_t st,
                     int k_splits = 1, size_t split_stride = 0, ScaleRef sa = no_scale(), ScaleRef sb = no_scale(), int item0 = 0, int items = 0,
                     const CUtensorMap* Bhalf = nullptr) {
    EpiStore e{C, ldc, accumulate, split_stride, sa, sb};
    if (f16 && bq == 256 && Bhalf != nullptr && items <= 0 && P % (2 * TC_BP) == 0 && tc_cg2_enabled()) return tc_launch_cg2(P, Q, K, A, Bhalf, e, st, k_splits);
    if (f16) return tc_launch_epi<EpiStore, true>(bq, P, Q, K, A, Bm, e, st, k_splits, item0, items);
    return tc_launch_epi<EpiStore, false>(bq, P, Q, K, A, Bm, e, st, k_splits, item0, items);
}
// fused launch: forward step (rows = N, or N+128 when the readout rows are appended)
template <int MODEL, bool GEN>
// `src_buf`: which of the two source buffers holds src_t (the epilogue writes src_{t+1} into the OTHER one: CTAs of a launch finish
// their K loops at different times, and a tile that is done must not overwrite operand columns another CTA has yet to load)
inline int tc_forward_step(TcWorkspace* w, const EpiFwd<MODEL, GEN>& epi, bool readout_rows, cudaStream_t st, int src_buf = 0) {
    const int P = w->N + (readout_rows ? TC_BP : 0);
    const CUtensorMap* mb = src_buf ? w->m_src2 : w->m_src;
    if (w->f16) return tc_launch_epi<EpiFwd<MODEL, GEN>, true>(w->bq_fwd, P, w->B, w->N, w->m_W, mb, epi, st);
    return tc_launch_epi<EpiFwd<MODEL, GEN>, false>(w->bq_fwd, P, w->B, w->N, w->m_W, mb, epi, st);
}
// second source buffer + step counters of the persistent multi-step forward kernel (allocated on first use)
inline int tc_workspace_ensure_persist(TcWorkspace* w, size_t* bytes) {
    if (w->src2_hi) return 0;
    const size_t bn = (size_t)w->B * w->ldk * w->esize();
    if (tc_alloc(&w->src2_hi, bn, bytes) || tc_alloc(&w->src2_lo, bn, bytes)) return 1;
    if (tc_alloc(reinterpret_cast<void**>(&w->fwd_done), 64 * sizeof(unsigned int), bytes)) return 1;
    if (tc_make_map(&w->m_src2[0], w->src2_hi, w->f16, w->B, w->N, w->ldk, w->bq_fwd) || tc_make_map(&w->m_src2[1], w->src2_lo, w->f16, w->B, w->N, w->ldk, w->bq_fwd)) return 1;
    if (tc_make_map(&w->m_src_h[0], w->src_hi, w->f16, w->B, w->N, w->ldk, 128) || tc_make_map(&w->m_src_h[1], w->src_lo, w->f16, w->B, w->N, w->ldk, 128)) return 1;
    if (tc_make_map(&w->m_src2_h[0], w->src2_hi, w->f16, w->B, w->N, w->ldk, 128) || tc_make_map(&w->m_src2_h[1], w->src2_lo, w->f16, w->B, w->N, w->ldk, 128)) return 1;
    return 0;
}
// whole horizon in one cooperative launch (binary16 operands, 256-trial tiles); returns 2 when the grid cannot be co-resident
template <int MODEL, bool GEN>
inline int tc_forward_persistent(TcWorkspace* w, const EpiFwd<MODEL, GEN>& epi, FwdPersist ps, int sm_count, cudaStream_t st) {
    using Epi = EpiFwd<MODEL, GEN>;
    auto kernel = k_gemm_fwd_persist<256, Epi, true>;
    auto kernel2 = k_gemm_fwd_persist_cg2<Epi>;
    static bool attr_done = false, attr2_done = false;
    const int gx1 = ps.n_row_tiles + (ps.readout ? 1 : 0);
    if ((int)(w->B / 256) + 1 > 64) return 2;
    ps.src_hi[0] = w->src_hi; ps.src_lo[0] = w->src_lo; ps.src_hi[1] = w->src2_hi; ps.src_lo[1] = w->src2_lo;
    ps.done = w->fwd_done;
    int kb = w->N / TcElt<true>::BK;
    Epi e = epi;
    // CTA-pair variant first (A/B on one box, alternating, 4 runs each: pass 30.43 vs 30.70 ms, forward stage 9.64 vs 9.82 ms per 100 steps,
    // bit-identical results); RP_NO_FWD_CG2=1, a grid of pairs that is not co-resident, or a refused cluster launch fall through to the
    // 1-CTA kernel.  Under Nsight Compute the cooperative cluster launch is refused asynchronously (LaunchFailed, a sticky error; the same
    // command runs clean without the profiler), so a profiled / injected process (NV_NSIGHT_INJECTION_TRANSPORT_TYPE, CUDA_INJECTION64_PATH)
    // gets the 1-CTA kernel: same protocol, same epilogue, bit-identical results.
    if (!getenv("RP_NO_FWD_CG2") && !getenv("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") && !getenv("CUDA_INJECTION64_PATH")) {
        if (!attr2_done) {
            if (cudaFuncSetAttribute(kernel2, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg2::SMEM_BYTES) != cudaSuccess) RP_TC_FAIL("cudaFuncSetAttribute failed");
            attr2_done = true;
        }
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3((gx1 + 1) & ~1, w->B / 256); cfg.blockDim = dim3(TC_THREADS); cfg.dynamicSmemBytes = TcCfg2::SMEM_BYTES; cfg.stream = st;
        cudaLaunchAttribute at[2];
        at[0].id = cudaLaunchAttributeClusterDimension; at[0].val.clusterDim.x = 2; at[0].val.clusterDim.y = 1; at[0].val.clusterDim.z = 1;
        at[1].id = cudaLaunchAttributeCooperative; at[1].val.cooperative = 1;
        cfg.attrs = at; cfg.numAttrs = 2;
        int ncl = 0;
        if (cudaOccupancyMaxActiveClusters(&ncl, kernel2, &cfg) == cudaSuccess && 2 * ncl >= (int)(cfg.gridDim.x * cfg.gridDim.y)) {
            if (cudaMemsetAsync(w->fwd_done, 0, 64 * sizeof(unsigned int), st) != cudaSuccess) RP_TC_FAIL("cudaMemsetAsync failed");
            if (cudaLaunchKernelEx(&cfg, kernel2, w->m_W[0], w->m_W[1], w->m_src_h[0], w->m_src_h[1], w->m_src2_h[0], w->m_src2_h[1], kb, e, ps) == cudaSuccess) return 0;
        }
        cudaGetLastError();
    }
    if (!attr_done) {
        if (cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, TcCfg<256>::SMEM_BYTES) != cudaSuccess) RP_TC_FAIL("cudaFuncSetAttribute failed");
        attr_done = true;
    }
    const dim3 grid(gx1, w->B / 256);
    int occ = 0;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kernel, TC_THREADS, TcCfg<256>::SMEM_BYTES) != cudaSuccess || occ < 1) return 2;
    if ((int)(grid.x * grid.y) > occ * sm_count) return 2;
    if (cudaMemsetAsync(w->fwd_done, 0, 64 * sizeof(unsigned int), st) != cudaSuccess) RP_TC_FAIL("cudaMemsetAsync failed");
    void* args[] = {&w->m_W[0], &w->m_W[1], &w->m_src[0], &w->m_src[1], &w->m_src2[0], &w->m_src2[1], &kb, &e, &ps};
    cudaError_t err = cudaLaunchCooperativeKernel(reinterpret_cast<void*>(kernel), grid, dim3(TC_THREADS), args, TcCfg<256>::SMEM_BYTES, st);
    if (err != cudaSuccess) { cudaGetLastError(); return 2; }      // e.g. not co-resident next to another tenant's kernels: per-step launches instead
    return 0;
}
// scale descriptors of the binary16 operands (no-ops on the tf32 path)
inline ScaleRef tc_scale_W(const TcWorkspace* w) { return w->f16 ? ScaleRef{w->meta + TCM_AMAX_W, 0.f, CV_HG} : no_scale(); }
inline ScaleRef tc_scale_Wout(const TcWorkspace* w) { return w->f16 ? ScaleRef{w->meta + TCM_AMAX_WOUT, 0.f, CV_HG} : no_scale(); }
inline ScaleRef tc_scale_srcbound(const TcWorkspace* w) { return w->f16 ? ScaleRef{w->meta + TCM_SRC_BOUND, 0.f, CV_HSRC} : no_scale(); }

// mode TC_FWD:   C=u[b][i]     A = kW (hi,lo)      B = src (hi,lo)     K = N
// mode TC_DGRAD: C=Z[b][j]     A = (kW)^T          B = g               K = N
// mode TC_WGRAD: C=dW[i][j] += A = src^T [j][(t,b)] B = g^T [i][(t,b)] K = k_extent (columns filled in the chunk)
// sb: scale of the B operand (binary16 path): src_t / g_t / the weight-gradient chunk; the A scale follows from the mode
// WGRAD only: cb = operand buffer, item0/items = slice of the work items (items <= 0: all)
inline int tc_wgrad_items(const TcWorkspace* w) { return (w->N / TC_BP) * (w->N / w->bq_wg) * TC_WGRAD_SPLITS; }
inline int tc_gemm(TcWorkspace* w, int mode, float* C, int ldc, int k_extent, int accumulate, cudaStream_t st, ScaleRef sb = no_scale(),
                   int cb = 0, int item0 = 0, int items = 0) {
    switch (mode) {
        case TC_FWD:   return tc_launch(w->f16, w->bq_fwd, w->N, w->B, w->N, w->m_W, w->m_src, C, ldc, accumulate, st, 1, 0, tc_scale_W(w), sb);
        case TC_DGRAD: return tc_launch(w->f16, w->bq_fwd, w->N, w->B, w->N, w->m_WT, w->m_g, C, ldc, accumulate, st, 1, 0, tc_scale_W(w), sb, 0, 0, w->m_g_h);
        // 2-way split-K into two accumulation slices: N=4096 gives 512 tiles = 3.46 waves of 148 SMs (86 % filled);
        // 1024 work items = 6.92 waves (99 %).  The slices are summed by k_finish_wgrad.
        case TC_WGRAD: return tc_launch(w->f16, w->bq_wg, w->N, w->N, k_extent, w->m_srcT[cb], w->m_gT[cb], C, ldc, accumulate, st, TC_WGRAD_SPLITS, (size_t)w->N * ldc,
                                        tc_scale_srcbound(w), sb, item0, items, w->m_gT_h[cb]);
    }
    RP_TC_FAIL("tc_gemm: unknown mode %d", mode);
}

// test entry: arbitrary K-major fp32 operands (split on the fly into temporaries)
inline int tc_gemm_standalone(bool f16, int P, int Q, int K, const float* A, int lda, const float* B, int ldb, float* C, int ldc,
                              int accumulate, cudaStream_t st) {
    if (P % TC_BP || Q % 128 || K <= 0) RP_TC_FAIL("split-3 GEMM needs P %% 128 == 0, Q %% 128 == 0 (got P=%d Q=%d K=%d)", P, Q, K);
    const int BK = f16 ? TcElt<true>::BK : TcElt<false>::BK;
    const int Kp = (K + BK - 1) / BK * BK;
    const int bq = (Q % 256 == 0) ? 256 : 128;
    const size_t es = f16 ? 2 : 4;
    uint8_t* tmp = nullptr;
    float* amax = nullptr;
    const size_t na = (size_t)P * Kp * es, nb = (size_t)Q * Kp * es;
    if (cudaMalloc(reinterpret_cast<void**>(&tmp), 2 * (na + nb)) != cudaSuccess) RP_TC_FAIL("cudaMalloc of split temporaries failed");
    if (cudaMalloc(reinterpret_cast<void**>(&amax), 2 * sizeof(float)) != cudaSuccess) { cudaFree(tmp); RP_TC_FAIL("cudaMalloc failed"); }
    cudaMemsetAsync(amax, 0, 2 * sizeof(float), st);
    void *a_hi = tmp, *a_lo = tmp + na, *b_hi = tmp + 2 * na, *b_lo = tmp + 2 * na + nb;
    ScaleRef sa = no_scale(), sb = no_scale();
    if (f16) {
        k_amax_2d<<<256, 256, 0, st>>>(P, K, A, (size_t)lda, nullptr, 0, amax);
        k_amax_2d<<<256, 256, 0, st>>>(Q, K, B, (size_t)ldb, nullptr, 0, amax + 1);
        sa = ScaleRef{amax, 0.f, CV_HG}; sb = ScaleRef{amax + 1, 0.f, CV_HG};
    }
    k_split_matrix<<<1024, 256, 0, st>>>(P, K, A, lda, a_hi, a_lo, Kp, P, f16 ? 1 : 0, sa);
    k_split_matrix<<<1024, 256, 0, st>>>(Q, K, B, ldb, b_hi, b_lo, Kp, Q, f16 ? 1 : 0, sb);
    CUtensorMap mA[2], mB[2];
    int rc = tc_make_map(&mA[0], a_hi, f16, P, Kp, Kp, TC_BP) || tc_make_map(&mA[1], a_lo, f16, P, Kp, Kp, TC_BP) ||
             tc_make_map(&mB[0], b_hi, f16, Q, Kp, Kp, bq) || tc_make_map(&mB[1], b_lo, f16, Q, Kp, Kp, bq);
    CUtensorMap mBh[2];
    if (!rc && f16 && bq == 256) rc = tc_make_map(&mBh[0], b_hi, f16, Q, Kp, Kp, 128) || tc_make_map(&mBh[1], b_lo, f16, Q, Kp, Kp, 128);
    if (!rc) rc = tc_launch(f16, bq, P, Q, Kp, mA, mB, C, ldc, accumulate, st, 1, 0, sa, sb, 0, 0, (f16 && bq == 256) ? mBh : nullptr);
    cudaStreamSynchronize(st);
    cudaFree(tmp);
    cudaFree(amax);
    return rc;
}

}  // namespace rp
