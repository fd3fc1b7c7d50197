// C-ABI entry points of the rectipy_b200 engine (see include/rectipy_b200.h for the contract and the
// reference lines each entry replaces).  Host side: plan/workspace management and the per-step launch
// sequences; device side: rp_kernels.cuh (element-wise), rp_gemm_simt.cuh (FFMA), rp_gemm_tc.cuh (tcgen05).
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <algorithm>
#include <new>
#include <vector>

#include "../../include/rectipy_b200.h"
#include "rp_kernels.cuh"
#include "rp_gemm_simt.cuh"
#include "rp_gemm_tc.cuh"
#include "rp_rls.cuh"
#include "rp_persistent.cuh"
#include "rp_fp32_paths.cuh"
#include <stdlib.h>

namespace {

thread_local char g_err[512] = "";

int fail(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
    return 1;
}

#define RP_CUDA(expr)                                                                          \
    do {                                                                                       \
        cudaError_t _e = (expr);                                                               \
        if (_e != cudaSuccess) return fail("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

#define RP_LAUNCH_CHECK()                                                                      \
    do {                                                                                       \
        cudaError_t _e = cudaGetLastError();                                                   \
        if (_e != cudaSuccess) return fail("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
    } while (0)

inline int nsv_of(int model) {
    switch (model) {
        case RP_LI_TANH: case RP_LI_SIGMOID: return 1;
        case RP_QIF: case RP_LIF: return 2;
        case RP_QIF_SFA: case RP_IK: case RP_IKU: return 3;
        case RP_IK_BIEXP: return 4;
        default: return -1;
    }
}
inline bool spiking(int model) { return model == RP_QIF || model == RP_QIF_SFA || model == RP_LIF || rp::is_ik(model); }
inline int nhist_of(int model) { return nsv_of(model) + (rp::is_ik(model) ? 1 : 0); }
// plan-aware variants (RP_JIT: the description carries what the template id implies for the compiled fields)
inline int nsv_of(const rp_desc& d) { return d.model == RP_JIT ? d.jit_nsv : nsv_of(d.model); }
inline bool spiking(const rp_desc& d) { return d.model == RP_JIT ? d.jit_spiking != 0 : spiking(d.model); }
inline int nhist_of(const rp_desc& d) { return d.model == RP_JIT ? d.jit_nsv + 1 : nhist_of(d.model); }      // JIT: + recurrent drive
// plane of the state that the recurrent weights project (-1: an expression of the state kept in plan->src)
inline int src_plane_of(const rp_desc& d) { return d.model == RP_JIT ? d.jit_src_plane : (spiking(d.model) ? 1 : -1); }
inline int round_up(int x, int m) { return (x + m - 1) / m * m; }

}  // namespace

struct rp_plan {
    rp_desc d;
    int nsv;
    int sm_count;
    long long launches;
    size_t ws_bytes;
    bool use_tc;           // tcgen05 split-3 contractions (tf32 or binary16 operands, see tc.f16)
    bool per_trial = false;// some template parameter has one value per trial (parameter sweep)
    // fp32 workspaces
    float* Wk = nullptr;     // [N][ldw]   k_i * W
    float* WkT = nullptr;    // [N][ldw]
    int ldw = 0;
    float* u = nullptr;      // [B][ldu]   GEMM output (recurrent drive / Z)
    int ldu = 0;
    float* g = nullptr;      // [B][N]
    float* src = nullptr;    // [B][N]     rate models: act(v)
    float* adj = nullptr;    // [nsv][B][N]
    float* win_acc = nullptr;// [B][N]
    float* pp = nullptr;     // [2][nsv][B][N] ping-pong state when no history is kept
    float2* mf = nullptr;    // iku: [B] per-trial {mean v, mean spike} of the state being stepped
    float2* asum = nullptr;  // iku: [B] per-trial means of the recovery-variable adjoint terms
    float* dWraw = nullptr;  // [N][ldw]
    float* dwout_part = nullptr;   // fused reverse kernel: [B/32][k][N] partial sums of dW_out
    // RP_JIT: module with the run-time compiled kernels of the template
    CUmodule jit_mod = nullptr;
    CUfunction jit_init_src = nullptr, jit_fwd = nullptr, jit_adj = nullptr, jit_fwd_rows = nullptr;
    // optional per-stage timing (rp_plan_stage_timing): a CUDA event on the launching stream before every launch of the step loops
    bool timing = false;
    std::vector<cudaEvent_t> ev_pool;
    std::vector<int> ev_cat;
    size_t ev_used = 0;
    float* wg_g = nullptr;   // few-trial FFMA path: g_t and r_t of several steps, [chunk*B][N] each, so that the
    float* wg_src = nullptr; // read-modify-write of dW happens once per chunk (rank chunk*B update) instead of every step
    int wg_chunk = 0;
    // persistent few-trial path (rp_persistent.cuh)
    bool persistent = false;
    int ps_rows = 0, ps_grid = 0, ps_npad = 0, ps_nrb = 0, ps_bl = 0, ps_ntb = 0;     // grid = ps_nrb row blocks x ps_ntb trial blocks of ps_bl trials
    float* ps_part = nullptr; size_t ps_part_floats = 0;      // per-row-block partial records (ordered reduction of readout / neuron means)
    int ps_fwd_wres = 0, ps_bwd_wres = 0, ps_bwd_dwres = 0;
    size_t ps_fwd_smem = 0, ps_bwd_smem = 0;
    float* ps_vec = nullptr;        // [2][B][Npad] x {value, tag}: flag-in-data exchange buffer (r_t forward, g_t backward)
    unsigned int* ps_bar = nullptr;
    // 3xTF32 operand buffers (leading dims padded to the TMA box)
    rp::TcWorkspace tc;
};

namespace {

// stage categories of rp_plan_stage_times
enum { ST_FWD = 0, ST_DGRAD = 1, ST_WGRAD = 2, ST_ADJ = 3, ST_OTHER = 4, ST_COUNT = RP_NUM_STAGES };
// everything enqueued on `st` from here to the next mark is attributed to `cat`
inline void stage_mark(rp_plan* p, int cat, cudaStream_t st) {
    if (!p->timing) return;
    if (p->ev_used == p->ev_pool.size()) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) { p->timing = false; return; }
        p->ev_pool.push_back(e); p->ev_cat.push_back(ST_OTHER);
    }
    p->ev_cat[p->ev_used] = cat;
    cudaEventRecord(p->ev_pool[p->ev_used++], st);
}

// ---- driver API through the runtime's entry-point query (the library must load on a box without libcuda) -------------------
typedef CUresult (*pfn_cuModuleLoadData)(CUmodule*, const void*);
typedef CUresult (*pfn_cuModuleGetFunction)(CUfunction*, CUmodule, const char*);
typedef CUresult (*pfn_cuModuleUnload)(CUmodule);
typedef CUresult (*pfn_cuLaunchKernel)(CUfunction, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, unsigned, CUstream, void**, void**);
template <typename F> F drv_entry(const char* name) {
    void* ptr = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint(name, &ptr, cudaEnableDefault, &q) != cudaSuccess || q != cudaDriverEntryPointSuccess) return nullptr;
    return reinterpret_cast<F>(ptr);
}
// launch one of the plan's run-time compiled kernels (single by-value argument record)
int jit_launch(CUfunction fn, int grid, int block, void* arg_record, cudaStream_t st) {
    static pfn_cuLaunchKernel launch = drv_entry<pfn_cuLaunchKernel>("cuLaunchKernel");
    if (!launch || !fn) return fail("run-time compiled kernel not available (rp_plan_set_jit_module not called?)");
    void* params[] = {arg_record};
    const CUresult r = launch(fn, (unsigned)grid, 1, 1, (unsigned)block, 1, 1, 0, reinterpret_cast<CUstream>(st), params, nullptr);
    if (r != CUDA_SUCCESS) return fail("cuLaunchKernel of a run-time compiled kernel failed with CUresult %d", (int)r);
    return 0;
}
struct JitInitSrcArgs { int N, B; const float* y; rp::ModelParams mp; float* src; int ld; };

int plan_alloc(rp_plan* p, float** ptr, size_t n_floats) {
    RP_CUDA(cudaMalloc(reinterpret_cast<void**>(ptr), n_floats * sizeof(float)));
    if (const char* poison = getenv("RP_POISON_ALLOC")) RP_CUDA(cudaMemset(*ptr, atoi(poison), n_floats * sizeof(float)));      // debug: uninitialised reads show up
    p->ws_bytes += n_floats * sizeof(float);
    return 0;
}

rp::ModelParams make_params(const rp_plan* p, const float* const* params) {
    rp::ModelParams mp;
    for (int q = 0; q < RP_NUM_PARAMS; ++q) {
        const int mode = p->d.param_per_neuron[q];      // 0 shared, 1 [n], 2 [B], 3 [B][n]
        mp.p[q] = params[q];
        mp.stride[q] = (mode & 1) ? 1 : 0;
        mp.bstride[q] = mode == 2 ? 1 : (mode == 3 ? p->d.n : 0);
    }
    return mp;
}

int check_params(const rp_plan* p, const float* const* params) {
    if (p->d.model == RP_JIT) return params[RP_P_K] ? 0 : fail("RP_JIT: parameter slot RP_P_K (the constant coupling factor 1) is NULL");
    if (rp::is_ik(p->d.model)) {
        static const int need_ik[] = {RP_P_C, RP_P_K, RP_P_VR, RP_P_VTH, RP_P_ETA, RP_P_G, RP_P_ER, RP_P_B, RP_P_TAU_U, RP_P_KAPPA, RP_P_TAU_S};
        for (int q : need_ik) if (!params[q]) return fail("ik_op parameter slot %d is NULL", q);
        if (p->d.model == RP_IK_BIEXP && !params[RP_P_TAU_X]) return fail("ik_biexp_op parameter tau_r (slot RP_P_TAU_X) is NULL");
        return 0;
    }
    static const int need_li[] = {RP_P_TAU, RP_P_K, RP_P_ETA};
    for (int q : need_li) if (!params[q]) return fail("parameter slot %d (tau/k/eta) is NULL", q);
    if (spiking(p->d.model) && !params[RP_P_TAU_S]) return fail("parameter tau_s is NULL");
    if (p->d.model == RP_QIF_SFA && (!params[RP_P_TAU_X] || !params[RP_P_ALPHA])) return fail("parameters tau_x/alpha are NULL");
    if (p->d.model == RP_LI_SIGMOID && (!params[RP_P_RMAX] || !params[RP_P_SIG_S] || !params[RP_P_V0]))
        return fail("parameters r_max/s/v0 are NULL");
    return 0;
}

inline int ew_grid(const rp_plan* p, size_t total) {
    size_t blocks = (total + 255) / 256;
    size_t cap = (size_t)p->sm_count * 8;
    return (int)std::max<size_t>(1, std::min(blocks, cap));
}

// ---- contraction dispatch:  C[q*ldc+p] (+)= sum_k Aop(p,k) Bop(q,k) ------------------------------------
int gemm_fp32(rp_plan* plan, bool kmajor, int P, int Q, int K, const float* A, int lda, const float* B, int ldb,
              float* C, int ldc, int accumulate, cudaStream_t st, long long* launches) {
    (void)plan;
    const cudaError_t e = rp::gemm_fp32_launch(kmajor, P, Q, K, A, lda, B, ldb, C, ldc, accumulate, st, launches);      // rp_fp32_paths.cu
    if (e != cudaSuccess) return fail("fp32 contraction launch failed: %s", cudaGetErrorString(e));
    return 0;
}

// ---- model dispatch helpers -----------------------------------------------------------------------------
#define RP_DISPATCH_MODEL(model, ...)                                  \
    switch (model) {                                                    \
        case RP_LI_TANH:    { constexpr int M_ = RP_LI_TANH;    __VA_ARGS__; } break; \
        case RP_LI_SIGMOID: { constexpr int M_ = RP_LI_SIGMOID; __VA_ARGS__; } break; \
        case RP_QIF:        { constexpr int M_ = RP_QIF;        __VA_ARGS__; } break; \
        case RP_QIF_SFA:    { constexpr int M_ = RP_QIF_SFA;    __VA_ARGS__; } break; \
        case RP_LIF:        { constexpr int M_ = RP_LIF;        __VA_ARGS__; } break; \
        case RP_IK:         { constexpr int M_ = RP_IK;         __VA_ARGS__; } break; \
        case RP_IKU:        { constexpr int M_ = RP_IKU;        __VA_ARGS__; } break; \
        case RP_IK_BIEXP:   { constexpr int M_ = RP_IK_BIEXP;   __VA_ARGS__; } break; \
        default: return fail("unknown model id %d", model);             \
    }

struct Window { int j; int first; int close; int len; };

// record windows of Network.run (network.py:590-597): records at t >= cutoff with t % S == 0, each holding the
// mean of the outputs since the previous record; a trailing window that never reaches a record step is dropped.
Window window_of(int t, int T, int S, int cutoff) {
    Window w{-1, 0, 0, 0};
    if (t < cutoff) return w;
    const int r0 = ((cutoff + S - 1) / S) * S;
    int j, start, rec;
    if (t <= r0) { j = 0; start = cutoff; rec = r0; }
    else { j = (t - r0 + S - 1) / S; rec = r0 + j * S; start = rec - S + 1; }
    if (rec >= T) return w;
    w.j = j; w.first = (t == start); w.close = (t == rec); w.len = rec - start + 1;
    return w;
}

// launch one of the persistent few-trial kernels (defined and compiled in rp_fp32_paths.cu) and translate its status
template <typename F, typename A>
int ps_launch_checked(F launcher, int model, const A* pa, int grid, size_t smem, cudaStream_t st, const char* what) {
    cudaError_t e = cudaSuccess;
    int occ = 0, sms = 0;
    switch (launcher(model, pa, grid, smem, st, &e, &occ, &sms)) {
        case 0: return 0;
        case 1: return fail("%s: persistent kernel launch failed: %s", what, cudaGetErrorString(e));
        case 2: return fail("%s: persistent grid of %d CTAs is not co-resident (occupancy %d x %d SMs)", what, grid, occ, sms);
        default: return fail("%s: no persistent kernel for model id %d", what, model);
    }
}

// geometry of the persistent few-trial kernels; leaves p->persistent false when the shape does not fit.
// Two-dimensional: n_tb trial blocks of BL <= PS_MAX_B trials x n_rb row blocks.  Among the feasible splits the one with the
// fewest trials per CTA wins (smallest per-step gather), i.e. the largest n_tb whose rows of kW still fit in shared memory.
// Decomposition of the persistent few-trial kernels for an n x batch problem, or false when the owned rows of kW do not fit shared memory
struct PersistShape { int rows, bl, ntb, nrb, npad, bwd_wres, bwd_dwres; size_t fwd_smem, bwd_smem; };
bool persistent_shape(int N, int B, int ldw, const cudaDeviceProp& prop, PersistShape* o) {
    const int sms = prop.multiProcessorCount;
    const int npad = round_up(N, 4);
    const size_t budget = std::min<size_t>(prop.sharedMemPerBlockOptin, 200 * 1024);
    for (int bl = 1; bl <= std::min(B, rp::PS_MAX_B); ++bl) {
        const int ntb = (B + bl - 1) / bl;
        if (ntb > sms) continue;
        const int nrb_max = sms / ntb;
        int rows = (N + nrb_max - 1) / nrb_max;
        rows = std::max(8, round_up(rows, 8));
        if (rows > rp::PS_MAX_ROWS || rows * bl > rp::PS_THREADS) continue;
        const size_t wbytes = (size_t)rows * ldw * sizeof(float);
        const size_t base_f = rp::ps_fwd_base_floats(bl, npad) * sizeof(float);
        const size_t base_b = ((size_t)bl * npad + 2 * rp::PS_MAX_ROWS * rp::PS_MAX_B) * sizeof(float);
        // Only worth it while the owned rows of kW stay in shared memory: streaming them from L2 with one warp per row is
        // slower than the per-step launch sequence (measured: N=4096, B=1: 156 vs 116 us/step).
        if (base_f + wbytes > budget || base_b > budget) continue;
        o->rows = rows; o->bl = bl; o->ntb = ntb; o->nrb = (N + rows - 1) / rows; o->npad = npad;
        o->fwd_smem = base_f + wbytes;
        o->bwd_dwres = (base_b + 2 * wbytes <= budget) ? 1 : 0;
        o->bwd_wres = (o->bwd_dwres || base_b + wbytes <= budget) ? 1 : 0;
        o->bwd_smem = base_b + (o->bwd_wres ? wbytes : 0) + (o->bwd_dwres ? wbytes : 0);
        return true;
    }
    return false;
}

int persistent_setup(rp_plan* p, const cudaDeviceProp& prop) {
    const int B = p->d.batch;
    int coop = 0, dev = 0;
    RP_CUDA(cudaGetDevice(&dev));
    RP_CUDA(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev));
    if (!coop) return 0;
    PersistShape ps;
    if (!persistent_shape(p->d.n, B, p->ldw, prop, &ps)) return 0;
    p->ps_rows = ps.rows; p->ps_bl = ps.bl; p->ps_ntb = ps.ntb; p->ps_nrb = ps.nrb;
    p->ps_grid = ps.nrb * ps.ntb;
    p->ps_npad = ps.npad;
    p->ps_fwd_wres = 1;
    p->ps_fwd_smem = ps.fwd_smem;
    p->ps_bwd_dwres = ps.bwd_dwres; p->ps_bwd_wres = ps.bwd_wres; p->ps_bwd_smem = ps.bwd_smem;
    if (plan_alloc(p, &p->ps_vec, 4 * (size_t)B * p->ps_npad)) return 1;
    RP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->ps_bar), sizeof(unsigned int)));
    p->persistent = true;
    return 0;
}

int persistent_forward(rp_plan* p, const rp_fwd_args* a, const rp::ModelParams& mp, cudaStream_t st) {
    const rp_desc& d = p->d;
    const int N = d.n, B = d.batch, nsv = p->nsv;
    const size_t plane = (size_t)B * N, slot = (size_t)nsv * plane;
    const int T_tot = a->T_total > 0 ? a->T_total : a->T;
    const int n_rec = rp_num_records(T_tot, a->sampling_steps, a->cutoff);
    if (a->history) RP_CUDA(cudaMemcpyAsync(a->history, a->y0, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    if (a->T == 0) { RP_CUDA(cudaMemcpyAsync(a->yT, a->y0, slot * sizeof(float), cudaMemcpyDeviceToDevice, st)); return 0; }
    RP_CUDA(cudaMemsetAsync(p->ps_vec, 0, 4 * (size_t)B * p->ps_npad * sizeof(float), st));   // clear stale step tags
    // Readout / neuron-mean records: every row block leaves one partial per record and k_sum_row_blocks adds them in a fixed order
    // (bit-reproducible); only when that buffer would exceed 256 MiB do the kernels fall back to atomics on the zeroed records.
    const bool readout = a->out_rec && d.out_mode == RP_OUT_READOUT && n_rec > 0;
    int n_red = 0;
    for (int r = 0; r < a->n_rec_vars; ++r) if (a->rec_reduce[r]) ++n_red;
    const size_t part_out = readout ? (size_t)n_rec * p->ps_nrb * B * d.n_out : 0;
    const size_t part_one = (size_t)n_rec * p->ps_nrb * B;
    const size_t part_need = part_out + (size_t)n_red * part_one;
    const bool ordered = part_need > 0 && part_need * sizeof(float) <= ((size_t)256 << 20);
    if (ordered && part_need > p->ps_part_floats) {
        if (p->ps_part) { RP_CUDA(cudaFree(p->ps_part)); p->ps_part = nullptr; p->ps_part_floats = 0; }
        RP_CUDA(cudaMalloc(reinterpret_cast<void**>(&p->ps_part), part_need * sizeof(float)));
        p->ps_part_floats = part_need;
    }
    if (!ordered && a->t_offset == 0) {      // accumulated with atomics: clear once, when the first segment of a run is integrated
        if (readout) RP_CUDA(cudaMemsetAsync(a->out_rec, 0, (size_t)n_rec * B * d.n_out * sizeof(float), st));
        for (int r = 0; r < a->n_rec_vars; ++r)
            if (a->rec_reduce[r] && n_rec > 0) RP_CUDA(cudaMemsetAsync(a->rec_buf[r], 0, (size_t)n_rec * B * sizeof(float), st));
    }
    rp::PersistFwdArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.N = N; pa.B = B; pa.T = a->T; pa.m = d.n_in; pa.k = d.n_out; pa.in_mode = d.in_mode; pa.in_target = d.in_target;
    pa.out_mode = d.out_mode; pa.out_var = d.out_var; pa.S = a->sampling_steps; pa.cutoff = a->cutoff;
    pa.t_offset = a->t_offset; pa.T_total = T_tot;
    pa.dt = d.dt; pa.theta = d.theta; pa.v_reset = d.v_reset;
    pa.Wk = p->Wk; pa.ldw = p->ldw; pa.rows_per_cta = p->ps_rows; pa.w_resident = p->ps_fwd_wres;
    pa.n_rb = p->ps_nrb; pa.BL = p->ps_bl;
    if (ordered) {
        float* cur = p->ps_part;
        if (readout) { pa.out_part = cur; cur += part_out; }
        for (int r = 0; r < a->n_rec_vars; ++r) if (a->rec_reduce[r]) { pa.rec_part[r] = cur; cur += part_one; }
    }
    pa.x = a->x; pa.W_in = a->W_in; pa.W_out = a->W_out; pa.mp = mp; pa.y0 = a->y0; pa.yT = a->yT; pa.history = a->history;
    pa.srcbuf = reinterpret_cast<uint2*>(p->ps_vec); pa.Npad = p->ps_npad; pa.out_rec = a->out_rec; pa.n_rec_vars = a->n_rec_vars;
    for (int r = 0; r < a->n_rec_vars; ++r) { pa.rec_var[r] = a->rec_var[r]; pa.rec_reduce[r] = a->rec_reduce[r]; pa.rec_buf[r] = a->rec_buf[r]; }
    pa.rec_post = spiking(d.model) ? 0 : 1;
    pa.barrier = p->ps_bar;
    if (ps_launch_checked(rp::ps_launch_fwd, d.model, &pa, p->ps_grid, p->ps_fwd_smem, st, "rp_forward")) return 1;
    ++p->launches;
    if (ordered) {
        // records that closed inside this segment: [j0, j1)
        const int j0 = rp_num_records(a->t_offset, a->sampling_steps, a->cutoff);
        const int j1 = rp_num_records(std::min(a->t_offset + a->T, T_tot), a->sampling_steps, a->cutoff);
        const int nj = std::min(j1, n_rec) - j0;
        if (nj > 0) {
            if (readout) {
                const int width = B * d.n_out;
                rp::k_sum_row_blocks<<<ew_grid(p, (size_t)nj * width), 256, 0, st>>>(pa.out_part + (size_t)j0 * p->ps_nrb * width, nj, p->ps_nrb, width, 1.0f,
                                                                                   a->out_rec + (size_t)j0 * width);
                ++p->launches;
            }
            for (int r = 0; r < a->n_rec_vars; ++r) {
                if (!a->rec_reduce[r]) continue;
                rp::k_sum_row_blocks<<<ew_grid(p, (size_t)nj * B), 256, 0, st>>>(pa.rec_part[r] + (size_t)j0 * p->ps_nrb * B, nj, p->ps_nrb, B, 1.0f / (float)N,
                                                                              a->rec_buf[r] + (size_t)j0 * B);
                ++p->launches;
            }
            RP_LAUNCH_CHECK();
        }
    }
    return 0;
}

int persistent_backward(rp_plan* p, const rp_bwd_args* a, const rp::ModelParams& mp, bool need_dW, cudaStream_t st) {
    const rp_desc& d = p->d;
    const int N = d.n, B = d.batch;
    const int fold = rp::fold_slot(d.model);
    const int kstride = d.param_per_neuron[fold] ? 1 : 0;
    if (a->T == 0) {
        const size_t slot = (size_t)p->nsv * B * N;
        if (a->g_y0) {
            if (a->g_yT) RP_CUDA(cudaMemcpyAsync(a->g_y0, a->g_yT, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
            else RP_CUDA(cudaMemsetAsync(a->g_y0, 0, slot * sizeof(float), st));
        }
        if (a->dW) RP_CUDA(cudaMemsetAsync(a->dW, 0, (size_t)N * N * sizeof(float), st));
        return 0;
    }
    RP_CUDA(cudaMemsetAsync(p->ps_vec, 0, 4 * (size_t)B * p->ps_npad * sizeof(float), st));   // clear stale step tags
    if (need_dW && !p->ps_bwd_dwres) RP_CUDA(cudaMemsetAsync(p->dWraw, 0, (size_t)p->ps_ntb * N * p->ldw * sizeof(float), st));
    rp::PersistBwdArgs pa;
    memset(&pa, 0, sizeof(pa));
    pa.N = N; pa.B = B; pa.T = a->T; pa.m = d.n_in; pa.k = d.n_out; pa.in_mode = d.in_mode; pa.in_target = d.in_target;
    pa.out_mode = d.out_mode; pa.out_var = d.out_var; pa.S = a->sampling_steps; pa.cutoff = a->cutoff; pa.truncate = a->truncate_steps;
    pa.t_offset = a->t_offset; pa.T_total = a->T_total > 0 ? a->T_total : a->T;
    pa.dt = d.dt; pa.theta = d.theta; pa.slope = d.slope;
    pa.WkT = p->WkT; pa.ldw = p->ldw; pa.rows_per_cta = p->ps_rows; pa.w_resident = p->ps_bwd_wres; pa.dw_resident = p->ps_bwd_dwres;
    pa.n_rb = p->ps_nrb; pa.BL = p->ps_bl;
    pa.need_dW = need_dW ? 1 : 0;
    pa.x = a->x; pa.W_in = a->W_in; pa.W_out = a->W_out; pa.mp = mp; pa.history = a->history; pa.g_out_rec = a->g_out_rec; pa.g_yT = a->g_yT;
    pa.gbuf = reinterpret_cast<uint2*>(p->ps_vec); pa.Npad = p->ps_npad; pa.dWrawT = p->dWraw;
    for (int q = 0; q < RP_NUM_PARAMS; ++q) pa.dparams[q] = (q == fold) ? nullptr : a->dparams[q];
    pa.dW_in = a->dW_in; pa.dW_out = a->dW_out; pa.g_y0 = a->g_y0; pa.g_x = a->g_x; pa.barrier = p->ps_bar;
    if (ps_launch_checked(rp::ps_launch_bwd, d.model, &pa, p->ps_grid, p->ps_bwd_smem, st, "rp_backward")) return 1;
    ++p->launches;
    if (need_dW) {
        dim3 grid((N + 31) / 32, (N + 31) / 32);
        rp::k_finish_wgrad_T<<<grid, 256, 0, st>>>(N, p->dWraw, p->ldw, a->W, a->params[fold], kstride, a->dW, a->dparams[fold], p->ps_ntb);
        ++p->launches;
        RP_LAUNCH_CHECK();
    }
    return 0;
}

}  // namespace

extern "C" {

int rp_abi_version(void) { return RP_ABI_VERSION; }
const char* rp_last_error(void) { return g_err; }
int rp_num_state_vars(int model) { return nsv_of(model); }
int rp_num_history_planes(int model) { return nsv_of(model) < 0 ? -1 : nhist_of(model); }

int rp_num_records(int T, int S, int cutoff) {
    if (S <= 0 || T <= 0) return 0;
    const int r0 = ((std::max(cutoff, 0) + S - 1) / S) * S;
    if (r0 > T - 1) return 0;
    return (T - 1 - r0) / S + 1;
}

int rp_plan_path(const rp_desc* d) {
    if (!d) { fail("rp_plan_path: null argument"); return -1; }
    if (d->precision == RP_PREC_3XTF32 || d->precision == RP_PREC_3XF16) return rp::tc_supported(d->n, d->batch) ? 2 : -1;
    if (rp::is_mean_field(d->model) || d->model == RP_JIT || getenv("RP_NO_PERSISTENT")) return 0;
    int dev = 0, coop = 0;
    cudaDeviceProp prop;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&prop, dev) != cudaSuccess) { fail("rp_plan_path: no CUDA device"); return -1; }
    cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, dev);
    PersistShape ps;
    return (coop && persistent_shape(d->n, d->batch, round_up(d->n, 4), prop, &ps)) ? 1 : 0;
}

int rp_plan_create(const rp_desc* d, rp_plan** out) {
    if (!d || !out) return fail("rp_plan_create: null argument");
    const int nsv = nsv_of(*d);
    if (nsv < 0 || (d->model == RP_JIT && (nsv < 1 || nsv > RP_MAX_SV))) return fail("rp_plan_create: unknown model %d / bad state-variable count", d->model);
    if (d->model == RP_JIT) {
        if (d->precision != RP_PREC_FP32) return fail("rp_plan_create: run-time compiled fields run on the per-step fp32 path (precision must be RP_PREC_FP32)");
        if (d->jit_spiking < 0 || d->jit_spiking > nsv) return fail("rp_plan_create: RP_JIT with %d spike variables but %d state variables", d->jit_spiking, nsv);
        if (d->jit_src_plane >= nsv || d->out_var >= 3 || d->out_var >= nsv) return fail("rp_plan_create: RP_JIT source / output plane out of range (output planes 0..2)");
    }
    if (d->n <= 0 || d->batch <= 0) return fail("rp_plan_create: n and batch must be positive");
    if (d->in_mode == RP_IN_PROJ && (d->n_in <= 0 || d->n_in > RP_MAX_IN)) return fail("rp_plan_create: n_in must be in [1,%d] for RP_IN_PROJ", RP_MAX_IN);
    if (d->out_mode == RP_OUT_READOUT && (d->n_out <= 0 || d->n_out > RP_MAX_OUT)) return fail("rp_plan_create: n_out must be in [1,%d] for RP_OUT_READOUT", RP_MAX_OUT);
    if (d->out_var < 0 || d->out_var > RP_VAR_R) return fail("rp_plan_create: bad out_var");
    if (d->out_var == RP_VAR_R && (spiking(*d) || d->model == RP_JIT)) return fail("rp_plan_create: RP_VAR_R only exists on rate models");
    if (d->out_var != RP_VAR_R && d->out_var >= nsv) return fail("rp_plan_create: out_var %d not a state variable of model %d", d->out_var, d->model);
    if (d->in_target == 1 && d->model != RP_LIF) return fail("rp_plan_create: in_target=1 (s_ext) only exists on the lif template");
    if (!(d->dt > 0.f)) return fail("rp_plan_create: dt must be positive");
    for (int q = 0; q < RP_NUM_PARAMS; ++q)
        if (d->param_per_neuron[q] < 0 || d->param_per_neuron[q] > 3) return fail("rp_plan_create: param_per_neuron[%d] must be 0..3", q);
    if (d->param_per_neuron[rp::fold_slot(d->model)] > 1)
        return fail("rp_plan_create: the coupling constant (slot %d) is folded into the shared weights and cannot differ between trials", rp::fold_slot(d->model));
    int dev = 0, count = 0;
    if (cudaGetDeviceCount(&count) != cudaSuccess || count == 0) return fail("rp_plan_create: no CUDA device available (the engine has no CPU fallback)");
    RP_CUDA(cudaGetDevice(&dev));
    cudaDeviceProp prop;
    RP_CUDA(cudaGetDeviceProperties(&prop, dev));
    if (prop.major != 10) return fail("rp_plan_create: device is sm_%d%d, this library is built for sm_100a only", prop.major, prop.minor);

    rp_plan* p = new (std::nothrow) rp_plan();
    if (!p) return fail("rp_plan_create: out of host memory");
    p->d = *d;
    p->nsv = nsv;
    p->sm_count = prop.multiProcessorCount;
    p->launches = 0;
    p->ws_bytes = 0;
    const int N = d->n, B = d->batch;
    p->use_tc = (d->precision == RP_PREC_3XTF32 || d->precision == RP_PREC_3XF16);
    // binary16 operands need a bound on the source variable one step ahead; the lif s_ext input adds an unbounded term to s'
    const bool f16 = d->precision == RP_PREC_3XF16 && !(d->model == RP_LIF && d->in_target == 1) && !getenv("RP_NO_F16");
    for (int q = 0; q < RP_NUM_PARAMS; ++q) if (d->param_per_neuron[q] > 1) p->per_trial = true;
    if (p->use_tc && !rp::tc_supported(N, B)) {
        delete p;
        return fail("rp_plan_create: RP_PREC_3XTF32 / RP_PREC_3XF16 need n %% 128 == 0 and batch %% 128 == 0 (got n=%d batch=%d)", N, B);
    }
    p->ldw = round_up(N, 4);
    p->ldu = round_up(N, 4);
    const size_t plane = (size_t)B * N;
    int rc = 0;
    rc |= plan_alloc(p, &p->u, (size_t)B * p->ldu);
    rc |= plan_alloc(p, &p->adj, (size_t)nsv * plane);
    rc |= plan_alloc(p, &p->win_acc, plane);
    rc |= plan_alloc(p, &p->pp, 2 * (size_t)nsv * plane);
    if (!p->use_tc) {
        rc |= plan_alloc(p, &p->Wk, (size_t)N * p->ldw);
        rc |= plan_alloc(p, &p->WkT, (size_t)N * p->ldw);
        rc |= plan_alloc(p, &p->g, plane);
        if (src_plane_of(*d) < 0) rc |= plan_alloc(p, &p->src, (d->model == RP_JIT ? 2 : 1) * plane);      // RP_JIT: double buffered (row kernel)
    } else {
        size_t bytes = 0;
        rc |= rp::tc_workspace_create(&p->tc, N, B, f16, !spiking(d->model), &bytes);
        if (rc) fail("rp_plan_create: tensor-core workspace allocation failed: %s", rp::tc_last_error());
        p->ws_bytes += bytes;
    }
    if (rc) { rp_plan_destroy(p); return 1; }
    if (rp::is_mean_field(d->model)) {
        if (cudaMalloc(reinterpret_cast<void**>(&p->mf), (size_t)B * sizeof(float2)) != cudaSuccess ||
            cudaMalloc(reinterpret_cast<void**>(&p->asum), (size_t)B * sizeof(float2)) != cudaSuccess) {
            rp_plan_destroy(p);
            return fail("rp_plan_create: cudaMalloc of the mean-field buffers failed");
        }
        p->ws_bytes += 2 * (size_t)B * sizeof(float2);
    }
    // (iku_op needs a per-step reduction over all neurons of a trial: per-step launch sequences only)
    if (!p->use_tc && !rp::is_mean_field(d->model) && d->model != RP_JIT && !getenv("RP_NO_PERSISTENT")) {
        if (persistent_setup(p, prop)) { rp_plan_destroy(p); return 1; }
    }
    *out = p;
    return 0;
}

void rp_plan_destroy(rp_plan* p) {
    if (!p) return;
    float* bufs[] = {p->Wk, p->WkT, p->u, p->g, p->src, p->adj, p->win_acc, p->pp, p->dWraw, p->ps_vec, p->wg_g, p->wg_src, p->dwout_part};
    for (float* b : bufs) if (b) cudaFree(b);
    if (p->ps_bar) cudaFree(p->ps_bar);
    if (p->ps_part) cudaFree(p->ps_part);
    if (p->jit_mod) { static pfn_cuModuleUnload unload = drv_entry<pfn_cuModuleUnload>("cuModuleUnload"); if (unload) unload(p->jit_mod); }
    if (p->mf) cudaFree(p->mf);
    if (p->asum) cudaFree(p->asum);
    for (cudaEvent_t e : p->ev_pool) cudaEventDestroy(e);
    rp::tc_workspace_destroy(&p->tc);
    delete p;
}

long long rp_plan_workspace_bytes(const rp_plan* p) { return p ? (long long)p->ws_bytes : 0; }
long long rp_plan_launch_count(const rp_plan* p) { return p ? p->launches : 0; }

int rp_plan_set_jit_module(rp_plan* p, const void* image, long long nbytes) {
    if (!p || !image || nbytes <= 0) return fail("rp_plan_set_jit_module: bad argument");
    if (p->d.model != RP_JIT) return fail("rp_plan_set_jit_module: the plan was not created with model RP_JIT");
    static pfn_cuModuleLoadData load = drv_entry<pfn_cuModuleLoadData>("cuModuleLoadData");
    static pfn_cuModuleGetFunction getf = drv_entry<pfn_cuModuleGetFunction>("cuModuleGetFunction");
    if (!load || !getf) return fail("rp_plan_set_jit_module: CUDA driver entry points not available");
    RP_CUDA(cudaFree(0));                                   // make sure the primary context is current
    CUresult r = load(&p->jit_mod, image);
    if (r != CUDA_SUCCESS) return fail("rp_plan_set_jit_module: cuModuleLoadData failed with CUresult %d", (int)r);
    if (getf(&p->jit_fwd, p->jit_mod, "rp_jit_fwd_step") != CUDA_SUCCESS || getf(&p->jit_adj, p->jit_mod, "rp_jit_adj_step") != CUDA_SUCCESS ||
        getf(&p->jit_init_src, p->jit_mod, "rp_jit_init_src") != CUDA_SUCCESS)
        return fail("rp_plan_set_jit_module: the image does not define rp_jit_fwd_step / rp_jit_adj_step / rp_jit_init_src");
    if (getf(&p->jit_fwd_rows, p->jit_mod, "rp_jit_fwd_step_rows") != CUDA_SUCCESS) p->jit_fwd_rows = nullptr;      // optional
    return 0;
}

// ------------------------------------------------------------------------------------------------------
int rp_forward(rp_plan* p, const rp_fwd_args* a, void* stream) {
    if (!p || !a) return fail("rp_forward: null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const rp_desc& d = p->d;
    const int N = d.n, B = d.batch, nsv = p->nsv;
    const size_t plane = (size_t)B * N, slot = (size_t)nsv * plane;
    if (a->T < 0 || a->sampling_steps <= 0 || a->cutoff < 0) return fail("rp_forward: bad T/sampling_steps/cutoff");
    if (!a->W || !a->y0 || !a->yT) return fail("rp_forward: W, y0 and yT are required");
    if (d.in_mode != RP_IN_NONE && !a->x && a->T > 0) return fail("rp_forward: x is NULL but in_mode != RP_IN_NONE");
    if (d.in_mode == RP_IN_PROJ && !a->W_in) return fail("rp_forward: W_in is NULL for RP_IN_PROJ");
    if (d.out_mode == RP_OUT_READOUT && a->out_rec && !a->W_out) return fail("rp_forward: W_out is NULL for RP_OUT_READOUT");
    if (a->n_rec_vars < 0 || a->n_rec_vars > RP_MAX_REC) return fail("rp_forward: n_rec_vars out of range");
    for (int r = 0; r < a->n_rec_vars; ++r) {
        if (a->rec_var[r] < 0 || a->rec_var[r] >= nsv) return fail("rp_forward: recorded variable %d is not a state variable", a->rec_var[r]);
        if (!a->rec_buf[r]) return fail("rp_forward: rec_buf[%d] is NULL", r);
    }
    if (check_params(p, a->params)) return 1;
    const rp::ModelParams mp = make_params(p, a->params);
    const int fold = rp::fold_slot(d.model);
    const int kstride = d.param_per_neuron[fold] ? 1 : 0;
    stage_mark(p, ST_OTHER, st);

    // weights: fold k into W once per call
    {
        dim3 grid((N + 31) / 32, (N + 31) / 32);
        if (!p->use_tc) rp::k_prepare_weights<<<grid, 256, 0, st>>>(N, a->W, a->params[fold], kstride, p->Wk, nullptr, p->ldw, nullptr, nullptr, nullptr, nullptr);
        else if (!p->tc.f16) rp::k_prepare_weights<<<grid, 256, 0, st>>>(N, a->W, a->params[fold], kstride, nullptr, nullptr, p->tc.ldk,
                                                                        (float*)p->tc.W_hi, (float*)p->tc.W_lo, nullptr, nullptr);
        else {
            RP_CUDA(cudaMemsetAsync(p->tc.meta + rp::TCM_AMAX_W, 0, 2 * sizeof(float), st));       // AMAX_W, AMAX_WOUT
            rp::k_amax_2d<<<p->sm_count * 4, 256, 0, st>>>(N, N, a->W, (size_t)N, a->params[fold], kstride, p->tc.meta + rp::TCM_AMAX_W);
            // a call that keeps checkpoints is followed by rp_backward on the same weights: split the transpose in the same pass
            const bool with_T = a->history != nullptr;
            rp::k_prepare_weights_f16<<<grid, 256, 0, st>>>(N, a->W, a->params[fold], kstride, p->tc.ldk, p->tc.W_hi, p->tc.W_lo,
                                                            with_T ? p->tc.WT_hi : nullptr, with_T ? p->tc.WT_lo : nullptr, rp::tc_scale_W(&p->tc));
            p->tc.wt_W = with_T ? a->W : nullptr;
            ++p->launches;
        }
        ++p->launches;
        RP_LAUNCH_CHECK();
    }
    const int T_tot = a->T_total > 0 ? a->T_total : a->T;
    if (a->t_offset < 0 || a->t_offset + a->T > T_tot) return fail("rp_forward: bad t_offset/T_total");
    if (p->persistent) return persistent_forward(p, a, mp, st);
    float* base = a->history ? a->history : p->pp;
    const size_t hslot = (size_t)nhist_of(d) * plane;          // checkpoint slot stride (ik, run-time compiled fields: + recurrent-drive plane)
    auto slot_ptr = [&](int t) -> float* { return a->history ? base + (size_t)t * hslot : base + (size_t)(t & 1) * slot; };
    const bool jit = d.model == RP_JIT;
    const int src_plane = src_plane_of(d);
    if (jit && !p->jit_fwd) return fail("rp_forward: RP_JIT plan without a module (call rp_plan_set_jit_module first)");
    RP_CUDA(cudaMemcpyAsync(slot_ptr(0), a->y0, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));

    const bool spk = spiking(d);
    const bool f16 = p->use_tc && p->tc.f16;
    // binary16 operands: bound of the source variable.  Rate models: static (|tanh| <= 1, sigmoid <= max r_max).  Spiking models:
    // |s_{t+1}| <= |s_t| + 1 (one spike adds exactly 1, nodes.py:385), tracked per step in tc.amax_src.
    const rp::ScaleRef sc_rate = rp::tc_scale_srcbound(&p->tc);
    if (f16 && a->T > 0) {
        if (spk) {
            size_t bytes = 0;
            if (rp::tc_workspace_ensure_amax(&p->tc, a->T + 2, &bytes)) return fail("rp_forward: %s", rp::tc_last_error());
            p->ws_bytes += bytes;
            RP_CUDA(cudaMemsetAsync(p->tc.amax_src, 0, (size_t)(a->T + 2) * sizeof(float), st));
        } else if (d.model == RP_LI_TANH) {
            rp::k_fill<<<1, 32, 0, st>>>(p->tc.meta + rp::TCM_SRC_BOUND, 1, 1.0f);
            ++p->launches;
        } else {
            const int mode = d.param_per_neuron[RP_P_RMAX];
            const int count = mode == 0 ? 1 : (mode == 1 ? N : (mode == 2 ? B : B * N));
            RP_CUDA(cudaMemsetAsync(p->tc.meta + rp::TCM_SRC_BOUND, 0, sizeof(float), st));
            rp::k_amax_2d<<<32, 256, 0, st>>>(1, count, a->params[RP_P_RMAX], (size_t)count, nullptr, 0, p->tc.meta + rp::TCM_SRC_BOUND);
            ++p->launches;
        }
        RP_LAUNCH_CHECK();
    }
    if (jit && a->T > 0 && src_plane < 0) {          // source = expression of the state: src_0 from y_0
        JitInitSrcArgs ja{N, B, slot_ptr(0), mp, p->src, N};
        if (jit_launch(p->jit_init_src, ew_grid(p, plane), 256, &ja, st)) return 1;
        ++p->launches;
    }
    if (!jit && a->T > 0 && (!spk || p->use_tc)) {
        if (f16 && spk) {    // first pass: exact maximum of src_0
            RP_DISPATCH_MODEL(d.model, (rp::k_init_src<M_><<<ew_grid(p, plane), 256, 0, st>>>(N, B, slot_ptr(0), mp, nullptr, N, nullptr, nullptr,
                                        p->tc.ldk, 0, rp::no_scale(), p->tc.amax_src)));
            ++p->launches;
        }
        const rp::ScaleRef sc0 = f16 ? (spk ? rp::ScaleRef{p->tc.amax_src, 0.f, rp::CV_HSRC} : sc_rate) : rp::no_scale();
        RP_DISPATCH_MODEL(d.model, (rp::k_init_src<M_><<<ew_grid(p, plane), 256, 0, st>>>(N, B, slot_ptr(0), mp,
                                    (!spk && !p->use_tc) ? p->src : nullptr, N, p->use_tc ? p->tc.src_hi : nullptr,
                                    p->use_tc ? p->tc.src_lo : nullptr, p->tc.ldk, f16 ? 1 : 0, sc0, nullptr)));
        ++p->launches;
        RP_LAUNCH_CHECK();
    }
    const size_t x_stride = d.in_mode == RP_IN_DENSE ? plane : (d.in_mode == RP_IN_PROJ ? (size_t)B * d.n_in : 0);
    const size_t out_stride = d.out_mode == RP_OUT_READOUT ? (size_t)B * d.n_out : plane;
    if (p->use_tc && a->T > 0) {          // the source operand is double buffered by step parity on every tensor-core path (see tc_forward_step)
        size_t bytes = 0;
        if (rp::tc_workspace_ensure_persist(&p->tc, &bytes)) return fail("rp_forward: %s", rp::tc_last_error());
        p->ws_bytes += bytes;
    }

    // tensor-core path: the step (and, for spiking nets read out from s, the readout) is the contraction's epilogue
    const bool fuse_readout = p->use_tc && spk && d.out_var == RP_VAR_S && d.out_mode == RP_OUT_READOUT && a->out_rec != nullptr;
    if (fuse_readout) {
        const size_t roff = (size_t)N * p->tc.ldk * p->tc.esize();
        if (f16) {
            rp::k_amax_2d<<<32, 256, 0, st>>>(d.n_out, N, a->W_out, (size_t)N, nullptr, 0, p->tc.meta + rp::TCM_AMAX_WOUT);
            ++p->launches;
        }
        rp::k_split_matrix<<<64, 256, 0, st>>>(d.n_out, N, a->W_out, N, (char*)p->tc.W_hi + roff, (char*)p->tc.W_lo + roff, p->tc.ldk, d.n_out,
                                               f16 ? 1 : 0, rp::tc_scale_Wout(&p->tc));
        ++p->launches;
        RP_LAUNCH_CHECK();
    }
    // Round 2: the whole horizon of the batched path as ONE persistent cooperative launch (binary16 operands, 256-trial tiles, nothing
    // that needs a per-step side kernel: no recorded state variables, output = fused readout or none, no mean-field template)
    bool persisted = false;
    if (f16 && p->tc.bq_fwd == 256 && a->T > 1 && !rp::is_mean_field(d.model) && a->n_rec_vars == 0 && (fuse_readout || a->out_rec == nullptr) &&
        !getenv("RP_NO_FWD_PERSIST")) {
        size_t bytes = 0;
        if (rp::tc_workspace_ensure_persist(&p->tc, &bytes)) return fail("rp_forward: %s", rp::tc_last_error());
        p->ws_bytes += bytes;
        rp::FwdStepArgs fa;
        memset(&fa, 0, sizeof(fa));
        fa.N = N; fa.B = B; fa.m = d.n_in; fa.in_mode = d.in_mode; fa.in_target = d.in_target;
        fa.dt = d.dt; fa.theta = d.theta; fa.v_reset = d.v_reset; fa.u = p->u; fa.ldu = p->ldu;
        fa.W_in = a->W_in; fa.mp = mp; fa.ld_src = p->tc.ldk; fa.mf = p->mf;
        fa.sc_out = rp::no_scale(); fa.per_trial = p->per_trial ? 1 : 0; fa.no_lean = getenv("RP_NO_FWD_LEAN") ? 1 : 0;
        rp::FwdPersist ps;
        memset(&ps, 0, sizeof(ps));
        ps.T = a->T; ps.t_offset = a->t_offset; ps.T_total = T_tot; ps.S = a->sampling_steps; ps.cutoff = a->cutoff;
        ps.n_row_tiles = N / rp::TC_BP; ps.readout = fuse_readout ? 1 : 0;
        ps.y_hist = a->history; ps.hslot = hslot; ps.y_pp = p->pp; ps.slot = slot;
        ps.x = a->x; ps.x_stride = x_stride; ps.out_rec = a->out_rec; ps.out_stride = out_stride;
        ps.amax_src = spk ? p->tc.amax_src : nullptr; ps.sc_static = sc_rate;
        { const char* sk = getenv("RP_FWD_SKEW_US"); ps.skew_ns = sk ? atoi(sk) * 1000 : 0; }
        ps.nsv = nsv; ps.plane = plane; ps.urec = (rp::is_ik(d.model) && a->history) ? 1 : 0;
        int prc = 0;
        auto fill0 = [&](auto& epi) {
            epi.a = fa; epi.out_rec_j = nullptr; epi.k = d.n_out; epi.win_first = 0; epi.win_close = 0; epi.inv_len = 0.f;
            epi.sA = rp::tc_scale_W(&p->tc); epi.sA_ro = rp::tc_scale_Wout(&p->tc); epi.sB = sc_rate;
        };
        stage_mark(p, ST_FWD, st);
        if (p->per_trial || rp::is_ik(d.model)) {
            RP_DISPATCH_MODEL(d.model, { rp::EpiFwd<M_, true> epi; fill0(epi); prc = rp::tc_forward_persistent<M_, true>(&p->tc, epi, ps, p->sm_count, st); });
        } else {
            RP_DISPATCH_MODEL(d.model, { rp::EpiFwd<M_, false> epi; fill0(epi); prc = rp::tc_forward_persistent<M_, false>(&p->tc, epi, ps, p->sm_count, st); });
        }
        if (prc == 1) return fail("rp_forward: %s", rp::tc_last_error());
        if (prc == 0) { persisted = true; ++p->launches; }
    }
    const int env_no_lean = getenv("RP_NO_FWD_LEAN") ? 1 : 0;           // read once per call, not per step
    const bool jit_rows = jit && p->jit_fwd_rows && B <= 8 && !getenv("RP_JIT_NO_ROWS");
    for (int t = 0; t < (persisted ? 0 : a->T); ++t) {
        float* cur = slot_ptr(t);
        float* nxt = slot_ptr(t + 1);
        const Window w = window_of(a->t_offset + t, T_tot, a->sampling_steps, a->cutoff);
        rp::FwdStepArgs fa;
        fa.N = N; fa.B = B; fa.m = d.n_in; fa.in_mode = d.in_mode; fa.in_target = d.in_target;
        fa.dt = d.dt; fa.theta = d.theta; fa.v_reset = d.v_reset;
        fa.y_cur = cur; fa.y_next = nxt; fa.u = p->u; fa.ldu = p->ldu;
        fa.x_t = a->x ? a->x + (size_t)t * x_stride : nullptr; fa.W_in = a->W_in; fa.mp = mp;
        fa.src_next = (src_plane < 0 && !p->use_tc) ? p->src : nullptr;
        // tensor-core path: src_t is read from buffer t & 1, src_{t+1} is written into the other one
        fa.src_hi = p->use_tc ? ((t & 1) ? p->tc.src_hi : p->tc.src2_hi) : nullptr;
        fa.src_lo = p->use_tc ? ((t & 1) ? p->tc.src_lo : p->tc.src2_lo) : nullptr; fa.ld_src = p->tc.ldk;
        fa.sc_out = rp::no_scale(); fa.amax_out = nullptr;
        rp::ScaleRef sc_in = rp::no_scale();
        if (f16) {
            if (spk) {
                sc_in = t == 0 ? rp::ScaleRef{p->tc.amax_src, 0.f, rp::CV_HSRC} : rp::ScaleRef{p->tc.amax_src + (t - 1), 1.f, rp::CV_HSRC};
                fa.sc_out = rp::ScaleRef{p->tc.amax_src + t, 1.f, rp::CV_HSRC};
                fa.amax_out = p->tc.amax_src + (t + 1);
            } else {
                sc_in = sc_rate; fa.sc_out = sc_rate;
            }
        }
        fa.urec_out = ((rp::is_ik(d.model) || jit) && a->history) ? cur + (size_t)nsv * plane : nullptr;
        fa.mf = p->mf;
        if (rp::is_mean_field(d.model)) {
            rp::k_trial_means<<<B, 256, 0, st>>>(N, cur, d.theta, p->mf);
            ++p->launches;
            RP_LAUNCH_CHECK();
        }
        fa.per_trial = p->per_trial ? 1 : 0;
        stage_mark(p, ST_FWD, st);
        fa.no_lean = env_no_lean;
        if (!p->use_tc) {
            // u[b][i] = sum_j (kW)[i][j] src_t[b][j], then the element-wise step
            const float* srcp = src_plane >= 0 ? cur + (size_t)src_plane * plane : p->src;
            if (jit_rows) {
                // few trials: the generated kernel forms u = W . src itself (one warp per neuron row) -- one launch per step instead of two.
                // src_{t+1} goes to the other half of the double-buffered source (this launch still reads src_t)
                rp::JitRowsArgs ra;
                ra.a = fa; ra.W = p->Wk; ra.ldw = p->ldw; ra.src = srcp;
                if (src_plane < 0) { ra.src = p->src + (size_t)(t & 1) * plane; ra.a.src_next = p->src + (size_t)((t + 1) & 1) * plane; }
                if (jit_launch(p->jit_fwd_rows, (N + 7) / 8, 256, &ra, st)) return 1;
            } else {
                if (gemm_fp32(p, true, N, B, N, p->Wk, p->ldw, srcp, N, p->u, p->ldu, 0, st, &p->launches)) return 1;
                if (jit) { if (jit_launch(p->jit_fwd, ew_grid(p, plane), 256, &fa, st)) return 1; }
                else RP_DISPATCH_MODEL(d.model, (rp::k_fwd_step<M_><<<ew_grid(p, plane), 256, 0, st>>>(fa)));
            }
            ++p->launches;
            RP_LAUNCH_CHECK();
        } else {
            const bool ro = fuse_readout && w.j >= 0;
            auto fill = [&](auto& epi) {
                epi.a = fa;
                epi.out_rec_j = ro ? a->out_rec + (size_t)w.j * out_stride : nullptr;
                epi.k = d.n_out; epi.win_first = w.first; epi.win_close = w.close; epi.inv_len = w.j >= 0 ? 1.0f / (float)w.len : 0.f;
                epi.sA = rp::tc_scale_W(&p->tc); epi.sA_ro = rp::tc_scale_Wout(&p->tc); epi.sB = sc_in;
            };
            if (p->per_trial || rp::is_ik(d.model)) {
                RP_DISPATCH_MODEL(d.model, {
                    rp::EpiFwd<M_, true> epi; fill(epi);
                    if (rp::tc_forward_step<M_, true>(&p->tc, epi, ro, st, t & 1)) return fail("rp_forward: %s", rp::tc_last_error());
                });
            } else {
                RP_DISPATCH_MODEL(d.model, {
                    rp::EpiFwd<M_, false> epi; fill(epi);
                    if (rp::tc_forward_step<M_, false>(&p->tc, epi, ro, st, t & 1)) return fail("rp_forward: %s", rp::tc_last_error());
                });
            }
            ++p->launches;
        }
        const bool want_out = a->out_rec != nullptr && !fuse_readout;
        const bool want_rec = a->n_rec_vars > 0 && w.close;
        if (w.j >= 0 && (want_rec || (want_out && (w.close || w.len > 1)))) {
            rp::ObsArgs oa;
            oa.N = N; oa.B = B; oa.k = d.n_out; oa.out_mode = d.out_mode; oa.out_var = d.out_var; oa.model = d.model;
            oa.y_pre = (jit && d.jit_post_out) ? nxt : cur; oa.y_post = nxt; oa.W_out = a->W_out; oa.mp = mp; oa.win_acc = p->win_acc;
            oa.win_first = w.first; oa.win_close = w.close; oa.inv_len = 1.0f / (float)w.len;
            oa.out_rec_j = want_out ? a->out_rec + (size_t)w.j * out_stride : nullptr;
            oa.skip_out = want_out ? 0 : 1;
            oa.n_rec_vars = a->n_rec_vars;
            for (int r = 0; r < RP_MAX_REC; ++r) {
                oa.rec_var[r] = r < a->n_rec_vars ? a->rec_var[r] : 0;
                oa.rec_reduce[r] = r < a->n_rec_vars ? a->rec_reduce[r] : 0;
                oa.rec_buf_j[r] = r < a->n_rec_vars ? a->rec_buf[r] + (size_t)w.j * (a->rec_reduce[r] ? (size_t)B : plane) : nullptr;
            }
            oa.rec_post = (spk && !(jit && d.jit_post_out)) ? 0 : 1;     // RateNet, MultiSpikeResetNet: y is post-update (nodes.py:169,464); SpikeResetNet: pre-update (nodes.py:387)
            stage_mark(p, ST_OTHER, st);
            if (jit) rp::k_observe<RP_QIF><<<B, 256, 0, st>>>(oa);       // state planes only: any instantiation without an activation will do
            else RP_DISPATCH_MODEL(d.model, (rp::k_observe<M_><<<B, 256, 0, st>>>(oa)));
            ++p->launches;
            RP_LAUNCH_CHECK();
        }
    }
    stage_mark(p, ST_OTHER, st);
    if (f16) { p->tc.fwd_history = a->history; p->tc.fwd_T = a->T; }
    RP_CUDA(cudaMemcpyAsync(a->yT, slot_ptr(a->T), slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

// ------------------------------------------------------------------------------------------------------
int rp_backward(rp_plan* p, const rp_bwd_args* a, void* stream) {
    if (!p || !a) return fail("rp_backward: null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const rp_desc& d = p->d;
    const int N = d.n, B = d.batch, nsv = p->nsv;
    const size_t plane = (size_t)B * N, slot = (size_t)nsv * plane;
    if (a->T < 0 || a->sampling_steps <= 0 || a->cutoff < 0) return fail("rp_backward: bad T/sampling_steps/cutoff");
    if (!a->W || !a->history) return fail("rp_backward: W and history are required");
    if (d.in_mode == RP_IN_PROJ && !a->W_in) return fail("rp_backward: W_in is NULL for RP_IN_PROJ");
    if (d.in_mode != RP_IN_NONE && !a->x && a->T > 0) return fail("rp_backward: x is NULL but in_mode != RP_IN_NONE");
    if (d.out_mode == RP_OUT_READOUT && a->g_out_rec && !a->W_out) return fail("rp_backward: W_out is NULL for RP_OUT_READOUT");
    if (a->dW_in && d.in_mode != RP_IN_PROJ) return fail("rp_backward: dW_in requested without RP_IN_PROJ");
    if (a->dW_out && d.out_mode != RP_OUT_READOUT) return fail("rp_backward: dW_out requested without RP_OUT_READOUT");
    if (a->g_x && d.in_mode != RP_IN_DENSE) return fail("rp_backward: g_x requested without RP_IN_DENSE");
    if (check_params(p, a->params)) return 1;
    for (int q = 0; q < RP_NUM_PARAMS; ++q)
        if (a->dparams[q] && d.param_per_neuron[q] > 1) return fail("rp_backward: gradients of per-trial parameters (slot %d) are not provided", q);
    const rp::ModelParams mp = make_params(p, a->params);
    const int fold = rp::fold_slot(d.model);
    const int kstride = d.param_per_neuron[fold] ? 1 : 0;
    const bool spk = spiking(d);
    const bool jit = d.model == RP_JIT;
    const int src_plane = src_plane_of(d);
    if (jit && !p->jit_adj) return fail("rp_backward: RP_JIT plan without a module (call rp_plan_set_jit_module first)");
    const bool need_dW = a->dW != nullptr || a->dparams[rp::fold_slot(d.model)] != nullptr;
    stage_mark(p, ST_OTHER, st);

    const int wg_slices = p->use_tc ? rp::TC_WGRAD_SPLITS : (p->persistent ? std::max(1, p->ps_ntb) : 1);
    if (need_dW && !p->dWraw) { if (plan_alloc(p, &p->dWraw, (size_t)wg_slices * N * p->ldw)) return 1; }
    if (need_dW && p->use_tc) {
        size_t bytes = 0;
        if (rp::tc_workspace_ensure_wgrad(&p->tc, &bytes)) return fail("rp_backward: %s", rp::tc_last_error());
        p->ws_bytes += bytes;
    }
    {
        dim3 grid((N + 31) / 32, (N + 31) / 32);
        if (!p->use_tc) rp::k_prepare_weights<<<grid, 256, 0, st>>>(N, a->W, a->params[fold], kstride, nullptr, p->WkT, p->ldw, nullptr, nullptr, nullptr, nullptr);
        else if (!p->tc.f16) rp::k_prepare_weights<<<grid, 256, 0, st>>>(N, a->W, a->params[fold], kstride, nullptr, nullptr, p->tc.ldk, nullptr, nullptr,
                                                                        (float*)p->tc.WT_hi, (float*)p->tc.WT_lo);
        else if (p->tc.wt_W == a->W && p->tc.fwd_history == a->history && a->history != nullptr) {
            --p->launches;           // the forward call that wrote these checkpoints already split (kW)^T with the same scale
        } else {
            RP_CUDA(cudaMemsetAsync(p->tc.meta + rp::TCM_AMAX_W, 0, sizeof(float), st));
            rp::k_amax_2d<<<p->sm_count * 4, 256, 0, st>>>(N, N, a->W, (size_t)N, a->params[fold], kstride, p->tc.meta + rp::TCM_AMAX_W);
            rp::k_prepare_weights_f16<<<grid, 256, 0, st>>>(N, a->W, a->params[fold], kstride, p->tc.ldk, nullptr, nullptr, p->tc.WT_hi, p->tc.WT_lo, rp::tc_scale_W(&p->tc));
            ++p->launches;
        }
        ++p->launches;
        RP_LAUNCH_CHECK();
    }
    const bool f16 = p->use_tc && p->tc.f16;
    if (f16 && a->T > 0) {
        // bound of the source operand over the sweep: the maxima tracked by the forward pass that wrote these checkpoints,
        // else one pass over the checkpoints' source planes; rate models: static bound of the activation
        float* sb = p->tc.meta + rp::TCM_SRC_BOUND;
        if (spk) {
            if (p->tc.fwd_history == a->history && p->tc.fwd_T == a->T && p->tc.amax_src) {
                rp::k_max_of_array<<<1, 32, 0, st>>>(p->tc.amax_src, a->T + 1, sb);
            } else {
                RP_CUDA(cudaMemsetAsync(sb, 0, sizeof(float), st));
                rp::k_amax_2d<<<p->sm_count * 4, 256, 0, st>>>(a->T, (int)plane, a->history + plane, (size_t)nhist_of(d.model) * plane, nullptr, 0, sb);
            }
        } else if (d.model == RP_LI_TANH) {
            rp::k_fill<<<1, 32, 0, st>>>(sb, 1, 1.0f);
        } else {
            const int mode = d.param_per_neuron[RP_P_RMAX];
            const int count = mode == 0 ? 1 : (mode == 1 ? N : (mode == 2 ? B : B * N));
            RP_CUDA(cudaMemsetAsync(sb, 0, sizeof(float), st));
            rp::k_amax_2d<<<32, 256, 0, st>>>(1, count, a->params[RP_P_RMAX], (size_t)count, nullptr, 0, sb);
        }
        ++p->launches;
        RP_LAUNCH_CHECK();
        RP_CUDA(cudaMemsetAsync(p->tc.meta + rp::TCM_G_AMAX0, 0, 6 * sizeof(float), st));      // G_AMAX0/1, CHUNK[2][2]
        RP_CUDA(cudaMemsetAsync(p->tc.meta + rp::TCM_NB0, 0, 3 * sizeof(float), st));
    }
    if (need_dW) RP_CUDA(cudaMemsetAsync(p->dWraw, 0, (size_t)wg_slices * N * p->ldw * sizeof(float), st));
    for (int q = 0; q < RP_NUM_PARAMS; ++q)
        if (a->dparams[q]) RP_CUDA(cudaMemsetAsync(a->dparams[q], 0, (size_t)N * sizeof(float), st));
    if (a->dW_in) RP_CUDA(cudaMemsetAsync(a->dW_in, 0, (size_t)N * d.n_in * sizeof(float), st));
    if (a->dW_out) RP_CUDA(cudaMemsetAsync(a->dW_out, 0, (size_t)N * d.n_out * sizeof(float), st));
    if (p->persistent) return persistent_backward(p, a, mp, need_dW, st);
    if (a->g_yT) RP_CUDA(cudaMemcpyAsync(p->adj, a->g_yT, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    else RP_CUDA(cudaMemsetAsync(p->adj, 0, slot * sizeof(float), st));

    const size_t x_stride = d.in_mode == RP_IN_DENSE ? plane : (d.in_mode == RP_IN_PROJ ? (size_t)B * d.n_in : 0);
    const size_t out_stride = d.out_mode == RP_OUT_READOUT ? (size_t)B * d.n_out : plane;
    const size_t hslot = (size_t)nhist_of(d) * plane;
    const int T_tot = a->T_total > 0 ? a->T_total : a->T;
    const int tr = a->truncate_steps;
    const bool truncating = tr > 0 && tr < T_tot;
    const int wg_chunk = p->use_tc ? p->tc.wgrad_chunk : 1;

    // few trials on the FFMA path: batch the rank-B weight-gradient updates of `wg_chunk` steps into one contraction
    const bool wg_batched = !p->use_tc && need_dW && B <= 16;
    if (wg_batched && !p->wg_g) {
        p->wg_chunk = std::max(1, 64 / B);
        if (plan_alloc(p, &p->wg_g, (size_t)p->wg_chunk * plane) || plan_alloc(p, &p->wg_src, (size_t)p->wg_chunk * plane)) return 1;
    }
    int wg_pos = 0;                 // chunk slot holding (g_t, r_t) of the step being processed

    rp::AdjArgs aa;
    memset(&aa, 0, sizeof(aa));
    aa.N = N; aa.B = B; aa.m = d.n_in; aa.k = d.n_out; aa.in_mode = d.in_mode; aa.in_target = d.in_target;
    aa.out_mode = d.out_mode; aa.out_var = d.out_var; aa.dt = d.dt; aa.theta = d.theta; aa.slope = d.slope;
    aa.adj = p->adj; aa.Z = p->u; aa.ldz = p->ldu; aa.W_in = a->W_in; aa.W_out = a->W_out; aa.mp = mp;
    if (!p->use_tc) { aa.g = wg_batched ? p->wg_g : p->g; aa.src = src_plane >= 0 ? nullptr : (wg_batched ? p->wg_src : p->src); }
    else if (!f16) { aa.g_hi = (float*)p->tc.g_hi; aa.g_lo = (float*)p->tc.g_lo; aa.ld_g = p->tc.ldk; aa.ld_t = p->tc.ldt; }
    else { aa.g = p->tc.g32; aa.src = spk ? nullptr : p->tc.src32; }      // binary16: fp32 g + exact maximum, converted by k_adj_convert_f16
    int gslot = 1;                  // binary16: meta slot holding max |g_t| of the step being processed
    int cpar = 0;                   // binary16: parity of the weight-gradient chunk reference slot (within the buffer being filled)
    // binary16: the weight-gradient contraction of a finished K chunk runs on a side stream, cut into one slice of work items per
    // reverse step of the NEXT chunk: each slice starts when that step's adjoint product has finished and its CTAs share the SMs with
    // the HBM-bound adjoint kernels of the step (which are shaped to fit beside a resident GEMM CTA).
    // Opt-in (RP_WG_OVERLAP=1).  Measured on B200 (profiles/r1d_timeline_overlap.txt): co-residency is achieved once the GEMM
    // is capped at 144 registers, given the full shared-memory carveout and the adjoint block needs no shared memory, but a
    // 4-warp adjoint block beside a GEMM CTA streams ~3x slower (the CTA's TMA and MMA operand traffic keeps the SM's L1/shared
    // pipe ~70 % busy), so the step time does not improve: 35.5 ms vs 34.3 ms per 100-step pass without the overlap.
    const bool coresident = f16 && need_dW && getenv("RP_WG_OVERLAP");
    // Second opt-in variant (RP_WG_IDLE_SLICES=1): the adjoint product of a step runs on (n/128) x (batch/256) CTAs -- 128 of the 148
    // SMs at the headline shape -- and a slice of exactly the idle SMs' worth of weight-gradient work items of the previous K chunk is
    // launched beside it; what is left of the chunk runs as one ordinary launch when the next chunk's operands are complete.
    // Measured: with 16-step chunks a work item (K = 8192) takes 140 us, twice the product, so the slice keeps 20 SMs busy through
    // the adjoint step, whose persistent grid then needs a second wave (57 vs 31 us): 35.2 vs 34.1 ms per pass.  With 8-step chunks
    // the items fit (73 us) but only 16 % of a chunk moves off the critical path, which the longer chunk already gains.
    const int zgrid = p->use_tc ? (N / rp::TC_BP) * (B / p->tc.bq_fwd) : 0;
    const int idle_sms = p->sm_count - zgrid;
    const bool idle_slices = f16 && need_dW && !coresident && idle_sms >= 8 && getenv("RP_WG_IDLE_SLICES");
    const bool overlap = coresident || idle_slices;         // weight-gradient slices on the side stream
    int cb_fill = 0, chunks_done = 0;
    bool q_active = false; int q_cb = 0, q_K = 0, q_next = 0, q_ref = 0;
    const int q_total = p->use_tc ? rp::tc_wgrad_items(&p->tc) : 0;
    const int q_per_step = !p->use_tc ? 0 : (coresident ? (q_total + p->tc.wgrad_chunk - 1) / p->tc.wgrad_chunk : idle_sms);
    cudaStream_t ws = overlap ? p->tc.ws : st;
    // next `count` work items of the queued chunk on stream `on` (side stream: slices; main stream: the remainder, idle-SM mode)
    auto launch_slices = [&](int count, cudaStream_t on) -> int {
        count = std::min(count, q_total - q_next);
        if (count <= 0) return 0;
        const rp::ScaleRef sc{p->tc.meta + q_ref, 0.f, rp::CV_HCHUNK};
        if (rp::tc_gemm(&p->tc, rp::TC_WGRAD, p->dWraw, p->ldw, q_K, 1, on, sc, q_cb, q_next, count)) return fail("rp_backward: %s", rp::tc_last_error());
        ++p->launches;
        q_next += count;
        if (q_next == q_total) { q_active = false; if (cudaEventRecord(p->tc.ev_done[q_cb], ws) != cudaSuccess) return fail("cudaEventRecord failed"); }
        return 0;
    };
    cudaStream_t rest_stream = coresident ? ws : st;       // where the remainder of a chunk goes
    for (int q = 0; q < RP_NUM_PARAMS; ++q) aa.dparams[q] = (q == fold) ? nullptr : a->dparams[q];
    aa.dW_in = a->dW_in; aa.dW_out = a->dW_out;
    aa.per_trial = p->per_trial ? 1 : 0;
    aa.any_param_grad = (a->dW_in || a->dW_out) ? 1 : 0;
    for (int q = 0; q < RP_NUM_PARAMS; ++q) if (aa.dparams[q]) aa.any_param_grad = 1;

    bool pgrad = false;
    for (int q = 0; q < RP_NUM_PARAMS; ++q) if (aa.dparams[q]) pgrad = true;
    // vectorised stand-alone adjoint kernel: tensor-core shapes, no per-neuron parameter sums (dW_out then comes from k_readout_grad)
    const bool adj_v4 = p->use_tc && !pgrad && !a->dW_in && !a->g_x && !p->per_trial && !getenv("RP_NO_ADJ_V4");
    // one fused kernel per reverse step instead of k_adj_step_v5 + k_adj_convert_f16 (+ k_readout_grad): the spiking templates whose
    // a_{t-1}[v] does not involve Z_{t-1} (see k_adj_fused_f16)
    const bool fused = f16 && adj_v4 && spk && !rp::is_ik(d.model) && !overlap && aa.src == nullptr && (size_t)B * N < (1u << 31) &&
                       (d.out_mode == RP_OUT_READOUT || !a->g_out_rec) && !(d.model == RP_LIF && d.in_target == 1) && !getenv("RP_NO_ADJ_FUSED");
    const bool fold_ro = fused && a->dW_out && a->g_out_rec && a->T > 0 && d.out_mode == RP_OUT_READOUT && d.out_var == RP_VAR_S;
    int nbr = 0;                    // fused: meta slot TCM_NB0 + nbr holds the bound the next fused launch reads
    int g_scale_slot = rp::TCM_NB0; // fused: slot whose bound scaled the K-major g operand currently in the workspace
    if (fused && a->T > 0) {
        if (a->g_yT) {               // bound of g_{T-1} = dt * gate * a_T[v]
            rp::k_amax_2d<<<p->sm_count * 4, 256, 0, st>>>(1, (int)plane, p->adj, plane, nullptr, 0, p->tc.meta + rp::TCM_NB0);
            rp::k_scale_scalar<<<1, 32, 0, st>>>(p->tc.meta + rp::TCM_NB0, d.dt);
            p->launches += 2;
            RP_LAUNCH_CHECK();
        }
        if (fold_ro) {
            const size_t np = (size_t)(B / rp::FA_TB) * d.n_out * N;
            if (!p->dwout_part) { if (plan_alloc(p, &p->dwout_part, np)) return 1; }
            RP_CUDA(cudaMemsetAsync(p->dwout_part, 0, np * sizeof(float), st));
        }
    }
    aa.mf_t = p->mf; aa.asum = p->asum;
    dim3 agrid((N + rp::ADJ_TX - 1) / rp::ADJ_TX, (B + rp::ADJ_TY * rp::ADJ_BPT - 1) / (rp::ADJ_TY * rp::ADJ_BPT));
    dim3 ablock(rp::ADJ_TX, rp::ADJ_TY);
    int pending = 0;   // steps whose (g, src) columns sit in the tensor-core weight-gradient chunk

    // t == T is the "pre only" launch that builds g_{T-1}
    for (int t = a->T; t >= 0; --t) {
        if (t == a->T && a->T == 0) break;
        aa.do_post = (t < a->T) ? 1 : 0;
        aa.do_pre = (t > 0) ? 1 : 0;
        if (aa.do_post) {
            // Z[b][j] = sum_i (kW)^T[j][i] g_t[b][i]
            if (!p->use_tc) {
                const float* g_cur = wg_batched ? p->wg_g + (size_t)wg_pos * plane : p->g;
                if (gemm_fp32(p, true, N, B, N, p->WkT, p->ldw, g_cur, N, p->u, p->ldu, 0, st, &p->launches)) return 1;
                if (need_dW) {
                    const float* srcp = src_plane >= 0 ? a->history + (size_t)t * hslot + (size_t)src_plane * plane : p->src;
                    if (wg_batched) {
                        if (src_plane >= 0) RP_CUDA(cudaMemcpyAsync(p->wg_src + (size_t)wg_pos * plane, srcp, plane * sizeof(float), cudaMemcpyDeviceToDevice, st));
                        if (wg_pos + 1 == p->wg_chunk || t == 0) {
                            // dWraw[i][j] += sum over the chunk's (step, trial) rows of r[j] * g[i]
                            if (gemm_fp32(p, false, N, N, (wg_pos + 1) * B, p->wg_src, N, p->wg_g, N, p->dWraw, p->ldw, 1, st, &p->launches)) return 1;
                            wg_pos = -1;
                        }
                        ++wg_pos;
                        aa.g = p->wg_g + (size_t)wg_pos * plane;                 // where this launch's "pre" puts g_{t-1}, r_{t-1}
                        if (src_plane < 0) aa.src = p->wg_src + (size_t)wg_pos * plane;
                    } else if (B <= 16) {
                        RP_CUDA(rp::outer_acc_launch(N, B, p->g, N, srcp, N, p->dWraw, p->ldw, st));
                        ++p->launches;
                    } else {
                        // dWraw[i][j] += sum_b src[b][j] g[b][i]   (p = j, q = i)
                        if (gemm_fp32(p, false, N, N, B, srcp, N, p->g, N, p->dWraw, p->ldw, 1, st, &p->launches)) return 1;
                    }
                }
            }
            const Window w = window_of(a->t_offset + t, T_tot, a->sampling_steps, a->cutoff);
            aa.y_t = a->history + (size_t)t * hslot;
            aa.urec_t = (rp::is_ik(d.model) || jit) ? a->history + (size_t)t * hslot + (size_t)nsv * plane : nullptr;
            if (rp::is_mean_field(d.model)) {
                // population means of y_t and of the incoming adjoint of u, one block per trial, before the element-wise adjoint
                rp::k_trial_means<<<B, 256, 0, st>>>(N, aa.y_t, d.theta, p->mf);
                rp::k_trial_adj_sums<<<B, 256, 0, st>>>(N, B, p->adj + 2 * plane, mp, d.dt, p->asum);
                p->launches += 2;
                RP_LAUNCH_CHECK();
            }
            aa.x_t = a->x ? a->x + (size_t)t * x_stride : nullptr;
            aa.e_t = (a->g_out_rec && w.j >= 0) ? a->g_out_rec + (size_t)w.j * out_stride : nullptr;
            aa.e_scale = w.j >= 0 ? 1.0f / (float)w.len : 0.f;
            aa.g_x_t = a->g_x ? a->g_x + (size_t)t * plane : nullptr;
            aa.zero_after_post = (truncating && (a->t_offset + t) > 0 && (a->t_offset + t) % tr == 0) ? 1 : 0;
        }
        if (jit) {
            aa.y_t = a->history + (size_t)t * hslot;
            aa.e_tm1 = nullptr; aa.e_scale_tm1 = 0.f;
            if (t > 0 && a->g_out_rec) {
                const Window w1 = window_of(a->t_offset + t - 1, T_tot, a->sampling_steps, a->cutoff);
                if (w1.j >= 0) { aa.e_tm1 = a->g_out_rec + (size_t)w1.j * out_stride; aa.e_scale_tm1 = 1.0f / (float)w1.len; }
            }
        }
        if (aa.do_pre) {
            aa.y_tm1 = a->history + (size_t)(t - 1) * hslot;
            if (p->use_tc && need_dW && !f16) {
                aa.gT_hi = (float*)p->tc.gT_hi[0]; aa.gT_lo = (float*)p->tc.gT_lo[0]; aa.srcT_hi = (float*)p->tc.srcT_hi[0]; aa.srcT_lo = (float*)p->tc.srcT_lo[0];
                aa.t_col0 = pending * B;
            }
            if (f16) aa.g_amax = p->tc.meta + rp::TCM_G_AMAX0 + (gslot ^ 1);
        }
        {
            if (p->use_tc && aa.do_post) {
                const rp::ScaleRef sg = fused ? rp::ScaleRef{p->tc.meta + g_scale_slot, 0.f, rp::CV_HGB}
                                              : (f16 ? rp::ScaleRef{p->tc.meta + rp::TCM_G_AMAX0 + gslot, 0.f, rp::CV_HG} : rp::no_scale());
                // The slice becomes eligible together with this step's adjoint product (both wait for the previous kernel of the
                // chain) but is enqueued after it: the product takes its 128 SMs first, the slice's CTAs take the idle SMs and
                // then every SM the product frees -- before the adjoint kernels, which fit beside them, become eligible.
                const bool slice_now = overlap && q_active;
                if (slice_now) {
                    RP_CUDA(cudaEventRecord(p->tc.ev_z, st));
                    RP_CUDA(cudaStreamWaitEvent(ws, p->tc.ev_z, 0));
                }
                stage_mark(p, ST_DGRAD, st);
                if (rp::tc_gemm(&p->tc, rp::TC_DGRAD, p->u, p->ldu, 0, 0, st, sg)) return fail("rp_backward: %s", rp::tc_last_error());
                ++p->launches;
                if (slice_now && launch_slices(q_per_step, ws)) return 1;
            }
            stage_mark(p, ST_ADJ, st);
            if (fused) {
                rp::AdjArgs va = aa;
                va.dW_out = nullptr; va.any_param_grad = 0; va.g = nullptr; va.g_amax = nullptr;
                rp::FusedAdjArgs fa;
                memset(&fa, 0, sizeof(fa));
                fa.g_hi = p->tc.g_hi; fa.g_lo = p->tc.g_lo; fa.ld_g = p->tc.ldk;
                if (need_dW && aa.do_pre) {
                    fa.gT_hi = p->tc.gT_hi[0]; fa.gT_lo = p->tc.gT_lo[0]; fa.srcT_hi = p->tc.srcT_hi[0]; fa.srcT_lo = p->tc.srcT_lo[0];
                    fa.ld_t = p->tc.ldt; fa.t_col0 = pending * B;
                }
                fa.nb_in = p->tc.meta + rp::TCM_NB0 + nbr;
                fa.nb_out = p->tc.meta + rp::TCM_NB0 + (nbr + 1) % 3;
                fa.nb_clear = p->tc.meta + rp::TCM_NB0 + (nbr + 2) % 3;
                fa.chunk_ref_in = p->tc.meta + rp::TCM_CHUNK0 + cpar;
                fa.chunk_ref_out = p->tc.meta + rp::TCM_CHUNK0 + (cpar ^ 1);
                fa.chunk_first = pending == 0 ? 1 : 0;
                fa.sc_src = rp::tc_scale_srcbound(&p->tc);
                fa.flags = reinterpret_cast<int*>(p->tc.meta + rp::TCM_FLAGS);
                if (aa.do_pre && a->g_out_rec) {
                    const Window w1 = window_of(a->t_offset + t - 1, T_tot, a->sampling_steps, a->cutoff);
                    if (w1.j >= 0) { fa.e_tm1 = a->g_out_rec + (size_t)w1.j * out_stride; fa.e_scale_tm1 = 1.0f / (float)w1.len; }
                }
                fa.dwout_part = fold_ro ? p->dwout_part : nullptr;
                const dim3 fgrid(N / rp::FA_TN, B / rp::FA_TB);
                const bool full = aa.do_post && aa.do_pre;
#define RP_FUSED_LAUNCH(M_) (full ? rp::launch_pdl(rp::k_adj_fused_f16<M_, true>, fgrid, dim3(256), 0, st, va, fa) \
                                  : rp::launch_pdl(rp::k_adj_fused_f16<M_, false>, fgrid, dim3(256), 0, st, va, fa))
                switch (d.model) {
                    case RP_QIF:     RP_FUSED_LAUNCH(RP_QIF); break;
                    case RP_QIF_SFA: RP_FUSED_LAUNCH(RP_QIF_SFA); break;
                    case RP_LIF:     RP_FUSED_LAUNCH(RP_LIF); break;
                    default: return fail("rp_backward: internal error (fused adjoint kernel dispatch)");
                }
#undef RP_FUSED_LAUNCH
                g_scale_slot = rp::TCM_NB0 + nbr;
                nbr = (nbr + 1) % 3;
                if (aa.do_pre) cpar ^= 1;
            } else if (adj_v4) {
                rp::AdjArgs va = aa;
                va.dW_out = nullptr; va.any_param_grad = 0;
                // rolling-pipeline kernel: spiking templates whose source value is not needed here (the conversion kernel reads
                // s_{t-1} from the checkpoint itself), enough warps for one per neuron tile
                const bool v5 = f16 && !coresident && spk && !rp::is_mean_field(d.model) && va.src == nullptr && p->sm_count * 24 >= N / 128 && !getenv("RP_NO_ADJ_V5");
                if (v5) {
                    {
                        switch (d.model) {
                            case RP_QIF:     rp::launch_pdl(rp::k_adj_step_v5<RP_QIF>, dim3(p->sm_count * 3), dim3(256), 0, st, va); break;
                            case RP_QIF_SFA: rp::launch_pdl(rp::k_adj_step_v5<RP_QIF_SFA>, dim3(p->sm_count * 3), dim3(256), 0, st, va); break;
                            case RP_LIF:     rp::launch_pdl(rp::k_adj_step_v5<RP_LIF>, dim3(p->sm_count * 3), dim3(256), 0, st, va); break;
                            case RP_IK:      rp::launch_pdl(rp::k_adj_step_v5<RP_IK>, dim3(p->sm_count * 3), dim3(256), 0, st, va); break;
                            default: return fail("rp_backward: internal error (adjoint kernel dispatch)");
                        }
                    }
                } else if (f16 && !coresident) {
                    RP_DISPATCH_MODEL(d.model, (rp::launch_pdl(rp::k_adj_step_v4<M_, false, 8>, dim3(N / 128, B / rp::ADJ4_TB), ablock, 0, st, va)));
                } else if (f16) {
                    RP_DISPATCH_MODEL(d.model, {
                        // same shared-memory carveout as the GEMM CTAs, or the block cannot become resident beside one
                        static bool carve = false;
                        if (!carve) { cudaFuncSetAttribute(rp::k_adj_step_v4<M_, false, 4>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); carve = true; }
                        rp::launch_pdl(rp::k_adj_step_v4<M_, false, 4>, dim3(N / 128, B / 16), dim3(32, 4), 0, st, va);
                    });
                }
                else     { RP_DISPATCH_MODEL(d.model, (rp::launch_pdl(rp::k_adj_step_v4<M_, true, 8>, dim3(N / 128, B / rp::ADJ4_TB), ablock, 0, st, va))); }
            } else if (jit) {
                if (jit_launch(p->jit_adj, ew_grid(p, plane), 256, &aa, st)) return 1;
            } else {
                RP_DISPATCH_MODEL(d.model, (rp::k_adj_step<M_><<<agrid, ablock, 0, st>>>(aa)));
            }
            RP_LAUNCH_CHECK();
        }
        ++p->launches;
        if (f16 && aa.do_pre && !fused) {
            // the buffer about to be refilled must have been consumed by the side stream (chunk index - 2)
            if (overlap && pending == 0 && chunks_done >= 2) RP_CUDA(cudaStreamWaitEvent(st, p->tc.ev_done[cb_fill], 0));
            rp::ConvArgs ca;
            memset(&ca, 0, sizeof(ca));
            ca.N = N; ca.B = B; ca.g32 = p->tc.g32;
            ca.src32 = spk ? a->history + (size_t)(t - 1) * hslot + plane : p->tc.src32;
            ca.g_hi = p->tc.g_hi; ca.g_lo = p->tc.g_lo; ca.ld_g = p->tc.ldk;
            if (need_dW) {
                ca.gT_hi = p->tc.gT_hi[cb_fill]; ca.gT_lo = p->tc.gT_lo[cb_fill]; ca.srcT_hi = p->tc.srcT_hi[cb_fill]; ca.srcT_lo = p->tc.srcT_lo[cb_fill];
                ca.ld_t = p->tc.ldt; ca.t_col0 = pending * B;
            }
            ca.g_amax = p->tc.meta + rp::TCM_G_AMAX0 + (gslot ^ 1);
            ca.g_amax_clear = p->tc.meta + rp::TCM_G_AMAX0 + gslot;
            ca.chunk_ref_in = p->tc.meta + rp::TCM_CHUNK0 + 2 * cb_fill + cpar;
            ca.chunk_ref_out = p->tc.meta + rp::TCM_CHUNK0 + 2 * cb_fill + (cpar ^ 1);
            ca.chunk_first = pending == 0 ? 1 : 0;
            ca.sc_src = rp::tc_scale_srcbound(&p->tc);
            ca.flags = reinterpret_cast<int*>(p->tc.meta + rp::TCM_FLAGS);
            {
                static bool carve = false;
                if (!carve) { cudaFuncSetAttribute(rp::k_adj_convert_f16, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared); carve = true; }
            }
            rp::launch_pdl(rp::k_adj_convert_f16, dim3(N / rp::CV_TN, B / rp::CV_TB), dim3(256), 0, st, ca);
            RP_LAUNCH_CHECK();
            ++p->launches;
            gslot ^= 1; cpar ^= 1;
        }
        if (p->use_tc && need_dW && aa.do_pre) {
            ++pending;
            if (pending == wg_chunk || t == 1) {
                // dWraw[i][j] += sum_{(t,b) in chunk} g[b][i] src[b][j]
                const int ref_slot = rp::TCM_CHUNK0 + 2 * cb_fill + cpar;
                if (!overlap) {
                    const rp::ScaleRef sc = f16 ? rp::ScaleRef{p->tc.meta + ref_slot, 0.f, rp::CV_HCHUNK} : rp::no_scale();
                    stage_mark(p, ST_WGRAD, st);
                    if (rp::tc_gemm(&p->tc, rp::TC_WGRAD, p->dWraw, p->ldw, pending * B, 1, st, sc, f16 ? cb_fill : 0)) return fail("rp_backward: %s", rp::tc_last_error());
                    ++p->launches;
                } else {
                    if (q_active && launch_slices(q_total, rest_stream)) return 1;  // whatever is left of the previous chunk
                    RP_CUDA(cudaEventRecord(p->tc.ev_ops[cb_fill], st));           // (after that launch: the next slices touch the same dW tiles)
                    RP_CUDA(cudaStreamWaitEvent(ws, p->tc.ev_ops[cb_fill], 0));
                    q_active = true; q_cb = cb_fill; q_K = pending * B; q_next = 0; q_ref = ref_slot;
                    if (t == 1 && launch_slices(q_total, rest_stream)) return 1;   // last chunk: only the final post step is left to overlap with
                    cb_fill ^= 1; cpar = 0; ++chunks_done;
                }
                pending = 0;
            }
        }
    }
    stage_mark(p, ST_OTHER, st);
    if (overlap) {
        if (q_active && launch_slices(q_total, rest_stream)) return 1;
        RP_CUDA(cudaEventRecord(p->tc.ev_join, ws));
        RP_CUDA(cudaStreamWaitEvent(st, p->tc.ev_join, 0));
    }
    if (fold_ro) {
        const size_t nk = (size_t)d.n_out * N;
        rp::k_sum_parts<<<(int)std::min<size_t>((nk + 255) / 256, 1024), 256, 0, st>>>(p->dwout_part, B / rp::FA_TB, nk, a->dW_out);
        ++p->launches;
        RP_LAUNCH_CHECK();
    } else if (adj_v4 && a->dW_out && a->g_out_rec && a->T > 0) {
        for (int t0 = 0; t0 < a->T; t0 += 32768) {            // grid.y limit
            const int tn = std::min(32768, a->T - t0);
            dim3 rgrid((N + 127) / 128, tn, (B + rp::RG_TB - 1) / rp::RG_TB);
            RP_DISPATCH_MODEL(d.model, (rp::k_readout_grad<M_><<<rgrid, 128, 0, st>>>(N, B, tn, a->sampling_steps, a->cutoff, d.n_out, d.out_var,
                                                                                       a->history + (size_t)t0 * hslot, a->g_out_rec, mp, a->dW_out,
                                                                                       a->t_offset + t0, T_tot)));
            ++p->launches;
            RP_LAUNCH_CHECK();
        }
    }
    if (need_dW) {
        rp::k_finish_wgrad<<<N, 256, 0, st>>>(N, p->dWraw, p->ldw, a->W, a->params[fold], kstride, a->dW, a->dparams[fold], wg_slices);
        ++p->launches;
        RP_LAUNCH_CHECK();
    }
    if (a->g_y0) RP_CUDA(cudaMemcpyAsync(a->g_y0, p->adj, slot * sizeof(float), cudaMemcpyDeviceToDevice, st));
    return 0;
}

int rp_plan_time_contraction(rp_plan* p, int which, int iters, float* avg_ms, double* flops, void* stream) {
    if (!p || !avg_ms || !flops || iters <= 0 || which < 0 || which > 2) return fail("rp_plan_time_contraction: bad argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    const int N = p->d.n, B = p->d.batch;
    const size_t plane = (size_t)B * N;
    long long scratch_launches = 0;
    int kext = B;
    if (which == 2) {
        const size_t nraw = (size_t)(p->use_tc ? rp::TC_WGRAD_SPLITS : 1) * N * p->ldw;
        if (!p->dWraw) { if (plan_alloc(p, &p->dWraw, nraw)) return 1; RP_CUDA(cudaMemsetAsync(p->dWraw, 0, nraw * sizeof(float), st)); }
        if (p->use_tc) {
            size_t bytes = 0;
            if (rp::tc_workspace_ensure_wgrad(&p->tc, &bytes)) return fail("rp_plan_time_contraction: %s", rp::tc_last_error());
            p->ws_bytes += bytes;
            kext = p->tc.wgrad_chunk * B;
        }
    }
    if (!p->use_tc && !p->src) { if (plan_alloc(p, &p->src, plane)) return 1; RP_CUDA(cudaMemsetAsync(p->src, 0, plane * sizeof(float), st)); }
    auto launch = [&]() -> int {
        if (p->use_tc) {
            const rp::ScaleRef sb = p->tc.f16 ? rp::ScaleRef{p->tc.meta + rp::TCM_G_AMAX0, 0.f, rp::CV_HG} : rp::no_scale();
            if (which == 2) return rp::tc_gemm(&p->tc, rp::TC_WGRAD, p->dWraw, p->ldw, kext, 1, st, sb);
            return rp::tc_gemm(&p->tc, which == 0 ? rp::TC_FWD : rp::TC_DGRAD, p->u, p->ldu, 0, 0, st, sb);
        }
        if (which == 2) return gemm_fp32(p, false, N, N, B, p->src, N, p->g, N, p->dWraw, p->ldw, 1, st, &scratch_launches);
        return gemm_fp32(p, true, N, B, N, which == 0 ? p->Wk : p->WkT, p->ldw, which == 0 ? p->src : p->g, N, p->u, p->ldu, 0, st, &scratch_launches);
    };
    for (int i = 0; i < 2; ++i) if (launch()) return fail("rp_plan_time_contraction: %s", p->use_tc ? rp::tc_last_error() : g_err);
    cudaEvent_t e0, e1;
    RP_CUDA(cudaEventCreate(&e0));
    RP_CUDA(cudaEventCreate(&e1));
    RP_CUDA(cudaEventRecord(e0, st));
    for (int i = 0; i < iters; ++i) if (launch()) return fail("rp_plan_time_contraction: launch failed");
    RP_CUDA(cudaEventRecord(e1, st));
    RP_CUDA(cudaEventSynchronize(e1));
    float ms = 0.f;
    RP_CUDA(cudaEventElapsedTime(&ms, e0, e1));
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    *avg_ms = ms / (float)iters;
    *flops = which == 2 ? 2.0 * N * (double)N * kext : 2.0 * N * (double)N * B;
    return 0;
}

int rp_gemm_tn(int precision, int P, int Q, int K, const float* A, int lda, const float* B, int ldb,
               float* C, int ldc, int accumulate, void* stream) {
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    long long launches = 0;
    if (precision == RP_PREC_FP32) return gemm_fp32(nullptr, true, P, Q, K, A, lda, B, ldb, C, ldc, accumulate, st, &launches);
    if (precision == RP_PREC_3XTF32 || precision == RP_PREC_3XF16) {
        if (rp::tc_gemm_standalone(precision == RP_PREC_3XF16, P, Q, K, A, lda, B, ldb, C, ldc, accumulate, st)) return fail("rp_gemm_tn: %s", rp::tc_last_error());
        return 0;
    }
    return fail("rp_gemm_tn: unknown precision %d", precision);
}

/* Debug / profiling: arm (capacity > 0) or disarm (0) the CTA timeline of the traced kernels; rp_trace_read copies the records
 * (16 bytes: tag, smid, start ns, end ns as {u32, u32, u64, u64} = 24 bytes each) collected so far and returns their number. */
static rp::TraceRec* g_trace_dev = nullptr;
static unsigned g_trace_capacity = 0;
int rp_trace_enable(int capacity) {
    cudaDeviceSynchronize();
    rp::TraceRec* null_buf = nullptr;
    unsigned zero = 0;
    if (g_trace_dev) { cudaFree(g_trace_dev); g_trace_dev = nullptr; g_trace_capacity = 0; }
    if (capacity > 0) {
        RP_CUDA(cudaMalloc(reinterpret_cast<void**>(&g_trace_dev), (size_t)capacity * sizeof(rp::TraceRec)));
        RP_CUDA(cudaMemset(g_trace_dev, 0, (size_t)capacity * sizeof(rp::TraceRec)));
        g_trace_capacity = (unsigned)capacity;
    }
    RP_CUDA(cudaMemcpyToSymbol(rp::g_trace_n, &zero, sizeof(unsigned)));
    RP_CUDA(cudaMemcpyToSymbol(rp::g_trace_cap, &g_trace_capacity, sizeof(unsigned)));
    RP_CUDA(cudaMemcpyToSymbol(rp::g_trace_buf, g_trace_dev ? &g_trace_dev : &null_buf, sizeof(rp::TraceRec*)));
    return 0;
}
int rp_trace_read(void* host_buf, int max_records) {
    if (!g_trace_dev || !host_buf || max_records <= 0) return 0;
    cudaDeviceSynchronize();
    unsigned n = 0;
    if (cudaMemcpyFromSymbol(&n, rp::g_trace_n, sizeof(unsigned)) != cudaSuccess) return -1;
    n = std::min(n, std::min(g_trace_capacity, (unsigned)max_records));
    if (cudaMemcpy(host_buf, g_trace_dev, (size_t)n * sizeof(rp::TraceRec), cudaMemcpyDeviceToHost) != cudaSuccess) return -1;
    return (int)n;
}

int rp_plan_stage_timing(rp_plan* p, int enable) {
    if (!p) return fail("rp_plan_stage_timing: null argument");
    p->timing = enable != 0;
    p->ev_used = 0;
    return 0;
}

int rp_plan_stage_times(rp_plan* p, float* ms, int* marks, void* stream) {
    if (!p || !ms || !marks) return fail("rp_plan_stage_times: null argument");
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    for (int c = 0; c < ST_COUNT; ++c) { ms[c] = 0.f; marks[c] = 0; }
    if (p->ev_used == 0) return 0;
    const bool was = p->timing;
    p->timing = true;
    stage_mark(p, ST_OTHER, st);          // closing event
    p->timing = was;
    RP_CUDA(cudaEventSynchronize(p->ev_pool[p->ev_used - 1]));
    for (size_t e = 0; e + 1 < p->ev_used; ++e) {
        float dtm = 0.f;
        RP_CUDA(cudaEventElapsedTime(&dtm, p->ev_pool[e], p->ev_pool[e + 1]));
        ms[p->ev_cat[e]] += dtm;
        ++marks[p->ev_cat[e]];
    }
    p->ev_used = 0;
    return 0;
}

int rp_plan_status(rp_plan* p, void* stream) {
    if (!p) return fail("rp_plan_status: null argument");
    if (!(p->use_tc && p->tc.f16)) return 0;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    int flags = 0;
    int* dflags = reinterpret_cast<int*>(p->tc.meta + rp::TCM_FLAGS);
    RP_CUDA(cudaMemcpyAsync(&flags, dflags, sizeof(int), cudaMemcpyDeviceToHost, st));
    RP_CUDA(cudaStreamSynchronize(st));
    if (flags & 1) {
        RP_CUDA(cudaMemsetAsync(dflags, 0, sizeof(int), st));
        return fail("rp_plan_status: the adjoint grew by more than 2^14 within one weight-gradient chunk, which exceeds the binary16 "
                    "operand range of RP_PREC_3XF16; the gradients of this call are invalid -- use RP_PREC_3XTF32 for this problem");
    }
    return 0;
}

int rp_rls_run(int T, int n_in, int n_out, float beta_inv, const float* X, const float* Y, float* W, float* P,
               float* loss, float* pred, int update_every, void* stream) {
    if (T < 0 || n_in <= 0 || n_out <= 0 || !X || !Y || !W || !P) return fail("rp_rls_run: bad argument");
    if (n_out > RP_RLS_MAX_OUT) return fail("rp_rls_run: n_out must be <= %d", RP_RLS_MAX_OUT);
    if (update_every <= 0) update_every = 1;
    cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
    if (rp::rls_run(T, n_in, n_out, beta_inv, X, Y, W, P, loss, pred, update_every, st)) return fail("rp_rls_run: %s", rp::rls_last_error());
    return 0;
}

}  // extern "C"
