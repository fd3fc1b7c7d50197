/*
 * rectipy_b200 -- C ABI of the B200 (sm_100a) engine for RectiPy's time-stepped integration hot path.
 *
 * The reference (pyrates-neuroscience/RectiPy) is pure Python and has no FFI; what this library replaces is the
 * per-step operator boundary of the reference, executed for a whole horizon at once:
 *
 *   rp_forward   <->  Network.run step loop            rectipy/network.py:588-599
 *                     Network.forward / _backward      rectipy/network.py:462-478,962-977
 *                     Linear.forward (W_in, W_out)     rectipy/edges.py:48-49
 *                     RateNet.forward                  rectipy/nodes.py:166-170
 *                     SpikeResetNet.forward            rectipy/nodes.py:382-392
 *                     Spike.forward (heaviside)        rectipy/nodes.py:473-476
 *                     generated vector field f(t,y,*args), equations neuron_model_templates/ ** .yaml
 *                     Observer.record window means     rectipy/network.py:590-597, rectipy/observer.py:79-105
 *   rp_backward  <->  error.backward() over the unrolled tape      rectipy/network.py:1123-1130
 *                     Spike.backward (surrogate)       rectipy/nodes.py:478-481
 *                     truncated BPTT detach            rectipy/network.py:598-599, rectipy/nodes.py:176-196
 *   rp_rls_run   <->  RLS.update applied per step      rectipy/edges.py:227-234, rectipy/network.py:1093-1121
 *   rp_plan_set_jit_module  <->  the run function PyRates generates for a template that is not one of the compiled vector fields
 *                     (RateNet.from_pyrates / _circuit_from_yaml -> get_run_func)   rectipy/nodes.py:112-164,232-262
 *                     incl. MultiSpikeResetNet.forward (several spike / reset pairs, post-update outputs)   rectipy/nodes.py:451-465
 *
 * Conventions
 *   - Plain C: pointers + sizes, no torch types.  All data pointers are DEVICE pointers to fp32, contiguous,
 *     16-byte aligned, owned by the caller; the library borrows them for the duration of the call.
 *   - `stream` is a cudaStream_t passed as void*; all work is enqueued on it, nothing synchronises the host.
 *   - Every entry point returns 0 on success, non-zero on error; rp_last_error() gives a thread-local message.
 *   - State layout ("SoA planes"): y[var][trial][neuron], var in {0:v, 1:s, 2:x}; plane stride = batch*n.
 *     (ik_op / iku_op: plane 2 holds the recovery variable u; ik_biexp_op: planes v, s, u, x; RP_JIT: the reset variables first,
 *     then the remaining state variables in equation order.)
 *   - A plan is not re-entrant: one host thread / one stream at a time per plan.
 */
#ifndef RECTIPY_B200_H
#define RECTIPY_B200_H

#ifdef __cplusplus
extern "C" {
#endif

#define RP_ABI_VERSION 8
#define RP_MAX_IN 8      /* max fused input-projection width  m (wider inputs: use RP_IN_DENSE)  */
#define RP_MAX_OUT 8     /* max fused readout width           k (wider readouts: use RP_OUT_DENSE) */
#define RP_MAX_SV 4
#define RP_MAX_REC 4

/* vector fields (neuron_model_templates/rate_neurons/leaky_integrator.yaml, spiking_neurons/{qif,lif}.yaml) */
enum { RP_LI_TANH = 0, RP_LI_SIGMOID = 1, RP_QIF = 2, RP_QIF_SFA = 3, RP_LIF = 4, RP_IK = 5 /* spiking_neurons/ik.yaml ik_op */,
       RP_IKU = 6 /* ik.yaml iku_op: recovery variable driven by the per-trial population means of v and of the spikes */,
       RP_IK_BIEXP = 7 /* ik.yaml:42-70 ik_biexp_op: iku_op with a bi-exponential synapse s' = -s/tau_d + x, x' = -x/tau_r + spike;
                          tau_d travels in slot RP_P_TAU_S, tau_r in slot RP_P_TAU_X */,
       RP_JIT = 8 /* a vector field compiled at run time from a user template (rp_plan_set_jit_module); rp_desc.jit_* describe it */ };
/* parameter slots; each is a device pointer to 1, n, B or B*n floats (see rp_desc.param_per_neuron) */
enum { RP_P_TAU = 0, RP_P_K, RP_P_ETA, RP_P_TAU_S, RP_P_TAU_X, RP_P_ALPHA, RP_P_RMAX, RP_P_SIG_S, RP_P_V0,
       /* ik_op: */ RP_P_C, RP_P_VR, RP_P_VTH, RP_P_G, RP_P_ER, RP_P_B, RP_P_TAU_U, RP_P_KAPPA, RP_NUM_PARAMS };
/* The coupling constant that scales the recurrent input (k for li/qif/lif, g for ik) is folded into the weights by the
 * engine; its gradient is returned in its dparams slot as sum_j dWraw[i][j] * W[i][j]. */
enum { RP_IN_NONE = 0, RP_IN_DENSE = 1, RP_IN_PROJ = 2 };   /* x_t is [B,n] current | [B,m] projected by W_in[n,m] */
enum { RP_OUT_DENSE = 0, RP_OUT_READOUT = 1 };             /* record y[out] itself | W_out[k,n] . y[out]         */
enum { RP_VAR_V = 0, RP_VAR_S = 1, RP_VAR_X = 2, RP_VAR_R = 3 }; /* RP_VAR_R = activation(v) of a rate node       */
/* FFMA fp32 | tcgen05 error-compensated split products, operands split into two tf32 words | the same split carried by two
 * binary16 words with exact power-of-two operand scales (same 11-bit significands -> same accuracy, twice the MMA rate) */
enum { RP_PREC_FP32 = 0, RP_PREC_3XTF32 = 1, RP_PREC_3XF16 = 2 };

typedef struct rp_desc {
    int model;              /* RP_LI_TANH ...                                                   */
    int n;                  /* neurons                                                          */
    int batch;              /* independent trials B (reference: 1)                              */
    int in_mode;            /* RP_IN_*                                                          */
    int n_in;               /* m (RP_IN_PROJ only)                                              */
    int in_target;          /* 0: input enters v' (I_ext), 1: input enters s' (lif s_ext)       */
    int out_mode;           /* RP_OUT_*                                                         */
    int n_out;              /* k (RP_OUT_READOUT only)                                          */
    int out_var;            /* RP_VAR_*: which variable is the node output                      */
    int precision;          /* RP_PREC_*                                                        */
    float dt;
    float theta;            /* spike threshold   (nodes.py:338,349)                             */
    float v_reset;          /* reset value       (nodes.py:338,348)                             */
    float slope;            /* surrogate slope   (nodes.py:345-347)                             */
    int param_per_neuron[RP_NUM_PARAMS]; /* layout of each parameter: 0 one shared value, 1 [n] per neuron,
                                            2 [B] per trial, 3 [B][n] per trial and neuron (parameter sweeps; no gradients) */
    /* RP_JIT only: */
    int jit_nsv;            /* state variables (1..RP_MAX_SV); plane 0 is the reset variable of a spiking field */
    int jit_spiking;        /* number of spike variables: planes 0..jit_spiking-1 are thresholded / reset with the surrogate gradient
                               (1: SpikeResetNet, several: MultiSpikeResetNet), 0: RateNet */
    int jit_post_out;       /* 1: outputs are POST-update slices (MultiSpikeResetNet.forward, rectipy/nodes.py:451-465) */
    int jit_src_plane;      /* plane projected by the recurrent weights, or -1: the source is an expression of the state */
} rp_desc;

typedef struct rp_plan rp_plan;

typedef struct rp_fwd_args {
    int T;                  /* integration steps                                                */
    int sampling_steps;     /* S  (network.py:592)                                              */
    int cutoff;             /*    (network.py:590)                                              */
    const float* x;         /* [T,B,m] (PROJ) | [T,B,n] (DENSE) | NULL                          */
    const float* W;         /* [n,n] recurrent weights, row = target, col = source              */
    const float* W_in;      /* [n,m] or NULL                                                    */
    const float* W_out;     /* [k,n] or NULL                                                    */
    const float* params[RP_NUM_PARAMS];
    const float* y0;        /* [n_sv,B,n] initial state                                         */
    float* yT;              /* [n_sv,B,n] state after T steps                                   */
    float* out_rec;         /* [n_rec,B,k] | [n_rec,B,n] window means of the output, or NULL    */
    int n_rec_vars;
    int rec_var[RP_MAX_REC];     /* state plane index 0..n_sv-1 (RP_VAR_V/S/X, 3: ik_biexp_op x)  */
    int rec_reduce[RP_MAX_REC];  /* 1: mean over neurons -> [n_rec,B]; 0: [n_rec,B,n]           */
    float* rec_buf[RP_MAX_REC];
    float* history;         /* [(T+1),n_hist,B,n] state checkpoints for rp_backward (n_hist = rp_num_history_planes), or NULL */
    /* Segmented horizons (checkpoint + recompute): this call integrates global steps [t_offset, t_offset+T) of a run of
     * T_total steps (0: T_total = T).  Record windows are those of the whole run; out_rec / rec_buf point at record 0 of the
     * whole run.  Segment starts must not split a record window: t_offset = 0, or (t_offset-1) % sampling_steps == 0. */
    int t_offset;
    int T_total;
} rp_fwd_args;

typedef struct rp_bwd_args {
    int T, sampling_steps, cutoff;
    int truncate_steps;     /* 0/>=T: full BPTT; else adjoint cut after steps t%tr==tr-1        */
    const float* x;
    const float* W;
    const float* W_in;
    const float* W_out;
    const float* params[RP_NUM_PARAMS];
    const float* history;   /* as written by rp_forward                                         */
    const float* g_out_rec; /* dL/d out_rec, same shape as out_rec, or NULL                     */
    const float* g_yT;      /* dL/d yT [n_sv,B,n] or NULL                                       */
    float* dW;              /* [n,n]  (overwritten) or NULL                                     */
    float* dW_in;           /* [n,m]  (overwritten) or NULL                                     */
    float* dW_out;          /* [k,n]  (overwritten) or NULL                                     */
    float* dparams[RP_NUM_PARAMS]; /* each [n] per-neuron sums over (t,trial) (overwritten) or NULL;
                                      the caller reduces over neurons for shared parameters     */
    float* g_y0;            /* [n_sv,B,n] or NULL                                               */
    float* g_x;             /* RP_IN_DENSE only: dL/dx [T,B,n] or NULL                          */
    int t_offset;           /* as in rp_fwd_args; x / history / g_x are those of the segment, g_out_rec of the whole run */
    int T_total;
} rp_bwd_args;

#define RP_NUM_STAGES 5
#ifndef __CUDACC_RTC__   /* device code compiled at run time (RP_JIT) includes this header for the constants and records only */
int         rp_abi_version(void);
const char* rp_last_error(void);
int         rp_num_state_vars(int model);                       /* LI 1, QIF/LIF 2, QIF-SFA/IK/IKU 3, IK_BIEXP 4 */
int         rp_num_history_planes(int model);                   /* planes per checkpoint slot: n_sv (+1 for ik: the recurrent drive) */
int         rp_num_records(int T, int sampling_steps, int cutoff); /* records produced by a run  */

int  rp_plan_create(const rp_desc* desc, rp_plan** plan);
/* Which execution path a plan with this description would take on the current device: 0 per-step launches (fp32), 1 the persistent
 * few-trial kernels (fp32, the owned rows of kW resident in shared memory), 2 the tcgen05 path; -1 on error.  The host side uses it
 * to decide whether padding the trial / neuron axes to multiples of 128 (tensor-core path) beats the per-step fp32 path. */
int  rp_plan_path(const rp_desc* desc);
void rp_plan_destroy(rp_plan* plan);
/* bytes of device workspace the plan holds (diagnostics) */
long long rp_plan_workspace_bytes(const rp_plan* plan);
/* kernels launched by this plan since creation (for bench.py's gpu_launches) */
long long rp_plan_launch_count(const rp_plan* plan);

/* RP_JIT: hand the plan the device code of its vector field -- a cubin / PTX image (e.g. from NVRTC, see rectipy_b200/jit.py) that
 * defines the  extern "C" __global__  kernels `rp_jit_init_src`, `rp_jit_fwd_step` (one rp::FwdStepArgs record) and `rp_jit_adj_step`
 * (one rp::AdjArgs record) against the argument records of rectipy_b200/csrc/rp_jit_abi.cuh.  This replaces what PyRates' code generation does in the reference
 * (rectipy/nodes.py:232-262: `_circuit_from_yaml` -> `get_run_func`).  Must precede the first rp_forward of the plan. */
int rp_plan_set_jit_module(rp_plan* plan, const void* image, long long nbytes);

int rp_forward(rp_plan* plan, const rp_fwd_args* args, void* stream);
int rp_backward(rp_plan* plan, const rp_bwd_args* args, void* stream);
/* Synchronises `stream` and reports whether a range guard of the plan tripped since the last call (RP_PREC_3XF16: the adjoint
 * grew faster than the binary16 operand range allows within one weight-gradient chunk).  0: results are valid. */
int rp_plan_status(rp_plan* plan, void* stream);

/* Sequential recursive-least-squares readout training over a recorded state matrix (edges.py:227-234):
 * for t in [0,T): y_hat = W x_t ; z = beta_inv*P x_t ; kappa = 1/(1+x_t.z) ;
 *                 W += outer(y_t - kappa * (W + outer(y_t,z)) x_t , z) ; P -= kappa*outer(z,z) ; loss_t = |y_t-y_hat|^2
 * X [T,n_in], Y [T,n_out], W [n_out,n_in] in/out, P [n_in,n_in] in/out, loss [T] out (or NULL), pred [T,n_out] out (or NULL). */
int rp_rls_run(int T, int n_in, int n_out, float beta_inv, const float* X, const float* Y,
               float* W, float* P, float* loss, float* pred, int update_every, void* stream);

/* Profiling hook: launch the plan's dominant contraction (0: forward W.src, 1: adjoint W^T.g, 2: weight gradient over a
 * full K chunk) `iters` times back to back on `stream`, bracketed by CUDA events; *avg_ms receives the mean launch
 * duration, *flops the algorithmic (logical, not x3) flop count of one launch.  Operands are whatever the plan's
 * workspaces currently hold (timing is data independent).  Synchronises the stream. */
int rp_plan_time_contraction(rp_plan* plan, int which, int iters, float* avg_ms, double* flops, void* stream);

/* Profiling hook: per-stage device time of the calls as they actually run.  rp_plan_stage_timing(plan, 1) makes rp_forward /
 * rp_backward record a CUDA event on the launching stream before every launch of their step loops; rp_plan_stage_times
 * synchronises, adds the event-to-event intervals up per stage and disarms nothing (call rp_plan_stage_timing(plan, 0) for that).
 * ms[RP_NUM_STAGES] / marks[RP_NUM_STAGES]: 0 fused forward step (contraction + element-wise epilogue), 1 adjoint product
 * Z = (kW)^T g, 2 weight-gradient chunk, 3 reverse element-wise kernel(s), 4 everything else between the first and last mark
 * (per-call kernels, and whatever the caller enqueued between the two calls).  marks = intervals summed = launches of that stage. */
int rp_plan_stage_timing(rp_plan* plan, int enable);
int rp_plan_stage_times(rp_plan* plan, float* ms, int* marks, void* stream);

/* Debug / profiling aid: per-CTA timeline of the contraction and adjoint kernels.  rp_trace_enable(capacity > 0) arms a device
 * buffer of `capacity` records (0 disarms and frees it); rp_trace_read synchronises the device, copies up to max_records records
 * {uint32 tag, uint32 smid, uint64 start_ns, uint64 end_ns} (globaltimer) to host_buf and returns how many there are. */
int rp_trace_enable(int capacity);
int rp_trace_read(void* host_buf, int max_records);

/* Standalone GEMM used by the engine, exposed for testing:  C[q*ldc+p] (+)= sum_k A[p*lda+k]*B[q*ldb+k]
 * precision RP_PREC_FP32 -> FFMA kernel, RP_PREC_3XTF32 / RP_PREC_3XF16 -> tcgen05 kernel (needs p,q,k extents it supports). */
int rp_gemm_tn(int precision, int P, int Q, int K, const float* A, int lda, const float* B, int ldb,
               float* C, int ldc, int accumulate, void* stream);
#endif /* __CUDACC_RTC__ */

#ifdef __cplusplus
}
#endif
#endif /* RECTIPY_B200_H */
