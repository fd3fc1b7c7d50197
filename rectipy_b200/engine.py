"""Host side of the engine: plan cache and the whole-horizon autograd function over the C ABI.

`EngineRun.apply` is what `Network.run` / `Network.forward` / `fit_bptt` dispatch to.  One call integrates T Euler
steps on the device (rp_forward) and -- when gradients are required -- keeps the per-step state checkpoints so that
`backward` can run the fused reverse-time adjoint (rp_backward) instead of torch autograd over an unrolled Python
loop (reference: rectipy/network.py:588-599,1123-1130).
"""
from __future__ import annotations

import ctypes as C
import os
from dataclasses import dataclass
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import _cabi as abi


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


_raw_stream = getattr(torch._C, "_cuda_getCurrentRawStream", None)


def _stream():
    """cudaStream_t of torch's current stream on the current device (the raw getter is ~20x cheaper than building a Stream object,
    which matters for single-step calls: user loops over `forward`, FeedbackNetwork)."""
    if _raw_stream is not None:
        return C.c_void_p(_raw_stream(torch.cuda.current_device()))
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _f32c(t: torch.Tensor) -> torch.Tensor:
    if t.dtype != torch.float32:
        t = t.to(torch.float32)
    return t.contiguous()


@dataclass(frozen=True)
class PlanKey:
    model: int
    n: int
    batch: int
    in_mode: int
    n_in: int
    in_target: int
    out_mode: int
    n_out: int
    out_var: int
    precision: int
    dt: float
    theta: float
    v_reset: float
    slope: float
    per_neuron: Tuple[int, ...]
    device: int
    jit_key: str = ""        # RP_JIT: key of the compiled program (rectipy_b200.jit) and what rp_desc.jit_* carry
    jit_nsv: int = 0
    jit_spiking: int = 0
    jit_post_out: int = 0
    jit_src_plane: int = 0


class Plan:
    """Owns one `rp_plan` (device workspaces).  Not re-entrant; one stream at a time."""

    def __init__(self, key: PlanKey):
        self.key = key
        self.lib = abi.load()
        if not torch.cuda.is_available():
            raise RuntimeError("rectipy_b200: no CUDA device available; the engine has no CPU fallback")
        d = abi.rp_desc()
        d.model, d.n, d.batch = key.model, key.n, key.batch
        d.in_mode, d.n_in, d.in_target = key.in_mode, key.n_in, key.in_target
        d.out_mode, d.n_out, d.out_var = key.out_mode, key.n_out, key.out_var
        d.precision = key.precision
        d.dt, d.theta, d.v_reset, d.slope = key.dt, key.theta, key.v_reset, key.slope
        for i, v in enumerate(key.per_neuron):
            d.param_per_neuron[i] = v
        d.jit_nsv, d.jit_spiking, d.jit_post_out, d.jit_src_plane = key.jit_nsv, key.jit_spiking, key.jit_post_out, key.jit_src_plane
        handle = C.c_void_p()
        with torch.cuda.device(key.device):
            abi.check(self.lib.rp_plan_create(C.byref(d), C.byref(handle)), "rp_plan_create")
            self.handle = handle
            if key.model == abi.RP_JIT:
                from . import jit
                image = jit.program(key.jit_key).image
                abi.check(self.lib.rp_plan_set_jit_module(handle, image, len(image)), "rp_plan_set_jit_module")
        self.nsv = key.jit_nsv if key.model == abi.RP_JIT else int(self.lib.rp_num_state_vars(key.model))
        self.nh = key.jit_nsv + 1 if key.model == abi.RP_JIT else int(self.lib.rp_num_history_planes(key.model))

    def __del__(self):
        try:
            if getattr(self, "handle", None):
                self.lib.rp_plan_destroy(self.handle)
                self.handle = None
        except Exception:
            pass

    @property
    def launches(self) -> int:
        return int(self.lib.rp_plan_launch_count(self.handle))

    def time_contraction(self, which: int, iters: int = 20):
        """(avg launch ms, algorithmic flops per launch) of the plan's contraction kernel: 0 fwd, 1 adjoint, 2 wgrad."""
        ms, fl = C.c_float(), C.c_double()
        with torch.cuda.device(self.key.device):
            abi.check(self.lib.rp_plan_time_contraction(self.handle, which, iters, C.byref(ms), C.byref(fl), _stream()),
                      "rp_plan_time_contraction")
        return float(ms.value), float(fl.value)

    def stage_timing(self, enable: bool) -> None:
        """Arm / disarm the per-stage CUDA-event timing of rp_forward / rp_backward (profiling passes only)."""
        abi.check(self.lib.rp_plan_stage_timing(self.handle, int(enable)), "rp_plan_stage_timing")

    def stage_times(self):
        """{stage: (total ms, launches)} since timing was armed (synchronises the current stream)."""
        ms = (C.c_float * abi.RP_NUM_STAGES)()
        marks = (C.c_int * abi.RP_NUM_STAGES)()
        with torch.cuda.device(self.key.device):
            abi.check(self.lib.rp_plan_stage_times(self.handle, ms, marks, _stream()), "rp_plan_stage_times")
        return {name: (float(ms[i]), int(marks[i])) for i, name in enumerate(abi.STAGE_NAMES)}

    @property
    def workspace_bytes(self) -> int:
        return int(self.lib.rp_plan_workspace_bytes(self.handle))


_PLANS: Dict[PlanKey, Plan] = {}


def get_plan(key: PlanKey) -> Plan:
    plan = _PLANS.get(key)
    if plan is None:
        plan = Plan(key)
        _PLANS[key] = plan
    return plan


def clear_plans():
    _PLANS.clear()


def total_launches() -> int:
    return sum(p.launches for p in _PLANS.values())


def tc_supported(n: int, batch: int) -> bool:
    """Shapes the tcgen05 split-3 kernels (3xTF32 / 3xFP16) accept (rp_gemm_tc.cuh: 128-row tiles on both operands)."""
    return n >= 128 and batch >= 128 and n % 128 == 0 and batch % 128 == 0


def num_records(T: int, sampling_steps: int, cutoff: int) -> int:
    return int(abi.load().rp_num_records(T, sampling_steps, cutoff))


@dataclass
class RunConfig:
    T: int
    sampling_steps: int
    cutoff: int
    truncate_steps: int
    rec_vars: Tuple[int, ...]         # state-variable indices
    rec_reduce: Tuple[int, ...]
    want_out: bool
    param_slots: Tuple[int, ...]      # abi slot of each tensor in *params (same order)


def history_budget_bytes() -> int:
    """Checkpoint memory one autograd call may hold (env RECTIPY_B200_HISTORY_GB, default 64 GiB of the 180 GB HBM).
    Longer horizons are integrated in segments: only segment-boundary states are kept and each segment's per-step
    checkpoints are recomputed just before its reverse sweep."""
    return int(float(os.environ.get("RECTIPY_B200_HISTORY_GB", "64")) * (1 << 30))


def plan_segments(T: int, S: int, slot_bytes: int, budget: int) -> List[Tuple[int, int]]:
    """[(t_offset, length)] covering [0, T).  One segment if the full history fits the budget; otherwise segment starts
    sit right after a record step ((t_offset - 1) % S == 0) so that no Observer window is split."""
    if T <= 0 or (T + 1) * slot_bytes <= budget:
        return [(0, T)]
    steps = max(1, budget // slot_bytes - 2)
    L = max(S, (steps // S) * S)
    segs, t0 = [], 0
    first = min(T, L + 1)
    segs.append((0, first))
    t0 = first
    while t0 < T:
        n = min(L, T - t0)
        segs.append((t0, n))
        t0 += n
    return segs


def _fill_common(a, cfg, x_c, W_c, W_in_c, W_out_c, params_c, t0, n, T_total):
    a.T, a.sampling_steps, a.cutoff = n, cfg.sampling_steps, cfg.cutoff
    a.t_offset, a.T_total = t0, T_total
    a.x = _ptr(x_c[t0:t0 + n]) if x_c is not None else None
    a.W, a.W_in, a.W_out = _ptr(W_c), _ptr(W_in_c), _ptr(W_out_c)
    for slot, p in zip(cfg.param_slots, params_c):
        a.params[slot] = p.data_ptr()


class EngineRun(torch.autograd.Function):
    """(x, W, W_in, W_out, y0, *params) -> (out_rec, yT, *recorded_vars) for a whole horizon."""

    @staticmethod
    def forward(ctx, plan: Plan, cfg: RunConfig, x, W, W_in, W_out, y0, *params):
        key = plan.key
        lib = plan.lib
        dev = y0.device
        B, N = key.batch, key.n
        nh = plan.nh
        n_rec = lib.rp_num_records(cfg.T, cfg.sampling_steps, cfg.cutoff)
        # grad mode is off inside Function.forward; needs_input_grad already accounts for no_grad() at apply time
        needs_grad = any(ctx.needs_input_grad)

        x_c = None if x is None else _f32c(x)
        W_c, y0_c = _f32c(W), _f32c(y0)
        W_in_c = None if W_in is None else _f32c(W_in)
        W_out_c = None if W_out is None else _f32c(W_out)
        params_c = [_f32c(p) for p in params]

        out_w = key.n_out if key.out_mode == abi.RP_OUT_READOUT else N
        out_rec = torch.empty((n_rec, B, out_w), device=dev, dtype=torch.float32) if cfg.want_out else None
        recs = [torch.empty((n_rec, B) if red else (n_rec, B, N), device=dev, dtype=torch.float32) for red in cfg.rec_reduce]
        segs = plan_segments(cfg.T, cfg.sampling_steps, nh * B * N * 4, history_budget_bytes()) if needs_grad else [(0, cfg.T)]
        segmented = len(segs) > 1
        history = None
        if needs_grad and not segmented:
            history = torch.empty((cfg.T + 1, nh, B, N), device=dev, dtype=torch.float32)
        bounds = []
        y_cur = y0_c
        with torch.cuda.device(dev):
            for (t0, n) in segs:
                a = abi.rp_fwd_args()
                _fill_common(a, cfg, x_c, W_c, W_in_c, W_out_c, params_c, t0, n, cfg.T)
                yT = torch.empty_like(y0_c)
                a.y0, a.yT = _ptr(y_cur), _ptr(yT)
                a.out_rec = _ptr(out_rec)
                a.n_rec_vars = len(cfg.rec_vars)
                for i, (v, red) in enumerate(zip(cfg.rec_vars, cfg.rec_reduce)):
                    a.rec_var[i], a.rec_reduce[i], a.rec_buf[i] = v, int(red), recs[i].data_ptr()
                a.history = _ptr(history)
                abi.check(lib.rp_forward(plan.handle, C.byref(a), _stream()), "rp_forward")
                if segmented:
                    bounds.append(y_cur)
                y_cur = yT
        yT = y_cur

        ctx.plan, ctx.cfg = plan, cfg
        ctx.has = (x is not None, W_in is not None, W_out is not None)
        ctx.segs = segs
        ctx.param_shapes = [tuple(p.shape) for p in params]
        if needs_grad:
            keep = history if not segmented else torch.stack(bounds)      # [n_seg, nsv, B, N] segment-start states
            saved = [t for t in (x_c, W_c, W_in_c, W_out_c, keep) if t is not None] + params_c
            ctx.save_for_backward(*saved)
        if out_rec is None:
            out_rec = torch.empty((0,), device=dev)
            ctx.mark_non_differentiable(out_rec)
        ctx.mark_non_differentiable(*recs)
        return (out_rec, yT) + tuple(recs)

    @staticmethod
    def backward(ctx, g_out, g_yT, *g_recs):
        plan, cfg = ctx.plan, ctx.cfg
        key, lib = plan.key, plan.lib
        B, N = key.batch, key.n
        nsv, nh = plan.nsv, plan.nh
        saved = list(ctx.saved_tensors)
        has_x, has_win, has_wout = ctx.has
        x_c = saved.pop(0) if has_x else None
        W_c = saved.pop(0)
        W_in_c = saved.pop(0) if has_win else None
        W_out_c = saved.pop(0) if has_wout else None
        keep = saved.pop(0)
        params_c = saved
        dev = W_c.device
        segs = ctx.segs
        segmented = len(segs) > 1
        # needs_input_grad indices: plan, cfg, x, W, W_in, W_out, y0, *params
        need = ctx.needs_input_grad
        need_x, need_W, need_Win, need_Wout, need_y0 = need[2], need[3], need[4], need[5], need[6]
        need_p = need[7:]
        if need_x and has_x and key.in_mode != abi.RP_IN_DENSE:
            raise NotImplementedError("rectipy_b200: gradients w.r.t. projected inputs (RP_IN_PROJ) are not provided; "
                                      "detach the input or use a dense input current")

        g_out_c = _f32c(g_out) if (g_out is not None and cfg.want_out) else None
        g_state = _f32c(g_yT) if g_yT is not None else None
        g_x = torch.empty_like(x_c) if (need_x and has_x) else None
        totals = {}

        def accumulate(name, t):
            if t is None:
                return
            if name in totals:
                totals[name].add_(t)
            else:
                totals[name] = t

        with torch.cuda.device(dev):
            for si in range(len(segs) - 1, -1, -1):
                t0, n = segs[si]
                if segmented:
                    # recompute this segment's per-step checkpoints from its start state
                    history = torch.empty((n + 1, nh, B, N), device=dev, dtype=torch.float32)
                    f = abi.rp_fwd_args()
                    _fill_common(f, cfg, x_c, W_c, W_in_c, W_out_c, params_c, t0, n, cfg.T)
                    scratch = torch.empty((nsv, B, N), device=dev, dtype=torch.float32)
                    f.y0, f.yT, f.history = _ptr(keep[si]), _ptr(scratch), _ptr(history)
                    abi.check(lib.rp_forward(plan.handle, C.byref(f), _stream()), "rp_forward (recompute)")
                else:
                    history = keep
                b = abi.rp_bwd_args()
                _fill_common(b, cfg, x_c, W_c, W_in_c, W_out_c, params_c, t0, n, cfg.T)
                b.truncate_steps = cfg.truncate_steps
                b.history = _ptr(history)
                b.g_out_rec, b.g_yT = _ptr(g_out_c), _ptr(g_state)
                dW = torch.empty_like(W_c) if need_W else None
                dW_in = torch.empty_like(W_in_c) if (need_Win and has_win) else None
                dW_out = torch.empty_like(W_out_c) if (need_Wout and has_wout) else None
                g_y0 = torch.empty((nsv, B, N), device=dev, dtype=torch.float32) if (need_y0 or si > 0) else None
                b.dW, b.dW_in, b.dW_out, b.g_y0 = _ptr(dW), _ptr(dW_in), _ptr(dW_out), _ptr(g_y0)
                b.g_x = _ptr(g_x[t0:t0 + n]) if g_x is not None else None
                dps = []
                for slot, np_ in zip(cfg.param_slots, need_p):
                    buf = torch.empty((N,), device=dev, dtype=torch.float32) if np_ else None
                    if buf is not None:
                        b.dparams[slot] = buf.data_ptr()
                    dps.append(buf)
                abi.check(lib.rp_backward(plan.handle, C.byref(b), _stream()), "rp_backward")
                if key.precision == abi.RP_PREC_3XF16:      # binary16 operand range guard (synchronises the stream)
                    abi.check(lib.rp_plan_status(plan.handle, _stream()), "rp_backward")
                accumulate("W", dW), accumulate("W_in", dW_in), accumulate("W_out", dW_out)
                for i, buf in enumerate(dps):
                    accumulate(("p", i), buf)
                g_state = g_y0
                del history
        grads_p = []
        for i, shape in enumerate(ctx.param_shapes):
            buf = totals.get(("p", i))
            if buf is None:
                grads_p.append(None)
            else:
                numel = 1
                for sdim in shape:
                    numel *= sdim
                grads_p.append(buf.reshape(shape) if numel == N and N > 1 else buf.sum().reshape(shape))
        return (None, None, g_x, totals.get("W"), totals.get("W_in"), totals.get("W_out"),
                g_state if need_y0 else None) + tuple(grads_p)


def rls_run(X: torch.Tensor, Y: torch.Tensor, W: torch.Tensor, P: torch.Tensor, beta_inv: float,
            update_every: int = 1) -> Tuple[torch.Tensor, torch.Tensor]:
    """Sequential RLS over recorded states (edges.py:227-234); updates W [k,n] and P [n,n] in place."""
    lib = abi.load()
    T, n = X.shape
    k = Y.shape[1]
    assert W.shape == (k, n) and P.shape == (n, n) and W.is_contiguous() and P.is_contiguous()
    assert W.dtype == torch.float32 and P.dtype == torch.float32
    X, Y = _f32c(X), _f32c(Y)
    loss = torch.empty((T,), device=X.device, dtype=torch.float32)
    pred = torch.empty((T, k), device=X.device, dtype=torch.float32)
    with torch.cuda.device(X.device):
        abi.check(lib.rp_rls_run(T, n, k, float(beta_inv), _ptr(X), _ptr(Y), _ptr(W), _ptr(P), _ptr(loss), _ptr(pred),
                                 int(update_every), _stream()), "rp_rls_run")
    return loss, pred


def gemm_tn(A: torch.Tensor, Bm: torch.Tensor, precision: int = abi.RP_PREC_FP32, out: Optional[torch.Tensor] = None,
            accumulate: bool = False) -> torch.Tensor:
    """C[q, p] (+)= sum_k A[p, k] * B[q, k]  through the engine's contraction kernels (test hook)."""
    lib = abi.load()
    A, Bm = _f32c(A), _f32c(Bm)
    P_, K = A.shape
    Q = Bm.shape[0]
    if out is None:
        out = torch.zeros((Q, P_), device=A.device, dtype=torch.float32)
    with torch.cuda.device(A.device):
        abi.check(lib.rp_gemm_tn(precision, P_, Q, K, _ptr(A), A.stride(0), _ptr(Bm), Bm.stride(0), _ptr(out), out.stride(0),
                                 int(accumulate), _stream()), "rp_gemm_tn")
    return out
