#!/usr/bin/env python
"""Benchmark of the RectiPy time-stepping hot path on B200 (driver contract: see the task description).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

Workload (BASELINE.json configs[2], the configuration the metric is quoted on):
    recurrent QIF spiking network, N = 4096 neurons, batch = 1024 independent trials per GPU, m = 2 inputs, k = 3
    readouts, dt = 1e-3, BPTT with the surrogate spike gradient, training W and W_out, MSE loss.
One bench "step" = one BPTT pass over one batch: forward T_INNER Euler steps + the reverse-time adjoint
(= one `fit_bptt` epoch without the optimizer update).  metric = neuron-steps/s = N * B * T_INNER / seconds.

    value        inputs already resident in HBM, engine called through the autograd function (device tensors)
    e2e          same pass through the public API (`Network.run` + loss + `backward`) with HOST inputs/targets
                 (pinned), host->device copies and the device->host read of the loss inside the timed region
    fwd          extra: forward-only neuron-steps/s (Network.run, no grad) over the same shapes
    roofline     dominant kernel = tcgen05 split-3 contraction (binary16 words); achieved = logical flops/launch / CUDA-event time
    cpu_baseline the CPU oracle port (= the reference's eager-torch path restated) on this box's host cores

--impl reference times the reference's CPU algorithm (oracle port; the reference itself is a Python package that
cannot be installed offline because its vector fields come from the un-pinned third-party PyRates) on the same
metric/config: one unbatched trial (the reference has no trial axis) of a bounded number of steps per bench step.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_NEURONS, BATCH, N_IN, N_OUT, DT = 4096, 1024, 2, 3, 1e-3
T_INNER = 100
PRECISION = "auto"    # tensor-core operand format: auto -> 3xf16 (binary16 split words), or 3xtf32 / fp32
METRIC = "neuron-steps/sec (QIF N=4096, batch 1024, BPTT fwd+bwd)"
UNIT = "neuron-steps/s"
CPU_T = 40            # Euler steps per CPU-baseline sample (one trial, BPTT)


def _peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as fh:
            p = json.load(fh)
        return dict(bf16=p["bf16_tflops"], bf16_sustained=p["bf16_tflops_sustained"], hbm=p["hbm_gbs"], source="measured")
    return dict(bf16=1590.0, bf16_sustained=1400.0, hbm=6650.0, source="fallback")


def _ncu_traffic():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the dominant kernel, from the committed ncu --set full
    capture (profiles/ncu_traffic.json, written by tools/summarize_ncu.py); None when no capture is recorded."""
    path = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        with open(path) as fh:
            return json.load(fh)["dram_bytes_per_launch"]
    except (OSError, KeyError, ValueError):
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    def __init__(self, gpu_index: int):
        self.gpu = gpu_index
        self.proc = None
        self.lines = []

    def start(self):
        q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except OSError:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return dict(sm_mhz=None, sm_max_mhz=None, reasons=["nvidia-smi unavailable"])
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            pass
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0])); mx.append(float(f[1]))
            except ValueError:
                continue
            for nm, val in zip(names, f[3:7]):
                if val.lower().startswith("active"):
                    reasons.add(nm)
        sm.sort()
        return dict(sm_mhz=sm[len(sm) // 2] if sm else None, sm_max_mhz=max(mx) if mx else None,
                    reasons=sorted(reasons), samples=len(sm))


def make_problem(seed: int, n: int, batch: int, T: int):
    import numpy as np
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    rng = np.random.default_rng(seed)
    W = (2.0 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    w_in = rng.standard_normal((n, N_IN)).astype(np.float32)
    w_out = (rng.standard_normal((N_OUT, n)) / np.sqrt(n)).astype(np.float32)
    etas = (-5.0 + np.tan((np.pi / 2) * (2.0 * np.arange(1, n + 1) - n - 1) / (n + 1))).astype(np.float32)
    t = np.arange(T, dtype=np.float32) * DT
    amp = rng.uniform(5, 15, (1, batch, 1)).astype(np.float32)
    phase = rng.uniform(0, 2 * np.pi, (1, batch, N_IN)).astype(np.float32)
    omega = np.asarray([3.0, 5.0], dtype=np.float32)[None, None, :]
    x = (amp * np.sin(2 * np.pi * omega * t[:, None, None] + phase) + 8.0).astype(np.float32)
    targets = rng.standard_normal((T, batch, N_OUT)).astype(np.float32)
    return W, w_in, w_out, etas, x, targets


def spread_state(seed: int, n: int, batch: int):
    """Initial state of a network that is already active: membrane potentials spread over the cycle (v in [-50, 99), s = 0),
    so that threshold crossings, resets and surrogate-gradient terms all occur within the short benchmark horizon
    (from the template's rest state v = -2 the first spikes only appear after ~500 steps)."""
    import numpy as np
    rng = np.random.default_rng(seed)
    return np.concatenate([rng.uniform(-50.0, 99.0, (batch, n)), np.zeros((batch, n))], axis=1).astype(np.float32)


def build_network(W, w_in, w_out, etas, batch, device):
    import rectipy_b200 as rp
    net = rp.Network(DT, device=device, batch=batch, precision=PRECISION)
    node = net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=W, source_var="s",
                               target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v",
                               op="qif_op", node_vars={"eta": etas}, train_params=["weights"])
    net.add_func_node("inp", N_IN, "identity")
    net.add_edge("inp", "qif", weights=w_in)
    net.add_func_node("out", N_OUT, "identity")
    net.add_edge("qif", "out", weights=w_out, train="gd")
    net.compile()
    return net, node


# ---------------------------------------------------------------------------------------------------------------
def cpu_reference_rate(n: int, T: int, reps: int, seed: int = 0):
    """Reference algorithm on the host cores: oracle port (eager torch, one trial), run(enable_grad) + loss.backward()."""
    import numpy as np
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rectipy_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W, w_in, w_out, etas, x, targets = make_problem(seed, n, 1, T)
    times = []
    for rep in range(reps + 1):
        node = orc.make_node("qif", n, W, DT, params=dict(eta=etas), dtype=torch.float32, train_params=["weights"],
                             y0=spread_state(seed, n, 1)[0])
        net = orc.OracleNet(node, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out, requires_grad=True))
        xt, tg = torch.tensor(x[:, 0, :]), torch.tensor(targets[:, 0, :])
        t0 = time.perf_counter()
        res = net.run(xt, sampling_steps=1, enable_grad=True)
        loss = torch.nn.functional.mse_loss(torch.stack(res["out"]), tg)
        loss.backward()
        dt_ = time.perf_counter() - t0
        if rep > 0:
            times.append(dt_)
    times.sort()
    med = times[len(times) // 2]
    return n * T / med, cores, med


def cpu_batched_rate(n: int, B: int, T: int, reps: int, seed: int = 0):
    """A stronger CPU comparator than the reference itself (which has no trial axis): the same QIF BPTT pass restated with a
    trial axis, so that the recurrent products become [B,N]x[N,N] sgemm on all host cores.  Not reference code."""
    import torch
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import rectipy_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    W, w_in, w_out, etas, x, targets = make_problem(seed, n, B, T)
    y0 = spread_state(seed, n, B)
    theta, v_reset = torch.tensor(100.0), torch.tensor(-100.0)
    slope, center = torch.tensor(0.5), torch.tensor(1.0)
    times = []
    for rep_ in range(reps + 1):
        Wt = torch.tensor(W, requires_grad=True); Wo = torch.tensor(w_out, requires_grad=True)
        Wi, eta = torch.tensor(w_in), torch.tensor(etas)
        v, sv = torch.tensor(y0[:, :n]), torch.tensor(y0[:, n:])
        xt, tg = torch.tensor(x), torch.tensor(targets)
        t0 = time.perf_counter()
        outs = []
        for t in range(T):
            spk = orc.OracleSpike.apply(v - theta, slope, center)
            gate = spk.detach()
            outs.append(sv @ Wo.T)
            vt = v + DT * ((v * v + eta + xt[t] @ Wi.T) / 1.0 + 1.0 * (sv @ Wt.T))
            sv = sv + DT * (-sv / 1.0 + spk / DT)
            v = vt * (1.0 - gate) + gate * v_reset
        loss = torch.nn.functional.mse_loss(torch.stack(outs), tg)
        loss.backward()
        dt_ = time.perf_counter() - t0
        if rep_ > 0:
            times.append(dt_)
    times.sort()
    med = times[len(times) // 2]
    return n * B * T / med, cores, med


def run_reference(args):
    """`--impl reference`: rank 0 only."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    times = []
    import torch
    for i in range(args.warmup + args.steps):
        rate, cores, sec = cpu_reference_rate(N_NEURONS, CPU_T, 1, seed=i)
        if i >= args.warmup:
            times.append(sec)
    total = sum(times)
    value = N_NEURONS * CPU_T * len(times) / total
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * total / len(times), "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"QIF N={N_NEURONS} BPTT fwd+bwd, reference algorithm (eager torch, unbatched) on host CPU",
                   "n": N_NEURONS, "batch": 1, "t_inner": CPU_T, "n_in": N_IN, "n_out": N_OUT, "dt": DT},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": os.cpu_count(), "kind": "port",
                         "sample": f"1 trial x {CPU_T} Euler steps BPTT per step (reference has no trial axis)"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    _OUT.emit(json.dumps(line))


# ---------------------------------------------------------------------------------------------------------------
def run_ours(args):
    import numpy as np
    import torch
    import torch.distributed as dist
    from rectipy_b200 import engine, parallel, _cabi as abi

    if os.environ.get("NCCL_DEBUG", "VERSION").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"          # keep stdout to the single JSON line
    rank, local_rank, world = parallel.init_from_env("nccl")
    if not torch.cuda.is_available():
        raise RuntimeError("bench.py needs a CUDA device (the engine has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    device = f"cuda:{local_rank}"
    n, B, T = N_NEURONS, BATCH, T_INNER
    W, w_in, w_out, etas, x_np, tgt_np = make_problem(1234 + rank, n, B, T)   # every rank owns its own trials
    W, w_in, w_out, etas = make_problem(1234, n, 1, 1)[:4]                    # parameters are replicated
    net, node = build_network(W, w_in, w_out, etas, B, device)
    node.reset(spread_state(4321 + rank, n, B))
    edge_out = net.get_edge("qif", "out")
    params = [node["weights"], edge_out.weights]
    x_dev = torch.tensor(x_np, device=device)
    tgt_dev = torch.tensor(tgt_np, device=device)
    x_host = torch.tensor(x_np).pin_memory()
    tgt_host = torch.tensor(tgt_np).pin_memory()
    y0 = net.state

    def barrier():
        if world > 1:
            dist.barrier(device_ids=[local_rank])
        torch.cuda.synchronize()

    def step_device():
        net.reset(y0)
        for p in params:
            p.grad = None
        obs = net.run(x_dev, sampling_steps=1, verbose=False, enable_grad=True)
        loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt_dev)
        loss.backward()
        net.allreduce_gradients()        # the public hook fit_bptt itself uses before every optimizer step (no-op on one rank)
        return loss

    def step_e2e():
        net.reset(y0)
        for p in params:
            p.grad = None
        obs = net.run(x_host, sampling_steps=1, verbose=False, enable_grad=True)      # H2D inside run()
        tg = tgt_host.to(device, non_blocking=True)
        loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]), tg)
        loss.backward()
        net.allreduce_gradients()
        return float(loss.item())                                                        # D2H of the result

    def step_fwd():
        net.reset(y0)
        net.run(x_dev, sampling_steps=1, verbose=False, enable_grad=False)

    def timed(fn, steps, warmup):
        for _ in range(warmup):
            fn()
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        l0 = engine.total_launches()
        e0.record()
        for _ in range(steps):
            fn()
        e1.record()
        barrier()
        ms = torch.tensor([e0.elapsed_time(e1)], device=device)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item()), engine.total_launches() - l0

    # CPU-side group for the final wait: ranks > 0 must not spin inside an NCCL barrier kernel while rank 0 runs its CPU baseline
    cpu_group = dist.new_group(backend="gloo") if world > 1 else None
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    ms_dev, launches = timed(step_device, args.steps, args.warmup)
    clocks = sampler.stop() if rank == 0 else None
    ms_e2e, _ = timed(step_e2e, args.steps, max(1, args.warmup // 2))
    ms_fwd, _ = timed(step_fwd, args.steps, 1)
    # one more pass with a CUDA event before every launch of the step loops: device time per stage of the pass AS IT RUNS
    plan0 = next(iter(engine._PLANS.values()))
    plan0.stage_timing(True)
    step_device()
    stages = plan0.stage_times()
    plan0.stage_timing(False)

    work = float(n) * B * T * args.steps * world
    value = work / (ms_dev * 1e-3)
    e2e_value = work / (ms_e2e * 1e-3)
    fwd_value = work / (ms_fwd * 1e-3)

    if rank == 0:
        peaks = _peaks()
        plan = next(iter(engine._PLANS.values()))
        ms_f, fl_f = plan.time_contraction(0, 20)
        ms_b, fl_b = plan.time_contraction(1, 20)
        ms_w, fl_w = plan.time_contraction(2, 10)
        use_tc = plan.key.precision in (abi.RP_PREC_3XTF32, abi.RP_PREC_3XF16)
        f16 = plan.key.precision == abi.RP_PREC_3XF16
        # measured TF32 dense peak of this box: cuBLAS (torch.matmul, allow_tf32) 8192^3, best of 5 -- MEASURED_PEAKS.json only
        # carries bf16; BASELINE.md asks for the tf32 figure to be measured in the run
        torch.backends.cuda.matmul.allow_tf32 = True
        ma = torch.randn(8192, 8192, device=device); mb = torch.randn(8192, 8192, device=device)
        best = 1e9
        for _ in range(6):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); torch.matmul(ma, mb); e1.record(); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1))
        tf32_measured = 2.0 * 8192 ** 3 / (best * 1e-3) / 1e12
        torch.backends.cuda.matmul.allow_tf32 = False
        del ma, mb
        # tf32 tensor rate is half the bf16 rate; 3xTF32 issues 3 MMAs per logical product -> divide by 3 again.
        # These launches are timed inside a long, power-capped step -> compare with the sustained figure.
        # the contraction is timed alone (20 back-to-back launches, ~3 ms) -> burst figure; tf32 issues at half the bf16 rate
        # binary16 split words issue at the full 16-bit rate: peak = measured bf16 burst / 3
        if f16:
            peak_logical = peaks["bf16"] / 3.0
        else:
            peak_logical = max(peaks["bf16"] / 2.0, tf32_measured) / 3.0 if use_tc else 72.0
        ach = fl_f / (ms_f * 1e-3) / 1e12
        # share of one BPTT pass spent in the three contractions (per Euler step: 1 fwd + 1 adjoint + 1/chunk wgrad)
        chunk_steps = fl_w / (2.0 * n * n * B)
        pass_ms = ms_dev / args.steps
        # ---- per-stage entries from the event-timed pass (the kernels that actually run, gaps included) -----------------------
        hbm = peaks["hbm"]
        f_fwd, f_bptt = 2.0 * n + 2 * N_IN + 2 * N_OUT + 14, 6.0 * n + 60          # flops per neuron-step
        adj_bytes = 44.0                                                              # B per neuron-step of the fused reverse kernel
        fwd_cnt = stages.get("fwd_fused", (0.0, 0))[1]
        persist = f16 and 0 < fwd_cnt < T                     # one cooperative launch integrates the whole horizon
        cg2 = f16 and not os.environ.get("RP_NO_TC_CG2")      # adjoint product / weight gradient on 256x256 CTA-pair tiles (cta_group::2)
        gemm = "rp::k_gemm_split3_cg2<EpiStore> (256x256 CTA-pair tile, cta_group::2)" if cg2 else "rp::k_gemm_split3<256, EpiStore, f16=%d>" % int(f16)
        fwd_cg2 = persist and not (os.environ.get("RP_NO_FWD_CG2") or os.environ.get("NV_NSIGHT_INJECTION_TRANSPORT_TYPE") or os.environ.get("CUDA_INJECTION64_PATH"))
        names = {"fwd_fused": (("rp::k_gemm_fwd_persist_cg2<EpiFwd<QIF>> (CTA pairs, " if fwd_cg2 else "rp::k_gemm_fwd_persist<256, EpiFwd<QIF>, f16> (") +
                               "persistent over all T steps: W.s contraction + Euler step + readout per step)"
                               if persist else "rp::k_gemm_split3<256, EpiFwd<QIF>, f16=%d> (W.s contraction + Euler step + readout)" % int(f16)),
                 "dgrad": gemm + " (Z = (kW)^T g)",
                 "wgrad": gemm + " (dW += g (x) s, K = %d steps x batch)" % int(chunk_steps),
                 "adjoint_elementwise": "rp::k_adj_fused_f16<QIF> (adjoint step + split operands + dW_out partials)",
                 "other": "per-call kernels (weight split, maxima, dW finish) + the loss between the two calls"}
        stage_total = sum(v[0] for v in stages.values()) or 1.0
        kernels = {}
        for key, (ms_tot, cnt) in stages.items():
            ent = {"kernel": names[key], "launches": cnt, "ms_total": ms_tot, "ms_per_launch": ms_tot / max(cnt, 1), "share_of_pass": ms_tot / stage_total}
            if key in ("fwd_fused", "dgrad") and cnt:
                steps_cov = T if (key == "fwd_fused" and persist) else cnt       # Euler steps these launches cover
                ent["steps_covered"] = steps_cov
                ent["ms_per_step_of_stage"] = ms_tot / steps_cov
                tf = 2.0 * n * n * B / (ms_tot / steps_cov * 1e-3) / 1e12
                ent.update(bound="tensor", achieved_tflops=tf, frac_of_burst_peak=tf / peak_logical, frac_of_sustained_peak=tf / (peaks["bf16_sustained"] / 3.0))
            elif key == "wgrad" and cnt:
                steps_cov = T          # all chunks of the pass together cover T steps
                tf = 2.0 * n * n * B * steps_cov / (ms_tot * 1e-3) / 1e12
                ent.update(bound="tensor", achieved_tflops=tf, frac_of_burst_peak=tf / peak_logical, frac_of_sustained_peak=tf / (peaks["bf16_sustained"] / 3.0))
            elif key == "adjoint_elementwise" and cnt:
                gbs = adj_bytes * n * B / (ent["ms_per_launch"] * 1e-3) / 1e9
                ent.update(bound="hbm", algorithmic_bytes_per_launch=adj_bytes * n * B, achieved_gbs=gbs, frac_of_hbm_peak=gbs / hbm)
            kernels[key] = ent
        top = max((k for k in kernels if k != "other"), key=lambda k: kernels[k]["ms_total"])
        contr_ms = sum(kernels[k]["ms_total"] for k in ("fwd_fused", "dgrad", "wgrad"))
        top_ent = kernels[top]
        roofline = {
            "bound": "tensor", "kernel": names[top], "achieved": top_ent.get("achieved_tflops", ach), "peak": peak_logical, "unit": "TFLOP/s",
            "frac": top_ent.get("frac_of_burst_peak", ach / peak_logical), "traffic": _ncu_traffic(),
            "note": ("the dominant kernel of the pass as it runs (largest share of the event-timed pass): achieved = logical 2*N*N*B flops per Euler step / "
                     "mean time per step of that stage incl. its element-wise epilogue and any launch gap (the persistent forward kernel covers T steps per launch); the kernel issues 3 kind::f16 MMAs "
                     "(binary16 hi/lo split words, fp32 accumulate) per logical product, so peak = bf16_tflops burst (%s)/3; `kernels` holds every stage, "
                     "`isolated_contractions` the bare contractions timed alone back to back" % peaks["source"]),
            "kernels": kernels,
            "whole_pass_frac": {"burst": f_bptt * value / 1e12 / peak_logical, "sustained": f_bptt * value / 1e12 / (peaks["bf16_sustained"] / 3.0),
                                "fwd_only_burst": f_fwd * fwd_value / 1e12 / peak_logical,
                                "flops_per_neuron_step": {"fwd": f_fwd, "bptt": f_bptt},
                                "note": "algorithmic flops per neuron-step (SURVEY 8d) x measured neuron-steps/s / (measured bf16 peak / 3)"},
            "isolated_contractions": {"launch_ms": {"fwd": ms_f, "adjoint": ms_b, "wgrad_chunk": ms_w, "wgrad_steps_per_chunk": chunk_steps},
                                      "achieved_tflops": {"fwd": ach, "adjoint": fl_b / (ms_b * 1e-3) / 1e12, "wgrad": fl_w / (ms_w * 1e-3) / 1e12},
                                      "frac_of_burst_peak": {"fwd": ach / peak_logical, "adjoint": fl_b / (ms_b * 1e-3) / 1e12 / peak_logical,
                                                             "wgrad": fl_w / (ms_w * 1e-3) / 1e12 / peak_logical}},
            "tf32_cublas_tflops": tf32_measured,
            "contraction_share_of_step": contr_ms / stage_total,
            "stage_pass_ms": stage_total, "timed_pass_ms": pass_ms,
        }
        if args.no_cpu_baseline:
            cpu_rate, cores, cpu_sec = float("nan"), os.cpu_count(), float("nan")
            cpub_rate, cpub_sec = float("nan"), float("nan")
        else:
            cpu_rate, cores, cpu_sec = cpu_reference_rate(n, CPU_T, 3)
            cpub_rate, _, cpub_sec = cpu_batched_rate(n, 64, 10, 3)
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_dev / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 (tcgen05 split-3 contractions on binary16 hi/lo words with exact power-of-two scales = 22-bit significand "
                      "products, fp32 accumulate)" if f16 else ("f32 (tcgen05 3xTF32 contractions, fp32 accumulate)" if use_tc else "f32")),
            "data": "synthetic",
            "config": {"workload": f"QIF recurrent spiking net BPTT (BASELINE configs[{2 if N_NEURONS == 4096 else 4}]): N={N_NEURONS}, batch={BATCH} trials/GPU, "
                                   "m=2, k=3, dt=1e-3, T=%d Euler steps fwd + adjoint per step, train W and W_out" % T,
                       "n": n, "batch_per_gpu": B, "t_inner": T, "n_in": N_IN, "n_out": N_OUT, "dt": DT,
                       "parallelism": f"trial-sharded x{world}" + (", NCCL grad all-reduce" if world > 1 else ""),
                       "initial_state": "v uniform in [-50, 99), s = 0 (active network: spikes, resets and surrogate terms occur within T)",
                       "l2": "working set (%.1f GB of checkpoints + %d MB of split weights per pass) exceeds the 126 MB L2"
                             % ((T + 1) * 2 * B * n * 4 / 1e9, (8 if f16 else 16) * n * n // 2**20)},
            "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": ms_e2e / args.steps,
                    "h2d_bytes_per_step": int(x_host.numel() * 4 + tgt_host.numel() * 4), "d2h_bytes_per_step": 4},
            "fwd": {"value": fwd_value, "unit": UNIT, "ms_per_step": ms_fwd / args.steps},
            "gpu_launches": int(launches),
            "clocks": clocks,
            "roofline": roofline,
            "cpu_baseline": {"value": cpu_rate, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": f"oracle port of the reference path, 1 trial x {CPU_T} steps BPTT, median of 3 ({cpu_sec:.2f} s each)"},
            "cpu_baseline_batched": {"value": cpub_rate, "unit": UNIT, "cores": cores, "kind": "restatement (not reference code)",
                                     "sample": f"same pass restated with a trial axis (sgemm on all cores): 64 trials x 10 steps BPTT, median of 3 ({cpub_sec:.2f} s each)"},
        }
        _OUT.emit(json.dumps(line))
    if world > 1:
        dist.barrier(group=cpu_group)         # a CPU wait: no GPU spins while rank 0 finishes its host-side work
        dist.destroy_process_group()


class _QuietStdout:
    """Route everything libraries print on fd 1 (e.g. NCCL's version banner) to stderr; the JSON line is the only stdout."""

    def __enter__(self):
        sys.stdout.flush()
        self.saved = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self.saved, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.saved, 1)
        os.close(self.saved)


_OUT = None


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--t-inner", type=int, default=T_INNER, help="Euler steps per BPTT pass (profiling runs use fewer)")
    ap.add_argument("--no-cpu-baseline", action="store_true", help="skip the CPU oracle timing (profiling runs)")
    ap.add_argument("--neurons", dest="n", type=int, default=N_NEURONS, help="neurons (8192: BASELINE configs[4], the multi-GPU sweep shape)")
    ap.add_argument("--trials", dest="batch", type=int, default=BATCH, help="trials per GPU")
    ap.add_argument("--precision", default="auto", choices=["auto", "3xf16", "3xtf32", "fp32"], help="contraction path (A/B runs)")
    args = ap.parse_args()
    globals()["T_INNER"] = args.t_inner
    globals()["PRECISION"] = args.precision
    if (args.n, args.batch) != (N_NEURONS, BATCH):
        globals()["N_NEURONS"], globals()["BATCH"] = args.n, args.batch
        globals()["METRIC"] = f"neuron-steps/sec (QIF N={args.n}, batch {args.batch}, BPTT fwd+bwd)"
    global _OUT
    with _QuietStdout() as q:
        _OUT = q
        if args.impl == "reference":
            run_reference(args)
        else:
            run_ours(args)


if __name__ == "__main__":
    main()
