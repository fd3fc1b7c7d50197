"""Variance probe for fit_rls (round 1f): per-repetition wall times of the reservoir run and of the RLS kernel, plus SM clock."""
import os, sys, time, subprocess
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rectipy_b200 as rp
from rectipy_b200 import engine

n, T = 600, 5000
rng = np.random.default_rng(1)
W = rng.standard_normal((n, n)) / np.sqrt(n); x = rng.standard_normal((T, n)).astype(np.float32); targets = rng.standard_normal((T, 3)).astype(np.float32)
net = rp.Network(1e-2, device="cuda:0")
net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                    target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 1.0})
net.add_func_node("out", 3, "identity"); edge = net.add_edge("rnn", "out", train="rls")


def clock():
    try:
        return subprocess.run(["nvidia-smi", "--query-gpu=clocks.sm,power.draw,pstate", "--format=csv,noheader"], capture_output=True, text=True).stdout.strip()
    except OSError:
        return "?"


def wall(fn):
    torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize(); return (time.perf_counter() - t0) * 1e3, r


for rep in range(8):
    t_fit, _ = wall(lambda: net.fit_rls(x, targets, update_steps=1, sampling_steps=100, verbose=False))
    t_run, obs = wall(lambda: net.run(x, sampling_steps=1, verbose=False, enable_grad=False))
    X = torch.randn(T, n, device="cuda"); Y = torch.randn(T, 3, device="cuda")
    t_rls, _ = wall(lambda: engine.rls_run(X, Y, edge.weights, edge.P, edge.beta, 1))
    print(f"rep {rep}: fit_rls {t_fit:8.1f} ms   run {t_run:7.1f} ms   rls_run {t_rls:7.1f} ms   |P| {float(edge.P.abs().max()):.3e} |W| {float(edge.weights.abs().max()):.3e}"
          f" finite {bool(torch.isfinite(edge.P).all())}   [{clock()}]", flush=True)

# which part of fit_rls is slow?  (synchronising between the reservoir run and the RLS kernel)
_orig = engine.rls_run
_t = []


def _timed(*a, **k):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); t0 = time.perf_counter(); e0.record(); r = _orig(*a, **k); e1.record(); t_host = (time.perf_counter() - t0) * 1e3
    torch.cuda.synchronize(); _t.append(((time.perf_counter() - t0) * 1e3, e0.elapsed_time(e1), t_host))
    return r


engine.rls_run = _timed
for rep in range(16):
    t_fit, _ = wall(lambda: net.fit_rls(x, targets, update_steps=1, sampling_steps=100, verbose=False))
    print(f"split rep {rep}: fit_rls {t_fit:8.1f} ms   of which rls_run wall {_t[-1][0]:7.1f} ms, on the device (CUDA events) {_t[-1][1]:7.1f} ms, host time of the call {_t[-1][2]:6.2f} ms", flush=True)
