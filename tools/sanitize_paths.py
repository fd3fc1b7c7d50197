"""Small forward + BPTT runs on every execution path, meant to be run under compute-sanitizer:

    compute-sanitizer --tool memcheck  python tools/sanitize_paths.py
    compute-sanitizer --tool racecheck python tools/sanitize_paths.py      # shared-memory hazards
    compute-sanitizer --tool synccheck python tools/sanitize_paths.py

Shapes are tiny (the tools slow kernels down 10-100x) but select the same kernels as the full-size runs: the persistent
few-trial kernels (cooperative launch, flag-in-data step exchange), the per-step FFMA kernels, the tcgen05 contractions with
both operand formats (fused forward epilogue, adjoint product, weight-gradient chunks, rolling-pipeline adjoint, operand
conversion), the mean-field templates and the persistent RLS kernel.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import numpy as np
import torch

import rectipy_b200 as rp
from rectipy_b200 import engine

TEMPLATES = {
    "qif": ("neuron_model_templates.spiking_neurons.qif.qif", "qif_op", "s", "s_in", {}),
    "qif_sfa": ("neuron_model_templates.spiking_neurons.qif.qif_sfa", "qif_sfa_op", "s", "s_in", {}),
    "li_tanh": ("neuron_model_templates.rate_neurons.leaky_integrator.tanh", "li_op", "tanh_op/r", "li_op/r_in", {}),
    "ik_biexp": ("neuron_model_templates.spiking_neurons.ik.ik_biexp", "ik_biexp_op", "s", "s_in",
                 dict(spike_threshold=40.0, spike_reset=-60.0)),
}


def run(model, n, B, T, prec, S=2):
    path, op, svar, tvar, skw = TEMPLATES[model]
    rng = np.random.default_rng(n + B)
    m, k = 2, 3
    net = rp.Network(1e-3 if model.startswith("qif") else 1e-2, device="cuda:0", batch=B, precision=prec)
    kw = dict(weights=rng.standard_normal((n, n)) * 2.0 / np.sqrt(n), source_var=svar, target_var=tvar, input_var=f"{op}/I_ext",
              train_params=["weights"])
    if model == "li_tanh":
        kw.update(output_var="li_op/v")
    else:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **skw)
    node = net.add_diffeq_node("rnn", path, **kw)
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=rng.standard_normal((n, m)), train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=rng.standard_normal((k, n)) / np.sqrt(n), train="gd")
    if model != "li_tanh":          # active state: spikes, resets and surrogate terms occur within T
        nsv = node.spec.n_sv
        y0 = np.zeros((B, nsv * n), dtype=np.float32)
        y0[:, :n] = rng.uniform(-50.0, 99.0, (B, n)) if model.startswith("qif") else rng.uniform(-60.0, 39.0, (B, n))
        node.reset(y0 if B > 1 else y0[0])
    x = rng.standard_normal((T, B, m)).astype(np.float32) * 5.0 + 10.0
    obs = net.run(x if B > 1 else x[:, 0, :], sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", f"{op}/v", True)])
    out = torch.stack(obs["out"])
    out.square().mean().backward()
    torch.cuda.synchronize()
    ok = all(torch.isfinite(p.grad).all() for p in net.parameters())
    print(f"{model:9s} n={n:4d} B={B:4d} T={T:3d} {prec:7s}: out {tuple(out.shape)} finite={bool(torch.isfinite(out).all())} grads finite={ok} "
          f"launches so far {engine.total_launches()}", flush=True)
    assert ok


def main():
    run("qif", 64, 2, 24, "fp32")            # persistent forward + persistent reverse sweep
    run("li_tanh", 96, 1, 24, "fp32")        # persistent, rate template
    run("qif_sfa", 50, 20, 12, "fp32")       # per-step FFMA kernels, ragged N
    run("ik_biexp", 40, 6, 12, "fp32")       # mean-field template: trial means, generic adjoint kernel, 4 state planes
    run("qif", 128, 128, 20, "3xf16")        # tcgen05, binary16 operands: fused forward, rolling-pipeline adjoint, conversion, 16-step chunk
    run("qif", 128, 128, 10, "3xtf32")       # tcgen05, tf32 operands: transposing adjoint kernel
    run("ik_biexp", 128, 128, 10, "3xf16")   # tensor-core path with the general-element epilogue
    X = torch.randn(40, 32, device="cuda"); Y = torch.randn(40, 3, device="cuda")
    W = torch.zeros(3, 32, device="cuda"); P = torch.eye(32, device="cuda")
    loss, _ = engine.rls_run(X, Y, W, P, 1.0)
    torch.cuda.synchronize()
    print("rls persistent: loss finite", bool(torch.isfinite(loss).all()), flush=True)


if __name__ == "__main__":
    main()
