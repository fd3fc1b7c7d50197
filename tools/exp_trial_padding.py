"""Experiment: is the tensor-core path with the trial axis padded to 128 faster than the fp32 paths at the true trial count?
Prints ms per 200-step BPTT pass (QIF, m=2, k=3).  Run on a B200: python tools/exp_trial_padding.py"""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rectipy_b200 as rp

QIF = "neuron_model_templates.spiking_neurons.qif.qif"


def one(n, B, precision, T=200, reps=3, grad=True):
    rng = np.random.default_rng(0)
    W = rng.standard_normal((n, n)).astype(np.float32) * 2.0 / np.sqrt(n)
    net = rp.Network(1e-3, device="cuda:0", batch=B, precision=precision)
    node = net.add_diffeq_node("qif", QIF, weights=W, source_var="s", target_var="s_in", input_var="I_ext", output_var="s",
                               spike_var="spike", reset_var="v", op="qif_op", node_vars={"eta": rng.standard_normal(n) * 5, "k": 1.5},
                               train_params=["weights"] if grad else None)
    net.add_func_node("inp", 2, "identity"); net.add_edge("inp", "qif", weights=rng.standard_normal((n, 2)))
    net.add_func_node("out", 3, "identity"); net.add_edge("qif", "out", weights=rng.standard_normal((3, n)) / np.sqrt(n), train="gd" if grad else None)
    x = torch.tensor(rng.standard_normal((T, B, 2)).astype(np.float32) * 5 + 10, device="cuda:0")
    y0 = net.state
    best = 1e9
    for r in range(reps + 1):
        net.reset(y0)
        torch.cuda.synchronize(); t0 = time.perf_counter()
        obs = net.run(x, verbose=False, enable_grad=grad)
        if grad:
            node["weights"].grad = None
            torch.stack(obs["out"]).square().mean().backward()
        torch.cuda.synchronize()
        if r:
            best = min(best, (time.perf_counter() - t0) * 1e3)
    return best


import sys as _sys
small = len(_sys.argv) > 1 and _sys.argv[1] == "small"
for grad in (False, True):
    for n in ((1536, 2048, 4096) if small else (1024, 2048, 4096)):
        for B in ((1, 2, 4, 8, 16) if small else (32, 64, 96)):
            os.environ["RECTIPY_B200_NO_PADDING"] = "1"
            a = one(n, B, "fp32", grad=grad)
            b = one(n, 128, "auto", grad=grad)
            print(f"grad={int(grad)} n={n} B={B}: fp32 path {a:8.2f} ms   padded-to-128 tensor-core path {b:8.2f} ms   x{a / b:.2f}", flush=True)
