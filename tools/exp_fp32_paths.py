"""A/B probe of the fp32 execution paths (persistent few-trial kernels, per-step launches): us per Euler step, best of 3.
Usage: python tools/exp_fp32_paths.py   (run from the root of the tree under test)"""
import os, sys, time
sys.path.insert(0, os.getcwd())
import numpy as np
import torch
import rectipy_b200 as rp

QIF = "neuron_model_templates.spiking_neurons.qif.qif"
TANH = "neuron_model_templates.rate_neurons.leaky_integrator.tanh"


def best(fn, reps=3):
    fn(); torch.cuda.synchronize()
    b = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize()
        b = min(b, time.perf_counter() - t0)
    return b


def qif_fwd(n, B, T, S=10):
    rng = np.random.default_rng(3)
    W = (2.0 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    net = rp.Network(1e-3, device="cuda:0", batch=B, precision="fp32")
    net.add_diffeq_node("qif", QIF, weights=W, source_var="s", target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike",
                        reset_var="v", op="qif_op", node_vars={"eta": rng.standard_normal(n) * 5 - 5})
    net.add_func_node("inp", 2, "identity"); net.add_edge("inp", "qif", weights=rng.standard_normal((n, 2)))
    net.add_func_node("out", 3, "identity"); net.add_edge("qif", "out", weights=rng.standard_normal((3, n)) / np.sqrt(n))
    x = torch.tensor(rng.standard_normal((T, B, 2)).astype(np.float32) * 5 + 8, device="cuda")
    y0 = net.state

    def run():
        net.reset(y0); net.run(x, sampling_steps=S, verbose=False, enable_grad=False)
    return best(run) / T * 1e6


def tanh_bptt(n, T):
    rng = np.random.default_rng(0)
    J = rng.standard_normal((n, n)) / np.sqrt(n)
    net = rp.Network(1e-2, device="cuda:0", precision="fp32")
    node = net.add_diffeq_node("tanh", TANH, weights=J, source_var="tanh_op/r", target_var="li_op/r_in", input_var="li_op/I_ext",
                               output_var="li_op/v", train_params=["weights"])
    x = torch.tensor(rng.standard_normal((T, n)).astype(np.float32), device="cuda")
    y0 = net.state

    def run():
        net.reset(y0); node["weights"].grad = None
        obs = net.run(x, verbose=False, enable_grad=True)
        torch.stack(obs["out"]).square().mean().backward()
    return best(run) / T * 1e6


print("tree", os.getcwd())
print("qif fwd N=1000 B=1  T=20000: %.2f us/step" % qif_fwd(1000, 1, 20000), flush=True)
print("qif fwd N=1000 B=32 T=10000: %.2f us/step" % qif_fwd(1000, 32, 10000), flush=True)
print("tanh bptt N=200  B=1 T=5000: %.2f us/step" % tanh_bptt(200, 5000), flush=True)
print("tanh bptt N=4096 B=1 T=300 : %.2f us/step" % tanh_bptt(4096, 300), flush=True)
