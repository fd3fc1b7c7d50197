"""Trial-parallel BPTT of a recurrent spiking network with the surrogate spike gradient -- the scaled-up version of the
reference's documentation/bptt_spiking_neurons_recurrent.py (BASELINE config 3: QIF, N = 4096, 1024 trials).

    python examples/bptt_spiking_recurrent.py [N] [trials] [steps]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a source checkout
import sys
import time

import numpy as np
import torch

from rectipy_b200 import Network

N = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
B = int(sys.argv[2]) if len(sys.argv) > 2 else 1024
T = int(sys.argv[3]) if len(sys.argv) > 3 else 200
dt, n_in, n_out = 1e-3, 2, 3
rng = np.random.default_rng(0)
etas = -5.0 + np.tan((np.pi / 2) * (2.0 * np.arange(1, N + 1) - N - 1) / (N + 1))


def build(J, W_out, train):
    net = Network(dt, device="cuda:0", batch=B)
    net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=J, source_var="s", target_var="s_in",
                        input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op",
                        node_vars={"eta": etas}, train_params=["weights"] if train else None)
    net.add_func_node("inp", n_in, "identity")
    net.add_edge("inp", "qif", weights=W_in)
    net.add_func_node("out", n_out, "identity")
    net.add_edge("qif", "out", weights=W_out, train="gd" if train else None)
    return net


W_in = rng.standard_normal((N, n_in)).astype(np.float32)
target_net = build(2.0 * rng.standard_normal((N, N)).astype(np.float32) / np.sqrt(N), rng.standard_normal((n_out, N)) / np.sqrt(N), False)
learner = build(2.0 * rng.standard_normal((N, N)).astype(np.float32) / np.sqrt(N), rng.standard_normal((n_out, N)) / np.sqrt(N), True)
y0 = np.concatenate([rng.uniform(-50, 99, (B, N)), np.zeros((B, N))], axis=1).astype(np.float32)   # network already active

t = np.arange(T, dtype=np.float32) * dt
x = (rng.uniform(5, 15, (1, B, 1)) * np.sin(2 * np.pi * np.array([3.0, 5.0]) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, n_in))) + 8.0)
x = torch.tensor(x, dtype=torch.float32, device="cuda:0")

target_net.get_node("qif").reset(y0)
targets = torch.stack(target_net.run(x, verbose=False, enable_grad=False)["out"])

opt = torch.optim.Adam(learner.parameters(), lr=1e-3)
for epoch in range(5):
    learner.get_node("qif").reset(y0)
    t0 = time.perf_counter()
    obs = learner.run(x, verbose=False, enable_grad=True)
    loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]), targets)
    opt.zero_grad()
    loss.backward()
    opt.step()
    torch.cuda.synchronize()
    el = time.perf_counter() - t0
    print(f"epoch {epoch}: loss {loss.item():.5f}   {N * B * T / el:.3e} neuron-steps/s (fwd + bwd + optimizer)")
