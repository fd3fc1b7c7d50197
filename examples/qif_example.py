"""QIF-SFA network simulation -- the workload of the reference's documentation/qif_example.py (BASELINE config 1), written
against the current API (`reset_var="v"`; the upstream script still passes the removed `spike_def=` keyword).

    python examples/qif_example.py            # needs a B200 (sm_100a); the whole 40 000-step horizon is ONE kernel launch
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a source checkout
import time

import numpy as np

from rectipy_b200 import Network, random_connectivity

N, p = 1000, 0.1
np.random.seed(0)
W = random_connectivity(N, N, p, normalize=True)
etas = -5.0 + 1.0 * np.tan((np.pi / 2) * (2.0 * np.arange(1, N + 1) - N - 1) / (N + 1))
v_theta = 1e3

T, dt = 40.0, 1e-3
steps = int(T / dt)
inp = np.zeros((steps, 1), dtype=np.float32)
inp[int(10.0 / dt):int(30.0 / dt), 0] = 3.0

net = Network(dt, device="cuda:0")
net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif_sfa", weights=W, source_var="s", target_var="s_in",
                    input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_sfa_op",
                    spike_threshold=v_theta, spike_reset=-v_theta,
                    node_vars={"all/qif_sfa_op/eta": etas, "all/qif_sfa_op/alpha": 0.0, "all/qif_sfa_op/k": 15.0})
net.add_func_node("inp", 1, activation_function="tanh")
net.add_edge("inp", "qif")

t0 = time.perf_counter()
obs = net.run(inp, record_output=False, record_vars=[("qif", "s", True)], sampling_steps=100, verbose=False, enable_grad=False)
s_mean = obs.to_numpy(("qif", "s"))
print(f"{steps} steps of {N} neurons in {time.perf_counter() - t0:.3f} s; mean synaptic activity at the end of the stimulus: "
      f"{s_mean[299]:.4f}, after it: {s_mean[-1]:.4f}")
