"""Two LIF populations coupled by a feed-forward and a feedback edge -- the workload of the reference's
documentation/rnn_tryout.py, written against the current API (`reset_var="v"`; the upstream script still passes the removed
`spike_def=` keyword).

    python examples/feedback_network.py       # needs a B200 (sm_100a)

`FeedbackNetwork` advances the populations in lockstep (one engine call per population and step), so this path is host-stepped:
it covers the reference's API, the fast paths are the single-plan `Network` runs of examples/qif_example.py and
examples/bptt_spiking_recurrent.py.
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))      # run from a source checkout
import time

import numpy as np

from rectipy_b200 import FeedbackNetwork

dt = 1e-2
N, k = 100, 10.0
rng = np.random.default_rng(0)
neuron = "neuron_model_templates.spiking_neurons.lif.lif"

net = FeedbackNetwork(dt, device="cuda:0")
for name in ("p1", "p2"):
    net.add_diffeq_node(name, node=neuron, input_var="I_ext", output_var="s", weights=rng.standard_normal((N, N)), source_var="s",
                        target_var="s_in", op="lif_op", spike_var="spike", reset_var="v")
net.add_edge("p1", "p2", weights=k * rng.random((N, N)), train=None)                       # feed-forward: p1 excites p2
net.add_edge("p2", "p1", weights=-10.0 * k * rng.random((N, N)), feedback=True)            # feedback: p2 inhibits p1

steps = 2000
inp = np.zeros((steps, 1), dtype=np.float32) + 100.0
t0 = time.perf_counter()
obs = net.run(inputs=inp, sampling_steps=10, enable_grad=False, verbose=False)
out = obs.to_numpy("out")
print(f"{steps} lockstep steps of 2 x {N} LIF neurons in {time.perf_counter() - t0:.2f} s; "
      f"mean output of p2 over the last 50 samples: {out[-50:].mean():.4f}")
