"""A user-defined neuron model: the template is plain PyRates YAML, its equations match none of the engine's compiled vector fields,
so `add_diffeq_node` generates the CUDA kernels (Euler step + reverse-time adjoint) from the equations and compiles them with NVRTC.

    python examples/custom_template.py          (needs a B200; writes the template next to this script)

Two populations of quadratic integrate-and-fire neurons in ONE node, each with its own spike variable: lists for `spike_var` / `reset_var`
select the reference's MultiSpikeResetNet semantics (rectipy/nodes.py:404-465)."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import rectipy_b200 as rp  # noqa: E402

YAML = """
ei_op:
  base: OperatorTemplate
  equations:
    - "v_e' = (v_e^2 + eta_e + I_ext)/tau_e + J_ee*s_in - J_ei*s_i"
    - "s_e' = -s_e/tau_s + spike_e"
    - "v_i' = (v_i^2 + eta_i)/tau_i + J_ie*s_e"
    - "s_i' = -s_i/tau_s + spike_i"
  variables:
    s_e: output(0.0)
    v_e: variable(-2.0)
    v_i: variable(-2.0)
    s_i: variable(0.0)
    eta_e: -5.0
    tau_e: 1.0
    J_ee: 1.0
    J_ei: 2.0
    eta_i: -5.0
    tau_i: 0.5
    J_ie: 3.0
    tau_s: 0.8
    I_ext: input(0.0)
    spike_e: input(0.0)
    spike_i: input(0.0)
    s_in: input(0.0)
ei:
  base: NodeTemplate
  operators:
    - ei_op
"""

here = os.path.dirname(os.path.abspath(__file__))
os.makedirs(os.path.join(here, "my_templates"), exist_ok=True)
with open(os.path.join(here, "my_templates", "twopop.yaml"), "w") as fh:
    fh.write(YAML)
os.chdir(here)                                    # template paths resolve against the working directory, like PyRates' do

n, trials, T, dt = 200, 16, 2000, 1e-3
rng = np.random.default_rng(0)
W = rng.standard_normal((n, n)) * 2.0 / np.sqrt(n)
net = rp.Network(dt, device="cuda:0", batch=trials)
node = net.add_diffeq_node("ei", "my_templates.twopop.ei", weights=W, source_var="s_e", target_var="s_in", input_var="I_ext",
                           output_var="s_e", spike_var=["spike_e", "spike_i"], reset_var=["v_e", "v_i"], op="ei_op",
                           node_vars={"eta_e": rng.standard_cauchy(n).clip(-20, 20) + 8.0, "eta_i": rng.uniform(-2, 6, n)},
                           train_params=["weights", "J_ei"], spike_threshold=100.0, spike_reset=-100.0)
net.add_func_node("inp", 2, "identity"); net.add_edge("inp", "ei", weights=rng.standard_normal((n, 2)))
net.add_func_node("out", 1, "identity"); net.add_edge("ei", "out", weights=rng.standard_normal((1, n)) / np.sqrt(n), train="gd")
print(type(node).__name__, "on generated kernels:", node.spec.name, "| state planes", node.spec.planes)

t = np.arange(T) * dt
x = 10.0 * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, trials, 2)) * t[:, None, None]) + 12.0
obs = net.run(x, sampling_steps=5, verbose=False, enable_grad=True, record_vars=[("ei", "v_i", False)])
out = torch.stack(obs["out"])
loss = out.square().mean()
loss.backward()
print("records", tuple(out.shape), "| loss %.4e" % float(loss.detach()), "| |dW| %.3e" % float(node["weights"].grad.norm()),
      "| dJ_ei %.3e" % float(node["J_ei"].grad))
print("resets of the inhibitory population seen at record steps:", int((obs.to_numpy(("ei", "v_i")) == -100.0).sum()))
