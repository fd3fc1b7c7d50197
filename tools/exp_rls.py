import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np, torch
import rectipy_b200 as rp
from rectipy_b200 import engine
n, T = 600, 5000
rng = np.random.default_rng(1)
W = rng.standard_normal((n, n)) / np.sqrt(n); x = rng.standard_normal((T, n)).astype(np.float32); targets = rng.standard_normal((T, 3)).astype(np.float32)
def mk():
    net = rp.Network(1e-2, device="cuda:0")
    net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                        target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 1.0})
    net.add_func_node("out", 3, "identity"); net.add_edge("rnn", "out", train="rls")
    return net
def t(fn, reps=3):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return min(ts) * 1e3
print("build net ms", t(mk))
net = mk()
print("fit_rls ms", t(lambda: net.fit_rls(x, targets, update_steps=1, sampling_steps=100, verbose=False)))
net2 = rp.Network(1e-2, device="cuda:0")
net2.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                     target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 1.0})
print("run only ms", t(lambda: net2.run(x, sampling_steps=1, verbose=False, enable_grad=False)))
X = torch.randn(T, n, device="cuda"); Y = torch.randn(T, 3, device="cuda")
Wr = torch.zeros(3, n, device="cuda"); P = torch.eye(n, device="cuda").contiguous()
print("rls_run only ms", t(lambda: engine.rls_run(X, Y, Wr, P, 1.0, 1)))
os.environ["RP_NO_PERSISTENT"] = "1"
print("rls_run per-step ms", t(lambda: engine.rls_run(X, Y, Wr, P, 1.0, 1)))
