"""Is the batched forward bit-reproducible inside one process?  (python tools/exp_determinism.py)  Rate (LI-tanh, per-step launches) and
spiking (QIF: persistent kernel, or per-step launches with RP_NO_FWD_PERSIST=1), N=4096, 1024 trials, tensor-core paths vs fp32."""
import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rectipy_b200 as rp

n, B, T = 4096, 1024, 25
rng = np.random.default_rng(5)
W = (1.5 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
w_out = (rng.standard_normal((2, n)) / np.sqrt(n)).astype(np.float32)
w_in = rng.standard_normal((n, 2)).astype(np.float32)
x = torch.tensor(rng.standard_normal((T, B, n)).astype(np.float32), device="cuda")
x2 = torch.tensor((rng.standard_normal((T, B, 2)) * 5 + 10).astype(np.float32), device="cuda")
y0q = np.concatenate([rng.uniform(-50, 99, (B, n)), np.zeros((B, n))], axis=1).astype(np.float32)
for model in ("li_tanh", "qif"):
    for prec in ("3xtf32", "3xf16") + (("fp32",) if model == "li_tanh" else ()):
        hs = []
        for rep in range(3):
            if model == "li_tanh":
                net = rp.Network(1e-2, device="cuda:0", batch=B, precision=prec)
                net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                                    target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 0.5, "li_op/k": 1.3})
                net.add_func_node("out", 2, "identity"); net.add_edge("rnn", "out", weights=w_out)
                out = torch.stack(net.run(x, verbose=False)["out"])
            else:
                net = rp.Network(1e-3, device="cuda:0", batch=B, precision=prec)
                node = net.add_diffeq_node("rnn", "neuron_model_templates.spiking_neurons.qif.qif", weights=W * 2, source_var="s", target_var="s_in",
                                           input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op", node_vars={"eta": rng.standard_normal(n) * 0 - 3.0})
                net.add_func_node("inp", 2, "identity"); net.add_edge("inp", "rnn", weights=w_in)
                net.add_func_node("out", 2, "identity"); net.add_edge("rnn", "out", weights=w_out)
                node.reset(y0q)
                out = torch.stack(net.run(x2, verbose=False)["out"])
            hs.append(hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:10] + " %.9e" % float(out.double().sum()))
        print(model, prec, "NO_FWD_PERSIST=" + os.environ.get("RP_NO_FWD_PERSIST", "0"), hs, flush=True)
