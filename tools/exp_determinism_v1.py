import os, sys, hashlib
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import rectipy_b200 as rp
n, B, T, dt = 4096, 1024, 25, 1e-2
rng = np.random.default_rng(5)
W = (1.5 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
w_out = (rng.standard_normal((2, n)) / np.sqrt(n)).astype(np.float32)
SCALE = float(os.environ.get("XSCALE", "1"))
x = torch.tensor(rng.standard_normal((T, B, n)).astype(np.float32) * SCALE, device="cuda")
for prec in ("3xtf32", "3xf16", "fp32"):
    hs = []
    for rep in range(3):
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                            target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 0.5, "li_op/k": 1.3})
        net.add_func_node("out", 2, "identity"); net.add_edge("rnn", "out", weights=w_out)
        out = torch.stack(net.run(x, verbose=False)["out"])
        hs.append(hashlib.sha1(out.cpu().numpy().tobytes()).hexdigest()[:10] + " %.9e" % float(out.double().sum()))
    print(prec, hs, flush=True)
