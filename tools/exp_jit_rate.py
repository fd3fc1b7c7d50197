"""What the run-time compiled (per-step fp32) path costs at the reference's own script shape: a generated 2-variable rate field and the
generated two-population spiking field, N = 1000, one trial, forward.  python tools/exp_jit_rate.py  (B200)"""
import os, sys, time, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np
import torch
import rectipy_b200 as rp
from test_gpu_jit import YAML, EI_YAML

d = tempfile.mkdtemp(); os.makedirs(os.path.join(d, "mymodels"))
open(os.path.join(d, "mymodels", "custom.yaml"), "w").write(YAML)
open(os.path.join(d, "mymodels", "twopop.yaml"), "w").write(EI_YAML)
os.chdir(d)
n, T = 1000, 20000
rng = np.random.default_rng(0)
W = (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
for name in ("fhn", "ei"):
    net = rp.Network(1e-3 if name == "ei" else 1e-2, device="cuda:0")
    t0 = time.perf_counter()
    if name == "fhn":
        net.add_diffeq_node("rnn", "mymodels.custom.fhn", weights=W, source_var="r", target_var="r_in", input_var="I_ext", output_var="v")
    else:
        net.add_diffeq_node("rnn", "mymodels.twopop.ei", weights=W * 2, source_var="s_e", target_var="s_in", input_var="I_ext", output_var="s_e",
                            spike_var=["spike_e", "spike_i"], reset_var=["v_e", "v_i"], node_vars={"ei_op/eta_e": rng.standard_normal(n) * 5 + 5})
    t_gen = time.perf_counter() - t0
    x = torch.tensor(rng.standard_normal((T, n)).astype(np.float32), device="cuda")
    y0 = net.state
    best = 1e9
    for _ in range(3):
        net.reset(y0); torch.cuda.synchronize(); t0 = time.perf_counter()
        net.run(x, sampling_steps=100, verbose=False); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print(f"{name}: code generation + NVRTC {t_gen:.1f} s; forward N={n} B=1 T={T}: {best / T * 1e6:.1f} us/step = {n * T / best:.3e} neuron-steps/s", flush=True)
