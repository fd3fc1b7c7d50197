"""Time every BASELINE.json config on the engine (GPU) next to the CPU oracle port (bounded sample). Prints a markdown table."""
import os, sys, time, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import numpy as np, torch
import rectipy_b200 as rp
import rectipy_oracle as orc

torch.set_num_threads(os.cpu_count())
rows = []

def gpu_time(fn, reps=3):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2]

def cpu_time(fn, reps=1):
    fn(); ts = []
    for _ in range(reps):
        t0 = time.perf_counter(); fn(); ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2]

# ---- C1: QIF-SFA forward, N=1000, dt=1e-3, T=40000, S=100, neuron-mean of s (documentation/qif_example.py) -------
n, T, dt = 1000, 40000, 1e-3
np.random.seed(0)
W = rp.random_connectivity(n, n, 0.1, normalize=True); etas = orc.lorentzian_etas(n)
inp = np.zeros((T, 1), dtype=np.float32); inp[10000:30000, 0] = 3.0
net = rp.Network(dt, device="cuda:0")
net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif_sfa", weights=W, source_var="s", target_var="s_in",
                    input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_sfa_op",
                    node_vars={"eta": etas, "alpha": 0.0, "k": 15.0}, spike_threshold=1e3, spike_reset=-1e3)
net.add_func_node("inp", 1, "tanh"); net.add_edge("inp", "qif", weights=np.ones((n, 1)))
y0 = net.state
def c1():
    net.reset(y0); return net.run(inp, record_output=False, record_vars=[("qif", "s", True)], sampling_steps=100, verbose=False, enable_grad=False)
tg = gpu_time(c1)
Tc = 2000
def c1_cpu():
    node = orc.make_node("qif_sfa", n, W, dt, params=dict(eta=etas, alpha=0.0, k=15.0), dtype=torch.float32, spike_threshold=1e3, spike_reset=-1e3)
    onet = orc.OracleNet(node, w_in=torch.ones(n, 1), in_act="tanh")
    onet.run(torch.tensor(inp[:Tc]), sampling_steps=100, record_vars=[("s", True)], enable_grad=False)
tc = cpu_time(c1_cpu)
rows.append(("C1 QIF-SFA fwd N=1000 B=1 T=40000", n * T / tg, n * Tc / tc, tg / T * 1e6))

# ---- sweep: QIF forward N=1000, 32 trials (north_star: parameter sweeps of 8-64 trials), persistent 2-D grid ---------------------
n, T, dt, Bs = 1000, 20000, 1e-3, 32
rng = np.random.default_rng(3)
Ws = (2.0 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
nets = rp.Network(dt, device="cuda:0", batch=Bs)
nets.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=Ws, source_var="s", target_var="s_in", input_var="I_ext",
                     output_var="s", spike_var="spike", reset_var="v", op="qif_op", node_vars={"eta": orc.lorentzian_etas(n)})
nets.add_func_node("inp", 2, "identity"); nets.add_edge("inp", "qif", weights=rng.standard_normal((n, 2)))
nets.add_func_node("out", 3, "identity"); nets.add_edge("qif", "out", weights=rng.standard_normal((3, n)) / np.sqrt(n))
xs = torch.randn(T, Bs, 2, device="cuda") * 5 + 8; y0s = nets.state
def sweep():
    nets.reset(y0s); return nets.run(xs, sampling_steps=10, verbose=False, enable_grad=False)
tg = gpu_time(sweep)
rows.append((f"sweep QIF fwd N={n} B={Bs} T={T} (readout, S=10)", n * Bs * T / tg, float("nan"), tg / T * 1e6))

# ---- sweep, 64 trials: neither the tensor-core shapes (multiples of 128) nor the persistent kernels hold it -> precision="auto" pads to 1024 x 128
Bs2 = 64
nets2 = rp.Network(dt, device="cuda:0", batch=Bs2)
nets2.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=Ws, source_var="s", target_var="s_in", input_var="I_ext",
                      output_var="s", spike_var="spike", reset_var="v", op="qif_op", node_vars={"eta": orc.lorentzian_etas(n)})
nets2.add_func_node("inp", 2, "identity"); nets2.add_edge("inp", "qif", weights=rng.standard_normal((n, 2)))
nets2.add_func_node("out", 3, "identity"); nets2.add_edge("qif", "out", weights=rng.standard_normal((3, n)) / np.sqrt(n))
xs2 = torch.randn(T, Bs2, 2, device="cuda") * 5 + 8; y0s2 = nets2.state
def sweep2():
    nets2.reset(y0s2); return nets2.run(xs2, sampling_steps=10, verbose=False, enable_grad=False)
tg = gpu_time(sweep2)
rows.append((f"sweep QIF fwd N={n} B={Bs2} T={T} (readout, S=10; padded to 1024 x 128, tcgen05)", n * Bs2 * T / tg, float("nan"), tg / T * 1e6))
del nets2, xs2

# ---- C2: LI-tanh rate net BPTT, N=200, dt=1e-2, T=10000 (documentation/bptt_rate_neurons.py) -----------------------
for n in (200, 4096):
    T = 10000 if n == 200 else 500
    rng = np.random.default_rng(0)
    J = rng.standard_normal((n, n)); J /= np.max(np.abs(np.linalg.eigvals(J)))
    tau = rng.uniform(10, 20, n)
    t = np.linspace(0, T * 1e-2, T); x = (np.sin(2 * np.pi * 0.2 * t) * 10.0).astype(np.float32)[:, None]
    target = rng.standard_normal((T, n)).astype(np.float32)
    net = rp.Network(1e-2, device="cuda:0")
    node = net.add_diffeq_node("tanh", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=J, source_var="tanh_op/r",
                               target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", train_params=["weights"],
                               node_vars={"all/li_op/eta": 2.0, "all/li_op/tau": tau, "all/li_op/k": 2.0})
    y0 = net.state; tgt = torch.tensor(target, device="cuda")
    def c2():
        net.reset(y0); node["weights"].grad = None
        obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=True)
        torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt).backward()
    tg = gpu_time(c2)
    Tc = 500 if n == 200 else 40
    def c2_cpu():
        onode = orc.make_node("li_tanh", n, J, 1e-2, params=dict(eta=2.0, tau=tau, k=2.0), dtype=torch.float32, train_params=["weights"])
        onet = orc.OracleNet(onode)
        r = onet.run(torch.tensor(np.repeat(x[:Tc], n, axis=1)), enable_grad=True)
        torch.nn.functional.mse_loss(torch.stack(r["out"]), torch.tensor(target[:Tc])).backward()
    tc = cpu_time(c2_cpu)
    rows.append((f"C2 LI-tanh BPTT N={n} B=1 T={T}", n * T / tg, n * Tc / tc, tg / T * 1e6))

# ---- C4: reservoir + ridge (N=100, m=5, T=300000, S=1) and RLS (N=600) ----------------------------------------------
n, m, T = 100, 5, 300000
rng = np.random.default_rng(1)
W = rng.standard_normal((n, n)) / np.sqrt(n); w_in = rng.standard_normal((n, m))
x = rng.standard_normal((T, m)).astype(np.float32); targets = rng.standard_normal((T, 2)).astype(np.float32)
def c4():
    net = rp.Network(1e-2, device="cuda:0")
    net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                        target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 1.0})
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in)
    net.fit_ridge(x, targets, sampling_steps=1, alpha=1e-4, verbose=False, add_readout_node=False)
tg = gpu_time(c4, reps=2)
Tc = 3000
def c4_cpu():
    onode = orc.make_node("li_tanh", n, W, 1e-2, params=dict(tau=1.0), dtype=torch.float32)
    onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, dtype=torch.float32))
    r = onet.run(torch.tensor(x[:Tc]), enable_grad=False)
    orc.ridge_fit(torch.stack(r["out"]), torch.tensor(targets[:Tc]), 1e-4)
tc = cpu_time(c4_cpu)
rows.append((f"C4 reservoir+ridge N=100 m=5 T={T}", n * T / tg, n * Tc / tc, tg / T * 1e6))

n, T = 600, 5000
W = rng.standard_normal((n, n)) / np.sqrt(n); x = rng.standard_normal((T, n)).astype(np.float32); targets = rng.standard_normal((T, 3)).astype(np.float32)
def c4r():
    net = rp.Network(1e-2, device="cuda:0")
    net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                        target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 1.0})
    net.add_func_node("out", 3, "identity"); net.add_edge("rnn", "out", train="rls")
    net.fit_rls(x, targets, update_steps=1, sampling_steps=100, verbose=False)
tg = gpu_time(c4r, reps=2)
Tc = 300
def c4r_cpu():
    onode = orc.make_node("li_tanh", n, W, 1e-2, params=dict(tau=1.0), dtype=torch.float32)
    rls = orc.OracleRLS(n, 3, dtype=torch.float32)
    xt, yt = torch.tensor(x[:Tc]), torch.tensor(targets[:Tc])
    with torch.no_grad():
        for s in range(Tc):
            o = onode.forward(xt[s]); rls.update(o, yt[s], rls.forward(o))
tc = cpu_time(c4r_cpu)
rows.append((f"C4 reservoir+RLS N=600 T={T}", n * T / tg, n * Tc / tc, tg / T * 1e6))

# ---- C5 (one rank's share): QIF N=8192, 1024 trials, BPTT T=50 ------------------------------------------------------
n, B, T = 8192, 1024, 50
rng = np.random.default_rng(2)
net = rp.Network(1e-3, device="cuda:0", batch=B)
node = net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=(2 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32),
                           source_var="s", target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op",
                           node_vars={"eta": orc.lorentzian_etas(n)}, train_params=["weights"])
net.add_func_node("inp", 2, "identity"); net.add_edge("inp", "qif", weights=rng.standard_normal((n, 2)))
net.add_func_node("out", 3, "identity"); net.add_edge("qif", "out", weights=rng.standard_normal((3, n)) / np.sqrt(n), train="gd")
x = torch.randn(T, B, 2, device="cuda") * 5 + 8; tgt = torch.randn(T, B, 3, device="cuda"); y0 = net.state
def c5():
    net.reset(y0); node["weights"].grad = None
    obs = net.run(x, verbose=False, enable_grad=True)
    torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt).backward()
tg = gpu_time(c5, reps=2)
rows.append((f"C5 QIF BPTT N=8192 B=1024/GPU T={T} (1 GPU share)", n * B * T / tg, float("nan"), tg / T * 1e6))

print("| config | engine neuron-steps/s | CPU oracle neuron-steps/s (%d cores) | ratio | engine us/step |" % os.cpu_count())
print("|---|---:|---:|---:|---:|")
for name, g, c, us in rows:
    print(f"| {name} | {g:.3e} | {c:.3e} | {g / c:.0f}x | {us:.1f} |")
