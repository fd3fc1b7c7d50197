"""Repetition-to-repetition spread of the execution paths (wall clock with a device synchronise on both sides, ms): min / median / max over
`reps` calls.  Written after an intermittent stall was found in rp_rls_run (round 1f); prints one line per workload."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
import rectipy_b200 as rp


def spread(name, fn, reps=30):
    fn(); torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); fn(); torch.cuda.synchronize(); ts.append((time.perf_counter() - t0) * 1e3)
    ts.sort()
    print(f"| {name} | {reps} | {ts[0]:.2f} | {ts[len(ts) // 2]:.2f} | {ts[-1]:.2f} | {ts[-1] / ts[len(ts) // 2]:.2f} |", flush=True)


print("| workload | calls | min ms | median ms | max ms | max / median |\n|---|---:|---:|---:|---:|---:|")
rng = np.random.default_rng(0)
# persistent forward (C1 style)
n, T = 1000, 4000
net = rp.Network(1e-3, device="cuda:0")
net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif_sfa", weights=rp.random_connectivity(n, n, 0.1, normalize=True), source_var="s",
                    target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_sfa_op",
                    node_vars={"eta": -5.0 + np.tan(np.pi / 2 * (2.0 * np.arange(1, n + 1) - n - 1) / (n + 1)), "alpha": 0.0, "k": 15.0},
                    spike_threshold=1e3, spike_reset=-1e3)
net.add_func_node("inp", 1, "tanh"); net.add_edge("inp", "qif", weights=np.ones((n, 1)))
inp = torch.full((T, 1), 3.0, device="cuda"); y0 = net.state


def c1():
    net.reset(y0); net.run(inp, record_output=False, record_vars=[("qif", "s", True)], sampling_steps=100, verbose=False, enable_grad=False)


spread("persistent forward, QIF-SFA N=1000 B=1 T=4000", c1)
# persistent forward + reverse sweep (C2 style)
n2, T2 = 200, 2000
J = rng.standard_normal((n2, n2)); J /= np.max(np.abs(np.linalg.eigvals(J)))
net2 = rp.Network(1e-2, device="cuda:0")
net2.add_diffeq_node("tanh", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=J, source_var="tanh_op/r", target_var="li_op/r_in",
                     input_var="li_op/I_ext", output_var="li_op/v", train_params=["weights"], node_vars={"all/li_op/eta": 2.0, "all/li_op/k": 2.0})
x2 = torch.randn(T2, 1, device="cuda"); tgt2 = torch.randn(T2, n2, device="cuda"); y02 = net2.state


def c2():
    net2.reset(y02)
    for p in net2.parameters():
        p.grad = None
    obs = net2.run(x2, sampling_steps=1, verbose=False, enable_grad=True)
    torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt2).backward()


spread("persistent forward + reverse, LI-tanh N=200 B=1 T=2000 BPTT", c2)
# per-step FFMA path
n3, B3, T3 = 96, 20, 200
net3 = rp.Network(1e-3, device="cuda:0", batch=B3)
net3.add_diffeq_node("q", "neuron_model_templates.spiking_neurons.qif.qif", weights=rng.standard_normal((n3, n3)) * 2 / np.sqrt(n3), source_var="s",
                     target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op", train_params=["weights"])
x3 = torch.randn(T3, B3, n3, device="cuda") + 10.0; y03 = net3.state


def c3():
    net3.reset(y03)
    for p in net3.parameters():
        p.grad = None
    obs = net3.run(x3, sampling_steps=1, verbose=False, enable_grad=True)
    torch.stack(obs["out"]).square().mean().backward()


spread("per-step FFMA, QIF N=96 B=20 T=200 BPTT", c3)
# tensor-core path at the headline shape
n4, B4, T4 = 4096, 1024, 50
net4 = rp.Network(1e-3, device="cuda:0", batch=B4)
node4 = net4.add_diffeq_node("q", "neuron_model_templates.spiking_neurons.qif.qif", weights=rng.standard_normal((n4, n4)) * 2 / np.sqrt(n4), source_var="s",
                             target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op", train_params=["weights"])
net4.add_func_node("inp", 2, "identity"); net4.add_edge("inp", "q", weights=rng.standard_normal((n4, 2)))
net4.add_func_node("out", 3, "identity"); net4.add_edge("q", "out", weights=rng.standard_normal((3, n4)) / 64, train="gd")
node4.reset(np.concatenate([rng.uniform(-50, 99, (B4, n4)), np.zeros((B4, n4))], axis=1).astype(np.float32))
x4 = torch.randn(T4, B4, 2, device="cuda") * 5 + 10; tgt4 = torch.randn(T4, B4, 3, device="cuda"); y04 = net4.state


def c4():
    net4.reset(y04)
    for p in net4.parameters():
        p.grad = None
    obs = net4.run(x4, sampling_steps=1, verbose=False, enable_grad=True)
    torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt4).backward()


spread("tcgen05 path, QIF N=4096 B=1024 T=50 BPTT", c4, reps=20)
