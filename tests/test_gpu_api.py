"""GPU tests of the drop-in API semantics (mirrors rectipy_tests/test_network.py:293-420 and test_nodes.py)."""
import os

import numpy as np
import pytest
import torch

from golden_util import GOLDEN, rel_err, orc

pytestmark = pytest.mark.gpu

NODE = "neuron_model_templates.rate_neurons.leaky_integrator.tanh"


def _rate_net(n=10, dt=1e-2, seed=0, **kw):
    import rectipy_b200 as rp
    rng = np.random.default_rng(seed)
    W = rng.standard_normal((n, n))
    net = rp.Network(dt=dt, device="cuda:0")
    node = net.add_diffeq_node("rnn", NODE, weights=W, input_var="li_op/I_ext", output_var="li_op/v",
                               source_var="tanh_op/r", target_var="li_op/r_in", **kw)
    return net, node, W


def test_run_equals_manual_forward_loop():
    """rectipy_tests/test_network.py:293-339: run(sampling_steps=2) == forward loop with window means; recorded
    variables == get_var after each forward."""
    n, steps = 10, 100
    x = torch.randn(steps, n, generator=torch.Generator().manual_seed(1))
    net1, _, _ = _rate_net(n)
    net2, _, _ = _rate_net(n)
    net3, _, _ = _rate_net(n)
    net3.compile()
    res1 = net1.run(inputs=x, sampling_steps=2, verbose=False)
    res2 = net2.run(inputs=x, record_output=False, record_vars=[("rnn", "li_op/v", False)], verbose=False)
    res3, res4, buf = [], [], []
    for step in range(steps):
        out = net3.forward(x[step, :])
        buf.append(out.detach().cpu().numpy())
        if step % 2 == 0:
            res3.append(np.mean(buf, axis=0))
            buf = []
        res4.append(net3.get_var("rnn", var="li_op/v").detach().cpu().numpy().copy())
    a, b = res1.to_dataframe("out").values.flatten(), np.asarray(res3).flatten()
    assert np.max(np.abs(a - b)) < 1e-5
    a, b = res2.to_dataframe(("rnn", "li_op/v")).values.flatten(), np.asarray(res4).flatten()
    assert np.max(np.abs(a - b)) < 1e-5


def test_node_forward_and_state_protocol():
    net, node, W = _rate_net(12)
    assert len(node.y) == 12 and node.n_in == 12 and node.n_out == 12
    x = torch.ones(12)
    y_before = node.y.clone()
    out = node.forward(x)                      # returns the PRE-update slice (nodes.py:166-170)
    assert torch.allclose(out.cpu(), y_before.cpu())
    dt, tau = 1e-2, 10.0
    expect = y_before.cpu() + dt * (-y_before.cpu() / tau + torch.tensor(W, dtype=torch.float32) @ torch.tanh(y_before.cpu()) + 1.0)
    assert torch.allclose(node.y.cpu(), expect, atol=1e-6)
    with pytest.raises(RuntimeError):
        node.forward(torch.ones(5))
    snap = net.state
    node.forward(x)
    net.reset(snap)
    assert torch.allclose(node.y.cpu(), expect, atol=1e-6)
    net.reset()
    assert float(node.y.abs().sum()) == 0.0     # reset() -> zeros (nodes.py:199-200)
    with pytest.raises(KeyError):
        net.get_var("rnn", "nonexistent")


def test_spiking_node_len_and_reset_quirk():
    import rectipy_b200 as rp
    n = 16
    net = rp.Network(1e-3, device="cuda:0")
    node = net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=np.zeros((n, n)),
                               source_var="s", target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike",
                               reset_var="v", op="qif_op")
    assert isinstance(node, rp.SpikeResetNet) and len(node.y) == 2 * n      # test_nodes.py:71
    assert torch.allclose(net.get_var("qif", "v").cpu(), torch.full((n,), -2.0))
    # one neuron driven over threshold: s jumps by exactly 1 one step after v >= theta, v goes to v_reset
    node.reset(torch.cat([torch.full((n,), 150.0), torch.zeros(n)]))
    node.forward(torch.zeros(n))
    assert torch.allclose(net.get_var("qif", "v").cpu(), torch.full((n,), -100.0))
    assert torch.allclose(net.get_var("qif", "s").cpu(), torch.ones(n))


def test_fit_bptt_ridge_rls_recover_readout():
    """rectipy_tests/test_network.py:342-420 scaled down: each fitter recovers a 3 x 10 readout."""
    import rectipy_b200 as rp
    rng = np.random.default_rng(3)
    n, k, steps, dt = 10, 3, 400, 1e-2
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    w_true = rng.standard_normal((k, n))
    x = rng.standard_normal((steps, n)).astype(np.float32)

    def make(train=None, weights=None):
        net = rp.Network(dt=dt, device="cuda:0")
        net.add_diffeq_node("rnn", NODE, weights=W, input_var="li_op/I_ext", output_var="li_op/v",
                            source_var="tanh_op/r", target_var="li_op/r_in", node_vars={"li_op/tau": 1.0})
        if train is not False:
            net.add_func_node("out", k, "identity")
            net.add_edge("rnn", "out", weights=weights, train=train)
        return net

    target_net = make(None, w_true)
    targets = target_net.run(x, sampling_steps=1, verbose=False, enable_grad=False).to_numpy("out")

    net_ridge = make(False)
    obs = net_ridge.fit_ridge(x, targets, sampling_steps=1, alpha=1e-6, verbose=False, add_readout_node=True)
    w_ridge = obs["w_out"].T.cpu().numpy()
    assert np.mean((w_ridge - w_true) ** 2) < 0.5
    fit = net_ridge.run(x, sampling_steps=1, verbose=False, enable_grad=False).to_numpy("out")
    assert fit.shape == targets.shape

    net_rls = make("rls", None)
    net_rls.fit_rls(x, targets, update_steps=1, sampling_steps=10, verbose=False)
    w_rls = net_rls.get_edge("rnn", "out").weights.cpu().numpy()
    assert np.mean((w_rls - w_true) ** 2) < 0.5

    net_gd = make("gd", np.zeros((k, n)))
    net_gd.fit_bptt([x] * 150, [targets] * 150, optimizer="adam", lr=0.1, sampling_steps=1, verbose=False)
    w_gd = net_gd.get_edge("rnn", "out").weights.detach().cpu().numpy()
    assert np.mean((w_gd - w_true) ** 2) < 0.5


def test_rls_and_ridge_fixtures():
    from rectipy_b200 import engine
    z = np.load(os.path.join(GOLDEN, "edges.npz"))
    X = torch.tensor(z["xs"], dtype=torch.float32, device="cuda")
    Y = torch.tensor(z["ys"], dtype=torch.float32, device="cuda")
    n_in, n_out = X.shape[1], Y.shape[1]
    Wt = torch.zeros((n_out, n_in), device="cuda")
    P = float(z["rls_alpha"]) * torch.eye(n_in, device="cuda")
    loss, pred = engine.rls_run(X, Y, Wt, P, 1.0 / float(z["rls_beta"]))
    assert rel_err(Wt.cpu().numpy(), z["rls_w"][-1]) < 2e-4
    assert rel_err(P.cpu().numpy(), z["rls_p"][-1]) < 2e-4
    assert rel_err(loss.cpu().numpy(), z["rls_loss"]) < 2e-4

    import rectipy_b200 as rp
    r = np.load(os.path.join(GOLDEN, "ridge.npz"))
    net = rp.Network(float(r["dt"]), device="cuda:0")
    net.add_diffeq_node("rnn", NODE, weights=r["W"], input_var="li_op/I_ext", output_var="li_op/v",
                        source_var="tanh_op/r", target_var="li_op/r_in", node_vars={"li_op/tau": 1.0})
    net.add_func_node("inp", r["w_in"].shape[1], "identity")
    net.add_edge("inp", "rnn", weights=r["w_in"])
    obs = net.fit_ridge(r["inputs"], r["targets"], sampling_steps=1, alpha=float(r["alpha"]), verbose=False,
                        add_readout_node=False)
    assert rel_err(obs.to_numpy("out"), r["X"]) < 1e-5
    assert rel_err(obs["y"].cpu().numpy(), r["y"]) < 5e-3       # normal equations in fp32 vs fp64
    assert obs["w_out"].shape == r["w_out"].shape


def test_truncated_bptt_array_mode_runs():
    """fit_bptt with array inputs (network.py:1016-1048): optimizer step + detach every update_steps."""
    import rectipy_b200 as rp
    rng = np.random.default_rng(2)
    n, k, T = 20, 2, 300
    net = rp.Network(1e-2, device="cuda:0")
    net.add_diffeq_node("rnn", NODE, weights=rng.standard_normal((n, n)) / np.sqrt(n), input_var="li_op/I_ext",
                        output_var="li_op/v", source_var="tanh_op/r", target_var="li_op/r_in", train_params=["weights"])
    net.add_func_node("out", k, "identity")
    net.add_edge("rnn", "out", train="gd")
    w0 = net.get_node("rnn")["weights"].detach().clone()
    obs = net.fit_bptt(rng.standard_normal((T, n)), rng.standard_normal((T, k)), optimizer="sgd", lr=1e-2,
                       update_steps=50, sampling_steps=10, verbose=False)
    assert len(obs["steps"]) == 30 and len(obs["loss"]) == 30
    assert not torch.equal(w0, net.get_node("rnn")["weights"].detach())


def test_two_node_feed_forward_chain_matches_reference():
    """inp -> Linear -> LI-tanh -> Linear -> QIF -> Linear -> out: the reference executes such chains node by node inside its
    step loop (network.py:962-973); here every node is one whole-horizon engine call and autograd chains them.  Records of
    both nodes, the window means and ALL gradients (two recurrent matrices, tau, eta, three edges) vs the fixture generated
    by the reference's own Network (oracle/make_golden.py::save_two_node_chain)."""
    import ast
    import os
    import rectipy_b200 as rp
    from golden_util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "two_node_chain.npz"))
    m = ast.literal_eval(str(z["meta"]))
    net = rp.Network(m["dt"], device="cuda:0")
    rate = net.add_diffeq_node("rate", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=z["in_W1"],
                               source_var="tanh_op/r", target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v",
                               node_vars={"li_op/tau": z["p1_tau"], "li_op/k": m["p1_k"], "li_op/eta": m["p1_eta"]},
                               train_params=["weights", "li_op/tau"])
    spk = net.add_diffeq_node("spk", "neuron_model_templates.spiking_neurons.qif.qif", weights=z["in_W2"], source_var="s",
                              target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op",
                              node_vars={"eta": z["p2_eta"], "k": m["p2_k"], "tau_s": m["p2_tau_s"]}, train_params=["weights", "eta"])
    net.add_func_node("inp", m["m"], "identity"); net.add_func_node("out", m["k"], "identity")
    net.add_edge("inp", "rate", weights=z["in_w_in"], train="gd")
    net.add_edge("rate", "spk", weights=z["in_w12"], train="gd")
    net.add_edge("spk", "out", weights=z["in_w_out"], train="gd")
    obs = net.run(z["in_inputs"], sampling_steps=m["S"], cutoff=m["cutoff"], verbose=False, enable_grad=True,
                  record_vars=[("rate", "v", False), ("spk", "s", True)])
    out = torch.stack(obs["out"])
    loss = torch.nn.functional.mse_loss(out, torch.tensor(z["in_targets"], dtype=torch.float32, device="cuda"))
    loss.backward()
    ref = {k[len("float64_"):]: z[k] for k in z.files if k.startswith("float64_")}
    tol32 = lambda key: 10.0 * max(rel_err(z["float32_" + key], ref[key]), 1e-6)      # as close to fp64 as the reference's own fp32 run (x10)
    assert list(np.asarray(obs["steps"])) == list(ref["steps"])
    assert rel_err(out.detach().cpu().numpy(), ref["out"]) <= tol32("out")
    assert rel_err(obs.to_numpy(("rate", "v")), ref["var_rate_v"]) <= tol32("var_rate_v")
    assert rel_err(obs.to_numpy(("spk", "s")), ref["var_spk_s"]) <= tol32("var_spk_s")
    assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    got = dict(grad_W1=rate["weights"].grad, grad_tau1=rate["li_op/tau"].grad, grad_W2=spk["weights"].grad, grad_eta2=spk["eta"].grad,
               grad_w_in=net.get_edge("inp", "rate").weights.grad, grad_w12=net.get_edge("rate", "spk").weights.grad,
               grad_w_out=net.get_edge("spk", "out").weights.grad)
    for key, g in got.items():
        assert rel_err(g.cpu().numpy().reshape(ref[key].shape), ref[key]) <= tol32(key), key
    # single steps through the whole chain keep working (user loops / fit_rls style)
    o1 = net.forward(z["in_inputs"][0])
    assert o1.shape[-1] == m["k"] and torch.isfinite(o1).all()


@pytest.mark.parametrize("variant", ["A", "B"])
def test_feedback_network_matches_reference(variant):
    """`FeedbackNetwork` (rectipy/network.py:1196-1357): inp -> p1 -> p2 -> out with a feedback edge p2 -> p1, against the fixture
    produced by the reference's own FeedbackNetwork (oracle/make_golden.py::save_feedback_net).  Variant A: LI-tanh -> QIF, the
    feedback reads the spiking node's stale `y` (one step of delay, nodes.py:387); variant B: QIF -> LI-tanh, the feedback reads the
    rate node's current state.  Window means, records of both nodes and ALL gradients incl. the feedback edge."""
    import ast
    import os
    import rectipy_b200 as rp
    from golden_util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "feedback_net.npz"))
    m = ast.literal_eval(str(z["meta"]))
    zi = lambda key: z[f"{variant}_{key}"]
    net = rp.FeedbackNetwork(m["dt"], device="cuda:0")

    def add_rate():
        return net.add_diffeq_node("rate", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=zi("in_Wr"),
                                   source_var="tanh_op/r", target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v",
                                   node_vars={"li_op/tau": zi("pr_tau"), "li_op/k": m["pr_k"], "li_op/eta": m["pr_eta"]},
                                   train_params=["weights", "li_op/tau"])

    def add_spk():
        return net.add_diffeq_node("spk", "neuron_model_templates.spiking_neurons.qif.qif", weights=zi("in_Wq"), source_var="s",
                                   target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op",
                                   node_vars={"eta": zi("pq_eta"), "k": m["pq_k"], "tau_s": m["pq_tau_s"]}, train_params=["weights", "eta"])
    if variant == "A":
        rate, spk = add_rate(), add_spk()
        first, second = "rate", "spk"
    else:
        spk, rate = add_spk(), add_rate()
        first, second = "spk", "rate"
    net.add_func_node("inp", m["m"], "identity"); net.add_func_node("out", m["k"], "identity")
    net.add_edge("inp", first, weights=zi("in_w_in"), train="gd")
    net.add_edge(first, second, weights=zi("in_w12"), train="gd")
    net.add_edge(second, "out", weights=zi("in_w_out"), train="gd")
    net.add_edge(second, first, weights=zi("in_wfb"), train="gd", feedback=True)
    obs = net.run(zi("in_inputs"), sampling_steps=m["S"], cutoff=m["cutoff"], verbose=False, enable_grad=True,
                  record_vars=[("rate", "v", False), ("spk", "s", True)])
    out = torch.stack(obs["out"])
    loss = torch.nn.functional.mse_loss(out, torch.tensor(zi("in_targets"), dtype=torch.float32, device="cuda"))
    loss.backward()
    ref = {k[len(f"{variant}_float64_"):]: z[k] for k in z.files if k.startswith(f"{variant}_float64_")}
    tol32 = lambda key: 10.0 * max(rel_err(zi("float32_" + key), ref[key]), 1e-6)     # as close to fp64 as the reference's own fp32 run (x10)
    assert list(np.asarray(obs["steps"])) == list(ref["steps"])
    assert rel_err(out.detach().cpu().numpy(), ref["out"]) <= tol32("out")
    assert rel_err(obs.to_numpy(("rate", "v")), ref["var_rate_v"]) <= tol32("var_rate_v")
    assert rel_err(obs.to_numpy(("spk", "s")), ref["var_spk_s"]) <= tol32("var_spk_s")
    assert abs(float(loss.detach()) - float(ref["loss"])) <= 1e-4 * abs(float(ref["loss"]))
    got = dict(grad_Wr=rate["weights"].grad, grad_tau=rate["li_op/tau"].grad, grad_Wq=spk["weights"].grad, grad_eta=spk["eta"].grad,
               grad_w_in=net.get_edge("inp", first).weights.grad, grad_w12=net.get_edge(first, second).weights.grad,
               grad_wfb=net.get_edge(second, first).weights.grad, grad_w_out=net.get_edge(second, "out").weights.grad)
    for key, g in got.items():
        assert g is not None, key
        assert rel_err(g.cpu().numpy().reshape(ref[key].shape), ref[key]) <= tol32(key), key
    assert len(list(net.parameters())) == 8
    o1 = net.forward(zi("in_inputs")[0])                       # single steps keep working (user loops)
    assert o1.shape[-1] == m["k"] and torch.isfinite(o1).all()


def test_feedback_network_trial_axis_and_truncation():
    """FeedbackNetwork with a trial axis: trials are independent, so a batch of different inputs equals the one-trial runs; a
    truncated run (detach every 50 steps) keeps the forward result and still delivers gradients to the feedback edge."""
    import rectipy_b200 as rp
    rng = np.random.default_rng(5)
    n1, n2, T, B, dt = 12, 9, 160, 3, 1e-3
    W1, W2 = rng.standard_normal((n1, n1)) * 2.0 / np.sqrt(n1), rng.standard_normal((n2, n2)) * 2.0 / np.sqrt(n2)
    ff, fb = rng.standard_normal((n2, n1)) * 3.0, rng.standard_normal((n1, n2)) * 3.0
    x = (rng.standard_normal((T, B, n1)) * 5.0 + 30.0).astype(np.float32)
    kw = dict(source_var="s", target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op")

    def make(batch, trial=0):
        net = rp.FeedbackNetwork(dt, device="cuda:0", batch=batch)
        net.add_diffeq_node("p1", "neuron_model_templates.spiking_neurons.qif.qif", weights=W1, node_vars={"eta": 10.0}, **kw)
        net.add_diffeq_node("p2", "neuron_model_templates.spiking_neurons.qif.qif", weights=W2, node_vars={"eta": 5.0}, **kw)
        net.add_edge("p1", "p2", weights=ff, train="gd")
        net.add_edge("p2", "p1", weights=fb, train="gd", feedback=True)
        # spread membrane potentials: neurons cross threshold within the horizon (otherwise s == 0 and nothing is coupled)
        srng = np.random.default_rng(17)
        for name, n in (("p1", n1), ("p2", n2)):
            y0 = np.concatenate([srng.uniform(-50.0, 99.0, (B, n)), np.zeros((B, n))], axis=1).astype(np.float32)
            net.get_node(name).reset(y0[:batch] if batch > 1 else y0[trial])
        return net
    netB = make(B)
    outB = torch.stack(netB.run(x, sampling_steps=2, verbose=False, enable_grad=False)["out"]).cpu().numpy()
    assert outB.shape == (T // 2, B, n2) and np.abs(outB).max() > 0.0
    for b in range(B):
        net1 = make(1, trial=b)
        out1 = torch.stack(net1.run(x[:, b, :], sampling_steps=2, verbose=False, enable_grad=False)["out"]).cpu().numpy()
        assert rel_err(outB[:, b, :], out1) < 1e-6, b
    net_t = make(B)
    out_t = torch.stack(net_t.run(x, sampling_steps=2, verbose=False, enable_grad=True, truncate_steps=50)["out"])
    assert rel_err(out_t.detach().cpu().numpy(), outB) < 1e-6
    out_t.square().mean().backward()
    g = net_t.get_edge("p2", "p1").weights.grad
    assert g is not None and torch.isfinite(g).all() and float(g.abs().max()) > 0.0


def test_delay_and_filter_edges_in_a_network():
    """inp --LinearMemory(delays)--> LI-tanh --LinearFilter--> out, window means with cutoff: the reference's own Network output
    (tests/golden/edges_stateful.npz).  The stateful edges map the per-step series on either side of the engine call."""
    import ast
    import os
    import rectipy_b200 as rp
    from golden_util import GOLDEN
    z = np.load(os.path.join(GOLDEN, "edges_stateful.npz"))
    m = ast.literal_eval(str(z["net_meta"]))
    net = rp.Network(m["dt"], device="cuda:0")
    net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=z["net_W"], source_var="tanh_op/r",
                        target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v",
                        node_vars={"li_op/tau": m["tau"], "li_op/k": m["k_c"], "li_op/eta": m["eta"]})
    net.add_func_node("inp", m["m"], "identity"); net.add_func_node("out", m["k"], "identity")
    e_in = net.add_edge("inp", "rnn", weights=z["net_w_in"], delays=z["net_d_in"])
    e_out = net.add_edge("rnn", "out", weights=z["net_w_out"], filter_weights=z["net_f_out"])
    assert type(e_in).__name__ == "LinearMemory" and type(e_out).__name__ == "LinearFilter"
    obs = net.run(z["net_inputs"], sampling_steps=m["S"], cutoff=m["cutoff"], verbose=False, enable_grad=False)
    assert rel_err(obs.to_numpy("out"), z["net_out"]) < 1e-5
