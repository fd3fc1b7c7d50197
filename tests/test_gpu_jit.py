"""Run-time compiled vector fields (rectipy_b200/jit.py, RP_JIT): a user template whose equations match none of the compiled fields
runs through kernels generated from its equations.  The checker is the oracle's node classes (RateNet / SpikeResetNet restatements,
oracle/rectipy_oracle.py) around the same field written by hand in torch, fp64, one trial at a time -- exactly what the reference
does with the function PyRates generates (rectipy/nodes.py:58-90,166-170,382-392).

Bars: rate field <= 1e-5 relative (outputs, recorded state, gradients); spiking field: identical spike raster, outputs <= 1e-4,
gradients <= 2e-3 (fp32 engine vs fp64 truth across threshold dynamics, as for the compiled spiking fields).
"""
import numpy as np
import pytest
import torch

from golden_util import rel_err, orc

pytestmark = pytest.mark.gpu

YAML = """
fhn_op:
  base: OperatorTemplate
  equations:
    - "v' = v - v^3/3 - w + I_ext + g*r_in"
    - "w' = (v + a - b*w)/tau_w"
    - "r = 1/(1 + exp(-beta*v))"
  variables:
    v: output(0.1)
    w: variable(0.0)
    r: variable(0.0)
    g: 0.5
    a: 0.7
    b: 0.8
    tau_w: 12.5
    beta: 2.0
    I_ext: input(0.0)
    r_in: input(0.0)
fhn:
  base: NodeTemplate
  operators:
    - fhn_op
adex_op:
  base: OperatorTemplate
  equations:
    - "v' = (E_L - v + delta*exp((v - v_T)/delta) - w + I_ext)/tau + J*s_in"
    - "w' = (a*(v - E_L) - w)/tau_w + b*spike"
    - "s' = -s/tau_s + spike"
  variables:
    s: output(0.0)
    v: variable(-1.0)
    w: variable(0.0)
    E_L: -1.0
    delta: 0.5
    v_T: 0.0
    tau: 1.0
    J: 1.0
    a: 0.2
    tau_w: 5.0
    b: 0.1
    tau_s: 0.5
    I_ext: input(0.0)
    spike: input(0.0)
    s_in: input(0.0)
adex:
  base: NodeTemplate
  operators:
    - adex_op
"""


@pytest.fixture
def user_templates(tmp_path, monkeypatch):
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "custom.yaml").write_text(YAML)
    monkeypatch.chdir(tmp_path)


def _fhn_oracle(n, W, dt, p, train):
    names = ["weights", "g", "a", "b", "tau_w", "beta", "I_ext"]

    def f(t, y, weights, g, a, b, tau_w, beta, I_ext):
        v, w = y[:n], y[n:]
        r = 1.0 / (1.0 + torch.exp(-beta * v))
        return torch.cat((v - v ** 3 / 3 - w + I_ext + g * (weights @ r), (v + a - b * w) / tau_w), 0)
    args = [torch.cat((torch.full((n,), 0.1), torch.zeros(n))).double(), torch.tensor(W)]
    args += [torch.as_tensor(np.atleast_1d(p[k]), dtype=torch.float64).clone() for k in names[1:-1]] + [torch.zeros(n, dtype=torch.float64)]
    pm = {k: i for i, k in enumerate(names)}
    pm["in"] = pm["I_ext"]
    vm = {"v": (0, n), "w": (n, 2 * n), "out": (0, n)}
    return orc.OracleRateNode(f, args, vm, pm, dt, torch.float64, train)


def _adex_oracle(n, W, dt, p, train, thresh, reset):
    names = ["weights", "E_L", "delta", "v_T", "tau", "J", "a", "tau_w", "b", "tau_s", "I_ext", "spike"]

    def f(t, y, weights, E_L, delta, v_T, tau, J, a, tau_w, b, tau_s, I_ext, spike):
        v, w, s = y[:n], y[n:2 * n], y[2 * n:]
        dv = (E_L - v + delta * torch.exp((v - v_T) / delta) - w + I_ext) / tau + J * (weights @ s)
        return torch.cat((dv, (a * (v - E_L) - w) / tau_w + b * spike, -s / tau_s + spike), 0)
    args = [torch.cat((torch.full((n,), -1.0), torch.zeros(2 * n))).double(), torch.tensor(W)]
    args += [torch.as_tensor(np.atleast_1d(p[k]), dtype=torch.float64).clone() for k in names[1:-2]]
    args += [torch.zeros(n, dtype=torch.float64), torch.zeros(n, dtype=torch.float64)]
    pm = {k: i for i, k in enumerate(names)}
    pm["in"], pm["spike_var"] = pm["I_ext"], pm["spike"]
    vm = {"v": (0, n), "w": (n, 2 * n), "s": (2 * n, 3 * n), "out": (2 * n, 3 * n), "reset_var": (0, n)}
    return orc.OracleSpikeResetNode(f, args, vm, pm, dt, torch.float64, train, spike_threshold=thresh, spike_reset=reset)


@pytest.mark.parametrize("B", [1, 3, 20])
def test_rate_field_compiled_at_run_time(user_templates, B):
    import rectipy_b200 as rp
    n, m, k, T, S, dt = 45, 2, 3, 150, 3, 0.05
    rng = np.random.default_rng(7 + B)
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)) * 0.3, rng.standard_normal((k, n)) / np.sqrt(n)
    p = dict(g=rng.uniform(0.4, 0.6, n), a=rng.uniform(0.5, 0.9, n), b=0.8, tau_w=12.5, beta=2.0)      # tau_w: shared scalar, a / g: per neuron
    t = np.arange(T) * dt
    x = 0.5 * np.sin(2 * np.pi * rng.uniform(0.05, 0.3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + 0.3
    targets = rng.standard_normal((len(range(0, T, S)), B, k))

    net = rp.Network(dt, device="cuda:0", batch=B)
    node = net.add_diffeq_node("rnn", "mymodels.custom.fhn", weights=W, source_var="fhn_op/r", target_var="fhn_op/r_in",
                               input_var="fhn_op/I_ext", output_var="fhn_op/v", node_vars={"fhn_op/a": p["a"], "fhn_op/g": p["g"]},
                               train_params=["weights", "fhn_op/a", "fhn_op/tau_w", "fhn_op/g"])
    assert node.spec.model == rp._cabi.RP_JIT and node.spec.jit_program.src_plane == -1
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", "w", False)])
    out = torch.stack(obs["out"]).reshape(-1, B, k)          # batch=1 networks have no trial axis
    rec_w = obs.to_numpy(("rnn", "w")).reshape(out.shape[0], B, n)
    loss = torch.nn.functional.mse_loss(out, torch.tensor(targets, dtype=torch.float32, device="cuda:0"))
    loss.backward()

    train = ["weights", "a", "tau_w", "g"]
    g_ref, out_ref, w_ref = None, np.zeros(out.shape), np.zeros(rec_w.shape)
    for b in range(B):
        onode = _fhn_oracle(n, W, dt, p, train)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, requires_grad=True), w_out=torch.tensor(w_out, requires_grad=True))
        r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=True, record_vars=[("w", False)])
        pred = torch.stack(r["out"])
        out_ref[:, b, :] = pred.detach().numpy()
        w_ref[:, b, :] = torch.stack(r["vars"]["w"]).detach().numpy()
        (torch.nn.functional.mse_loss(pred, torch.tensor(targets[:, b, :]), reduction="sum") / out.numel()).backward()
        gs = [q.grad.numpy().copy() for q in onet.parameters()]
        g_ref = gs if g_ref is None else [a + c for a, c in zip(g_ref, gs)]
    g_eng = [node["weights"].grad, node["fhn_op/a"].grad, node["fhn_op/tau_w"].grad, node["fhn_op/g"].grad,
             net.get_edge("inp", "rnn").weights.grad, net.get_edge("rnn", "out").weights.grad]
    names = ["weights", "a", "tau_w", "g", "w_in", "w_out"]
    errs = {nm: rel_err(ge.detach().cpu().numpy().reshape(gr.shape), gr) for nm, ge, gr in zip(names, g_eng, g_ref)}
    errs["out"] = rel_err(out.detach().cpu().numpy(), out_ref)
    errs["rec_w"] = rel_err(rec_w, w_ref)
    print("jit fhn B=%d" % B, {k_: f"{v:.2e}" for k_, v in errs.items()})
    assert all(v <= 1e-5 for v in errs.values()), errs


@pytest.mark.parametrize("B,truncate", [(2, None), (5, 40)])
def test_spiking_field_compiled_at_run_time(user_templates, B, truncate):
    import rectipy_b200 as rp
    n, m, k, T, S, dt = 40, 2, 2, 400, 4, 0.01
    thresh, reset = 2.0, -1.5
    rng = np.random.default_rng(11 + B)
    W = rng.standard_normal((n, n)) * 1.5 / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    p = dict(E_L=-1.0, delta=0.5, v_T=0.0, tau=rng.uniform(0.8, 1.2, n), J=1.0, a=0.2, tau_w=5.0, b=0.1, tau_s=0.5)
    t = np.arange(T) * dt
    x = 1.0 * np.sin(2 * np.pi * rng.uniform(0.3, 2, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + 1.2
    targets = rng.standard_normal((len(range(0, T, S)), B, k))

    net = rp.Network(dt, device="cuda:0", batch=B)
    node = net.add_diffeq_node("rnn", "mymodels.custom.adex", weights=W, source_var="s", target_var="s_in", input_var="I_ext",
                               output_var="s", spike_var="spike", reset_var="v", node_vars={"adex_op/tau": p["tau"]},
                               train_params=["weights", "adex_op/b", "adex_op/tau"], spike_threshold=thresh, spike_reset=reset)
    assert node.spec.model == rp._cabi.RP_JIT and node.spec.jit_program.spiking
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    kw = {} if truncate is None else dict(truncate_steps=truncate)
    obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", "v", False)], **kw)
    out = torch.stack(obs["out"]).reshape(-1, B, k)
    rec_v = obs.to_numpy(("rnn", "v")).reshape(out.shape[0], B, n)
    torch.nn.functional.mse_loss(out, torch.tensor(targets, dtype=torch.float32, device="cuda:0")).backward()

    train = ["weights", "b", "tau"]
    g_ref, out_ref, v_ref = None, np.zeros(out.shape), np.zeros(rec_v.shape)
    for b in range(B):
        onode = _adex_oracle(n, W, dt, p, train, thresh, reset)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, requires_grad=True), w_out=torch.tensor(w_out, requires_grad=True))
        r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=True, record_vars=[("v", False)], truncate_steps=truncate)
        pred = torch.stack(r["out"])
        out_ref[:, b, :] = pred.detach().numpy()
        v_ref[:, b, :] = torch.stack(r["vars"]["v"]).detach().numpy()
        (torch.nn.functional.mse_loss(pred, torch.tensor(targets[:, b, :]), reduction="sum") / out.numel()).backward()
        gs = [q.grad.numpy().copy() for q in onet.parameters()]
        g_ref = gs if g_ref is None else [a + c for a, c in zip(g_ref, gs)]
    # the recorded v of a SpikeResetNet is the pre-update value: v >= theta marks a spike at that record step
    assert np.array_equal(rec_v >= thresh, v_ref >= thresh) and (v_ref >= thresh).sum() > 0
    g_eng = [node["weights"].grad, node["adex_op/b"].grad, node["adex_op/tau"].grad,
             net.get_edge("inp", "rnn").weights.grad, net.get_edge("rnn", "out").weights.grad]
    names = ["weights", "b", "tau", "w_in", "w_out"]
    errs = {nm: rel_err(ge.detach().cpu().numpy().reshape(gr.shape), gr) for nm, ge, gr in zip(names, g_eng, g_ref)}
    e_out = rel_err(out.detach().cpu().numpy(), out_ref)
    print("jit adex B=%d" % B, f"out {e_out:.2e}", {k_: f"{v:.2e}" for k_, v in errs.items()})
    assert e_out <= 1e-4
    assert all(v <= 2e-3 for v in errs.values()), errs


def test_jit_rejects_what_it_cannot_express(user_templates):
    import rectipy_b200 as rp
    net = rp.Network(1e-3, device="cuda:0")
    with pytest.raises(NotImplementedError):          # output must be a state variable
        net.add_diffeq_node("a", "mymodels.custom.fhn", weights=np.zeros((4, 4)), source_var="r", target_var="r_in",
                            input_var="I_ext", output_var="r")
    with pytest.raises(NotImplementedError):          # tensor-core precisions are for the compiled fields
        net2 = rp.Network(1e-3, device="cuda:0", batch=128, precision="3xf16")
        net2.add_diffeq_node("a", "mymodels.custom.fhn", weights=np.zeros((128, 128)), source_var="r", target_var="r_in",
                             input_var="I_ext", output_var="v")
        net2.run(np.zeros((3, 128, 128)), verbose=False)


# ---- MultiSpikeResetNet (rectipy/nodes.py:404-465): several spike / reset variable pairs, post-update outputs ------------------
EI_YAML = """
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


@pytest.fixture
def ei_template(tmp_path, monkeypatch):
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "twopop.yaml").write_text(EI_YAML)
    monkeypatch.chdir(tmp_path)


def _ei_engine(rp, z_or_none, n, W, dt, B, node_vars, train, thresh, reset, w_in, w_out):
    net = rp.Network(dt, device="cuda:0", batch=B)
    node = net.add_diffeq_node("rnn", "mymodels.twopop.ei", weights=W, source_var="s_e", target_var="s_in", input_var="I_ext",
                               output_var="s_e", spike_var=["spike_e", "spike_i"], reset_var=["v_e", "v_i"], op="ei_op",
                               node_vars=node_vars, train_params=train, spike_threshold=thresh, spike_reset=reset)
    assert type(node).__name__ == "MultiSpikeResetNet" and node.spec.jit_program.spiking == 2 and node.spec.jit_program.post_out
    net.add_func_node("inp", w_in.shape[1], "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
    net.add_func_node("out", w_out.shape[0], "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    return net, node


def test_multi_spike_reset_net_matches_reference_fixture(ei_template):
    """Fixture minted by the reference's own MultiSpikeResetNet inside its Network (oracle/make_golden.py::save_multispike)."""
    import os
    import rectipy_b200 as rp
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "multispike_ei.npz"))
    import ast
    meta = ast.literal_eval(str(z["meta"]))
    n, dt, S = meta["n"], meta["dt"], meta["S"]
    node_vars = {"eta_e": z["param_eta_e"], "eta_i": z["param_eta_i"], **{q: meta[q] for q in ("tau_e", "J_ee", "J_ei", "tau_i", "J_ie", "tau_s")}}
    net, node = _ei_engine(rp, z, n, z["in_W"], dt, 1, node_vars, meta["train"], meta["thresh"], meta["reset"], z["in_w_in"], z["in_w_out"])
    obs = net.run(z["in_inputs"], sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", "v_e", False), ("rnn", "v_i", False)])
    out = torch.stack(obs["out"])
    loss = torch.nn.MSELoss()(out, torch.tensor(z["in_targets"], dtype=torch.float32, device="cuda:0"))
    loss.backward()
    res = dict(out=out.detach().cpu().numpy(), var_v_e=obs.to_numpy(("rnn", "v_e")), var_v_i=obs.to_numpy(("rnn", "v_i")),
               y_final=node.y.detach().cpu().numpy(), loss=loss.detach().cpu().numpy(),
               grad_w_in=net.get_edge("inp", "rnn").weights.grad.cpu().numpy(), grad_w_out=net.get_edge("rnn", "out").weights.grad.cpu().numpy())
    for name in meta["train"]:
        res[f"grad_{name}"] = node[name].grad.detach().cpu().numpy()
    assert np.array_equal(np.asarray(obs["steps"]), z["float64_steps"])
    # post-update records: a reset shows as exactly the reset value -- the same (step, neuron) entries must have been reset
    for v in ("var_v_e", "var_v_i"):
        assert np.array_equal(res[v].reshape(z["float64_" + v].shape) == meta["reset"], z["float64_" + v] == meta["reset"])
    report = {}
    for key, val in res.items():
        ref = z["float64_" + key]
        err = rel_err(val.reshape(ref.shape), ref)
        bar = max(2e-5, 10.0 * rel_err(z["float32_" + key], ref))
        report[key] = (err, bar)
    print("multispike fixture", {k_: f"{e:.2e}/{b:.1e}" for k_, (e, b) in report.items()})
    assert not {k_: v for k_, v in report.items() if not v[0] <= v[1]}


def test_multi_spike_reset_net_batched_matches_per_trial_oracle(ei_template):
    import rectipy_b200 as rp
    n, m, k, T, S, dt, B = 30, 2, 3, 600, 2, 1e-3, 4
    thresh, reset = 50.0, -50.0
    rng = np.random.default_rng(99)
    W = rng.standard_normal((n, n)) * 2.0 / np.sqrt(n)
    params = dict(eta_e=orc.lorentzian_etas(n) + 8.0, eta_i=rng.uniform(-2.0, 6.0, n), tau_e=1.0, J_ee=1.3, J_ei=2.0, tau_i=0.5, J_ie=3.0, tau_s=0.8)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    t = np.arange(T) * dt
    x = 10.0 * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + 12.0
    targets = rng.standard_normal((len(range(0, T, S)), B, k))
    train = ["weights", "eta_e", "J_ie"]
    net, node = _ei_engine(rp, None, n, W, dt, B, params, train, thresh, reset, w_in, w_out)
    obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", "v_i", False)])
    out = torch.stack(obs["out"])
    rec = obs.to_numpy(("rnn", "v_i")).reshape(out.shape[0], B, n)
    torch.nn.functional.mse_loss(out, torch.tensor(targets, dtype=torch.float32, device="cuda:0")).backward()
    g_ref, out_ref, rec_ref = None, np.zeros(out.shape), np.zeros(rec.shape)
    for b in range(B):
        onode = orc.OracleMultiSpikeResetNode(*orc.build_ei_node_args(n, W, params, torch.float64), dt, torch.float64, train,
                                              spike_threshold=thresh, spike_reset=reset)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, requires_grad=True), w_out=torch.tensor(w_out, requires_grad=True))
        r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=True, record_vars=[("v_i", False)])
        pred = torch.stack(r["out"])
        out_ref[:, b, :] = pred.detach().numpy()
        rec_ref[:, b, :] = torch.stack(r["vars"]["v_i"]).detach().numpy()
        (torch.nn.functional.mse_loss(pred, torch.tensor(targets[:, b, :]), reduction="sum") / out.numel()).backward()
        gs = [q.grad.numpy().copy() for q in onet.parameters()]
        g_ref = gs if g_ref is None else [a + c for a, c in zip(g_ref, gs)]
    assert np.array_equal(rec == reset, rec_ref == reset) and (rec_ref == reset).sum() > 0
    g_eng = [node["weights"].grad, node["eta_e"].grad, node["J_ie"].grad,
             net.get_edge("inp", "rnn").weights.grad, net.get_edge("rnn", "out").weights.grad]
    errs = {nm: rel_err(ge.detach().cpu().numpy().reshape(gr.shape), gr) for nm, ge, gr in zip(train + ["w_in", "w_out"], g_eng, g_ref)}
    e_out = rel_err(out.detach().cpu().numpy(), out_ref)
    print("multispike batched", f"out {e_out:.2e}", {k_: f"{v:.2e}" for k_, v in errs.items()})
    assert e_out <= 1e-4 and all(v <= 2e-3 for v in errs.values()), errs


def test_generated_fields_single_steps_sweeps_and_uncoupled_nodes(user_templates):
    """The rest of the node protocol on generated kernels: `node.forward(x)` (T = 1 calls), per-trial parameter sweeps, a node without
    recurrent weights (`weights=None, N=n`, rectipy/nodes.py:134-140)."""
    import rectipy_b200 as rp
    n, dt = 12, 0.05
    rng = np.random.default_rng(3)
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    p = dict(g=0.5, a=0.7, b=0.8, tau_w=12.5, beta=2.0)
    # (a) single steps through node.forward, state protocol
    net = rp.Network(dt, device="cuda:0")
    node = net.add_diffeq_node("rnn", "mymodels.custom.fhn", weights=W, source_var="r", target_var="r_in", input_var="I_ext", output_var="v")
    onode = _fhn_oracle(n, W, dt, p, None)
    for step in range(6):
        x = rng.standard_normal(n)
        out = node.forward(torch.tensor(x, dtype=torch.float32, device="cuda:0")).cpu().numpy()
        ref = onode.forward(torch.tensor(x)).detach().numpy()
        assert rel_err(out, ref) <= 1e-5, step
    assert rel_err(node.y.cpu().numpy(), onode.y.detach().numpy()) <= 1e-5
    assert rel_err(node["w"].cpu().numpy(), onode.get("w").detach().numpy()) <= 1e-5
    # (b) a sweep: every trial its own `a` (shared over neurons) and its own per-neuron `tau_w`
    B, T = 5, 80
    a_sweep = rng.uniform(0.5, 0.9, (B, 1))
    tw_sweep = rng.uniform(8.0, 16.0, (B, n))
    net = rp.Network(dt, device="cuda:0", batch=B)
    net.add_diffeq_node("rnn", "mymodels.custom.fhn", weights=W, source_var="r", target_var="r_in", input_var="I_ext", output_var="v",
                        node_vars={"fhn_op/a": a_sweep, "fhn_op/tau_w": tw_sweep})
    x = rng.standard_normal((T, B, n)) * 0.3
    out = net.run(x, verbose=False).to_numpy("out")
    for b in range(B):
        onode = _fhn_oracle(n, W, dt, dict(p, a=float(a_sweep[b, 0]), tau_w=tw_sweep[b]), None)
        ref = torch.stack(orc.OracleNet(onode).run(torch.tensor(x[:, b, :]), enable_grad=False)["out"]).numpy()
        assert rel_err(out[:, b, :], ref) <= 1e-5, b
    # (c) no recurrent weights
    net = rp.Network(dt, device="cuda:0")
    net.add_diffeq_node("rnn", "mymodels.custom.fhn", N=n, input_var="I_ext", output_var="w")
    x1 = rng.standard_normal((40, n))
    out = net.run(x1, verbose=False).to_numpy("out")
    onode = _fhn_oracle(n, np.zeros((n, n)), dt, p, None)
    onode.start, onode.stop = n, 2 * n
    ref = torch.stack(orc.OracleNet(onode).run(torch.tensor(x1), enable_grad=False)["out"]).numpy()
    assert rel_err(out, ref) <= 1e-5
