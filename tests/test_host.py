"""CPU tests of the host logic (no compute): template front-end, graph construction/compile rules, error types,
Observer, utilities.  Mirrors the structural parts of rectipy_tests/test_network.py and test_edges.py."""
import os

import numpy as np
import pytest
import torch

import rectipy_b200 as rp
from rectipy_b200 import templates, _cabi as abi
from rectipy_b200.network import _window_mean, _record_steps
from golden_util import orc

QIF = "neuron_model_templates.spiking_neurons.qif.qif"
TANH = "neuron_model_templates.rate_neurons.leaky_integrator.tanh"


def test_builtin_templates_resolve():
    for path, name, nsv, model in [(TANH, "li_tanh", 1, abi.RP_LI_TANH),
                                   ("neuron_model_templates.rate_neurons.leaky_integrator.sigmoid", "li_sigmoid", 1, abi.RP_LI_SIGMOID),
                                   (QIF, "qif", 2, abi.RP_QIF), (QIF + "_sfa", "qif_sfa", 3, abi.RP_QIF_SFA),
                                   ("neuron_model_templates.spiking_neurons.lif.lif", "lif", 2, abi.RP_LIF),
                                   ("neuron_model_templates.spiking_neurons.ik.ik", "ik", 3, abi.RP_IK),
                                   ("neuron_model_templates.spiking_neurons.ik.iku", "iku", 3, abi.RP_IKU),
                                   ("neuron_model_templates.spiking_neurons.ik.ik_biexp", "ik_biexp", 4, abi.RP_IK_BIEXP)]:
        spec = templates.resolve_template(path)
        assert (spec.name, spec.n_sv, spec.model) == (name, nsv, model)
    # defaults of the reference YAML files
    spec = templates.resolve_template(QIF + "_sfa")
    assert spec.params["qif_sfa_op/tau_x"][1] == 10.0 and spec.params["qif_sfa_op/eta"][1] == -5.0
    assert dict(spec.state_vars)["qif_sfa_op/v"] == -2.0
    assert templates.resolve_template("neuron_model_templates.spiking_neurons.lif.lif").params["lif_op/tau_s"][1] == 0.5
    with pytest.raises(FileNotFoundError):
        templates.resolve_template("neuron_model_templates.nowhere.nothing.tanh")
    with pytest.raises(AttributeError):
        templates.resolve_template("neuron_model_templates.spiking_neurons.qif.invalid")


def test_user_yaml_template(tmp_path, monkeypatch):
    """A user file in PyRates template syntax with `base:` inheritance and changed defaults is parsed and matched."""
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "custom.yaml").write_text(
        "my_qif_op:\n  base: OperatorTemplate\n  equations:\n    - \"v' = (v^2 + eta + I_ext)/tau + k*s_in\"\n"
        "    - \"s' = -s/tau_s + spike\"\n  variables:\n    s: output(0.0)\n    v: variable(-1.0)\n    tau: 2.0\n    k: 3.0\n"
        "    tau_s: 0.5\n    eta: 1.5\n    I_ext: input(0.0)\n    spike: input(0.0)\n    s_in: input(0.0)\n"
        "my_sfa_op:\n  base: my_qif_op\n  equations:\n    replace:\n      eta: eta - x\n    add:\n      - \"x' = -x/tau_x + alpha*spike\"\n"
        "  variables:\n    x: variable(0.0)\n    alpha: 0.25\n    tau_x: 7.0\n"
        "my_neuron:\n  base: NodeTemplate\n  operators:\n    - my_sfa_op\n"
        "weird_op:\n  base: OperatorTemplate\n  equations: \"v' = -v^3\"\n  variables:\n    v: output(0.0)\n"
        "weird:\n  base: NodeTemplate\n  operators:\n    - weird_op\n")
    monkeypatch.chdir(tmp_path)
    spec = templates.resolve_template("mymodels.custom.my_neuron")
    assert spec.name == "qif_sfa" and spec.ops == ("my_sfa_op",)
    assert spec.params["my_sfa_op/tau"][1] == 2.0 and spec.params["my_sfa_op/alpha"][1] == 0.25
    assert dict(spec.state_vars)["my_sfa_op/v"] == -1.0
    weird = templates.resolve_template("mymodels.custom.weird")      # no compiled field: handed to the run-time code generator
    assert weird.model == abi.RP_JIT and weird.jit_field is not None and weird.jit_program is None


def test_user_yaml_equivalent_equations_match_symbolically(tmp_path, monkeypatch):
    """A user operator whose equations are algebraically identical to a compiled field, but written differently, resolves to it."""
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "rewritten.yaml").write_text(
        "q_op:\n  base: OperatorTemplate\n  equations:\n    - \"v' = k*s_in + (I_ext + eta + v*v)/tau\"\n"
        "    - \"s' = spike - (1/tau_s)*s\"\n  variables:\n    s: output(0.0)\n    v: variable(-3.0)\n    tau: 1.5\n    k: 2.0\n"
        "    tau_s: 0.5\n    eta: 0.5\n    I_ext: input(0.0)\n    spike: input(0.0)\n    s_in: input(0.0)\n"
        "q_neuron:\n  base: NodeTemplate\n  operators:\n    - q_op\n"
        "cubic_op:\n  base: OperatorTemplate\n  equations:\n    - \"v' = k*s_in + (I_ext + eta + v*v*v)/tau\"\n"
        "    - \"s' = spike - (1/tau_s)*s\"\n  variables:\n    s: output(0.0)\n    v: variable(-3.0)\n    tau: 1.5\n    k: 2.0\n"
        "    tau_s: 0.5\n    eta: 0.5\n    I_ext: input(0.0)\n    spike: input(0.0)\n    s_in: input(0.0)\n"
        "cubic_neuron:\n  base: NodeTemplate\n  operators:\n    - cubic_op\n")
    monkeypatch.chdir(tmp_path)
    spec = templates.resolve_template("mymodels.rewritten.q_neuron")
    assert spec.name == "qif" and spec.model == abi.RP_QIF and spec.ops == ("q_op",)
    assert spec.params["q_op/tau"][1] == 1.5 and dict(spec.state_vars)["q_op/v"] == -3.0
    cubic = templates.resolve_template("mymodels.rewritten.cubic_neuron")
    assert cubic.model == abi.RP_JIT and [k for k, _ in cubic.state_vars] == ["cubic_op/v", "cubic_op/s"]
    # names that sympy would read as constants / functions stay plain symbols
    assert templates._same_equation("v'=(E_r-v)*g*s_in/C", "v'=g*s_in*(E_r-v)/C")
    assert not templates._same_equation("v'=(E_r-v)*g*s_in/C", "v'=g*s_in*(E_r+v)/C")
    assert templates._same_equation("u'=(b*(mean(v)-v_r)-u)/tau_u+kappa*mean(spike)", "u'=kappa*mean(spike)-u/tau_u+b*(mean(v)-v_r)/tau_u")


def test_ik_state_order_matches_reference():
    """ik_op: the reference's y is [v, u, s] (equation order, ik.yaml:10-13); the engine keeps planes (v, s, u)."""
    net = rp.Network(1e-1, device="cpu")
    n = 4
    node = net.add_diffeq_node("ik", "neuron_model_templates.spiking_neurons.ik.ik", weights=np.zeros((n, n)), source_var="s",
                               target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="ik_op",
                               spike_threshold=40.0, spike_reset=-60.0, node_vars={"eta": 55.0})
    assert len(node.y) == 3 * n
    assert torch.allclose(node.y, torch.cat([torch.full((n,), -60.0), torch.zeros(2 * n)]))
    node.reset(np.arange(3 * n, dtype=np.float32))
    assert torch.allclose(node["v"], torch.arange(0., n)) and torch.allclose(node["u"], torch.arange(n, 2. * n))
    assert torch.allclose(node["s"], torch.arange(2. * n, 3 * n))
    assert torch.allclose(node.state[1, 0], torch.arange(2. * n, 3 * n)) and torch.allclose(node.state[2, 0], torch.arange(n, 2. * n))
    assert torch.allclose(node.y, torch.arange(0., 3 * n))
    assert node.var_index("u") == 2 and node.var_index("ik_op/s") == 1
    slots, tensors, per_neuron = node.param_slots()
    assert abi.RP_P_G in slots and abi.RP_P_KAPPA in slots and float(node["eta"]) == 55.0 and float(node["C"]) == 100.0


def test_ik_biexp_state_order_and_time_constant_slots():
    """ik_biexp_op (ik.yaml:42-70): the reference's y is [v, u, s, x]; engine planes (v, s, u, x); tau_d / tau_r travel in the tau_s /
    tau_x slots of the ABI."""
    net = rp.Network(1e-1, device="cpu")
    n = 3
    node = net.add_diffeq_node("ik", "neuron_model_templates.spiking_neurons.ik.ik_biexp", weights=np.zeros((n, n)), source_var="s",
                               target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="ik_biexp_op",
                               spike_threshold=40.0, spike_reset=-60.0, node_vars={"tau_r": 1.5})
    assert len(node.y) == 4 * n
    assert torch.allclose(node.y, torch.cat([torch.full((n,), -60.0), torch.zeros(3 * n)]))
    node.reset(np.arange(4 * n, dtype=np.float32))
    for j, name in enumerate(("v", "u", "s", "x")):
        assert torch.allclose(node[name], torch.arange(float(j * n), float((j + 1) * n)))
    assert [node.var_index(v) for v in ("v", "s", "u", "x")] == [0, 1, 2, 3]
    assert torch.allclose(node.state[3, 0], torch.arange(3. * n, 4 * n)) and torch.allclose(node.state[1, 0], torch.arange(2. * n, 3 * n))
    assert torch.allclose(node.y, torch.arange(0., 4 * n))
    slots, tensors, per_neuron = node.param_slots()
    assert float(tensors[slots.index(abi.RP_P_TAU_X)]) == 1.5 and float(tensors[slots.index(abi.RP_P_TAU_S)]) == 6.0
    assert float(node["tau_d"]) == 6.0 and float(node["kappa"]) == 10.0


def test_feedback_network_graph_bookkeeping():
    """`FeedbackNetwork.compile` sorts the edges into the feed-forward graph (input / output node, evaluation order) and the
    feedback graph; `get_edge` / `get_node` / `parameters` see both (rectipy/network.py:1204-1328)."""
    net = rp.FeedbackNetwork(1e-3, device="cpu")
    lif = "neuron_model_templates.spiking_neurons.lif.lif"
    kw = dict(source_var="s", target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="lif_op")
    net.add_diffeq_node("p1", lif, weights=np.random.randn(6, 6), train_params=["weights"], **kw)
    net.add_diffeq_node("p2", lif, weights=np.random.randn(4, 4), **kw)
    e_ff = net.add_edge("p1", "p2", weights=np.random.rand(4, 6), train=None)
    e_fb = net.add_edge("p2", "p1", weights=-np.random.rand(6, 4), train="gd", feedback=True)
    for _ in range(2):                                  # compiling twice must not lose or duplicate the feedback edge
        net.compile()
        assert (net._in_node, net._out_node) == ("p1", "p2")
        assert list(net.graph.edges) == [("p1", "p2")] and list(net._fb_graph.edges) == [("p2", "p1")]
    assert net.get_edge("p1", "p2") is e_ff and net.get_edge("p2", "p1") is e_fb
    assert net.get_node("p1") is net._fb_graph.nodes["p1"]["node"]
    params = list(net.parameters())
    assert len(params) == 2 and any(p is e_fb.weights for p in params)          # p1's weights once + the feedback edge
    assert net._get_path() == ["p1", "p2"] and net._feedback_sources("p1") == ["p2"] and net._feedback_sources("p2") == []
    with pytest.raises(RuntimeError):                   # no CPU fallback
        net.run(np.zeros((3, 1)), verbose=False, enable_grad=False)


def _qif_net(n=10, **kw):
    net = rp.Network(1e-3, device="cpu")
    node = net.add_diffeq_node("qif", QIF, weights=np.random.randn(n, n), source_var="s", target_var="s_in",
                               input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op", **kw)
    return net, node


def test_node_construction_and_protocol():
    net, node = _qif_net(10, node_vars={"eta": np.linspace(-1, 1, 10), "tau": 2.0}, train_params=["weights", "eta"],
                         spike_threshold=50.0, spike_reset=-50.0)
    assert isinstance(node, rp.SpikeResetNet)
    assert len(node.y) == 20 and node.n_in == 10 and node.n_out == 10          # test_nodes.py:70-71
    assert node.slope == pytest.approx(1.0) and node.theta == 50.0 and node.v_reset == -50.0   # nodes.py:345-347
    assert torch.allclose(node["v"], torch.full((10,), -2.0)) and torch.allclose(node["qif_op/s"], torch.zeros(10))
    assert node["eta"].shape == (10,) and node["tau"].shape == (1,) and float(node["tau"]) == 2.0
    assert len(list(node.parameters())) == 2 and all(p.requires_grad for p in node.parameters())
    assert net.get_var("qif", "eta") is node["qif_op/eta"]
    with pytest.raises(KeyError):
        node["nonexistent"]
    slots, tensors, per_neuron = node.param_slots()
    assert per_neuron[abi.RP_P_ETA] == 1 and per_neuron[abi.RP_P_TAU] == 0
    node.set_param("k", 3.0)
    assert float(node["k"]) == 3.0
    node.reset()
    assert float(node.y.abs().sum()) == 0.0                                    # reset() -> zeros (nodes.py:199-200)
    y = np.arange(20, dtype=np.float32)
    node.reset(y)
    assert torch.allclose(node["s"], torch.arange(10, 20, dtype=torch.float32))
    rate = rp.Network(1e-2, device="cpu").add_diffeq_node("rnn", TANH, weights=np.zeros((4, 4)), input_var="li_op/I_ext",
                                                          output_var="li_op/v", source_var="tanh_op/r", target_var="li_op/r_in")
    assert isinstance(rate, rp.RateNet) and not isinstance(rate, rp.SpikeResetNet) and len(rate.y) == 4


def test_add_diffeq_node_errors():
    net = rp.Network(1e-3, device="cpu")
    with pytest.raises(ValueError):        # spiking node without reset_var (network.py:293-295)
        net.add_diffeq_node("a", QIF, weights=np.zeros((3, 3)), source_var="s", target_var="s_in", input_var="I_ext",
                            output_var="s", spike_var="spike", op="qif_op")
    with pytest.raises(ValueError):        # weights without source/target variable (nodes.py:247-249)
        net.add_diffeq_node("b", QIF, weights=np.zeros((3, 3)), input_var="I_ext", output_var="s", op="qif_op")
    with pytest.raises(NotImplementedError):   # reset=False selects the upstream-broken SpikeNet
        net.add_diffeq_node("c", QIF, weights=np.zeros((3, 3)), source_var="s", target_var="s_in", input_var="I_ext",
                            output_var="s", spike_var="spike", reset_var="v", reset=False, op="qif_op")
    with pytest.raises(KeyError):
        net.add_diffeq_node("d", QIF, weights=np.zeros((3, 3)), source_var="s", target_var="s_in", input_var="I_ext",
                            output_var="s", spike_var="spike", reset_var="v", op="qif_op", node_vars={"bogus": 1.0})
    with pytest.raises(FileNotFoundError):
        net.add_diffeq_node("e", "no_such_pkg.file.tmpl", weights=np.zeros((3, 3)), source_var="s", target_var="s_in",
                            input_var="I_ext", output_var="s")
    with pytest.raises(ValueError):
        net.add_func_node("f", 3, "not_an_activation")


def test_edges_compile_and_parameters():
    """rectipy_tests/test_network.py:108-290 structure: edge class selection, shapes, compile rules, parameter counts."""
    net, node = _qif_net(10, train_params=["weights"])
    net.add_func_node("inp", 3, "tanh")
    e_in = net.add_edge("inp", "qif")
    assert isinstance(e_in, rp.Linear) and e_in.weights.shape == (10, 3) and not e_in.weights.requires_grad   # test_network.py:179
    net.add_func_node("out", 2, "softmax")
    e_out = net.add_edge("qif", "out", weights=np.random.randn(10, 2), train="gd")    # transposed weights are accepted
    assert e_out.weights.shape == (2, 10) and e_out.weights.requires_grad
    net.compile()
    assert net._in_node == "inp" and net._out_node == "out" and len(net._bwd_graph) == 2
    assert net.n_in == 3 and net.n_out == 2
    assert len(list(net.parameters())) == 2
    chain = net._get_chain()
    assert (chain.diffeq, chain.in_func, chain.out_func) == ("qif", "inp", "out")
    with pytest.raises(ValueError):
        net.add_edge("qif", "out", weights=np.random.randn(3, 3))
    with pytest.raises(ValueError):
        net.add_edge("qif", "out", train="invalid")
    delayed = net.add_edge("qif", "out", delays=np.ones(10, dtype=np.int64))      # edge class selection (network.py:375-383)
    assert isinstance(delayed, rp.LinearMemory) and delayed.buffer.shape == (10, 2) and net._is_multi()
    with pytest.raises(ValueError):
        net.add_edge("qif", "out", delays=np.ones(4, dtype=np.int64))             # one delay per source output
    masked = net.add_edge("inp", "qif", weights=np.ones((10, 3)), mask=np.eye(10, 3))
    assert isinstance(masked, rp.LinearMasked) and float(masked.effective_weights().sum()) == 3.0
    rls = net.add_edge("qif", "out", train="rls", alpha=2.0, beta=0.9)
    assert isinstance(rls, rp.RLS) and float(rls.P[0, 0]) == 2.0 and net._train_edge == ("qif", "out")
    # ambiguous input node -> ValueError (test_network.py:237-238)
    net.add_func_node("inp2", 3, "identity")
    net.add_edge("inp2", "qif")
    with pytest.raises(ValueError):
        net.compile()
    net.pop_node("inp2")
    net.compile()
    # run on a CPU network must fail loudly, not fall back
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        net.run(np.zeros((5, 3)), verbose=False)
    # two diffeq nodes are not an engine chain
    net2, _ = _qif_net(4)
    net2.add_diffeq_node("qif2", QIF, weights=np.zeros((4, 4)), source_var="s", target_var="s_in", input_var="I_ext",
                         output_var="s", spike_var="spike", reset_var="v", op="qif_op")
    net2.add_edge("qif", "qif2")
    net2.compile()
    with pytest.raises(NotImplementedError):
        net2._get_chain()


def test_linear_edge_known_answer():
    """rectipy_tests/test_edges.py:34-92: Linear == torch.nn.Linear (no bias), shape/dtype rules."""
    w = np.random.randn(4, 9)
    lin = rp.Linear(9, 4, weights=w, dtype=torch.float64)
    ref = torch.nn.Linear(9, 4, bias=False, dtype=torch.float64)
    with torch.no_grad():
        ref.weight.copy_(torch.tensor(w))
    x = torch.randn(9, dtype=torch.float64)
    assert torch.allclose(lin.forward(x), ref(x), atol=1e-12)
    assert rp.Linear(9, 4, weights=w.T).weights.shape == (4, 9)
    assert len(list(rp.Linear(9, 4, detach=False).parameters())) == 1 and len(list(rp.Linear(9, 4).parameters())) == 0
    with pytest.raises(ValueError):
        rp.Linear(9, 4, weights=np.zeros((5, 5)))
    with pytest.raises(RuntimeError):
        lin.forward(torch.randn(8, dtype=torch.float64))
    with pytest.raises(ValueError):
        rp.RLS(3, 2, beta=1.5)
    rls, orls = rp.RLS(5, 2, dtype=torch.float64, beta=0.95, alpha=1.5), orc.OracleRLS(5, 2, beta=0.95, alpha=1.5)
    for _ in range(10):
        xx, yy = torch.randn(5, dtype=torch.float64), torch.randn(2, dtype=torch.float64)
        rls.update(xx, yy, rls.forward(xx)); orls.update(xx, yy, orls.forward(xx))
    assert torch.allclose(rls.weights, orls.weights, atol=1e-12) and torch.allclose(rls.P, orls.P, atol=1e-12)


def test_square_edge_weights_follow_the_reference_transposition():
    """rectipy/edges.py:22-23,160-161: weights (and masks) of shape (n_in, n_out) are transposed, which for a square matrix is
    ALWAYS the case -- a user's N x N matrix acts as its transpose.  Fixture: outputs of the unmodified reference classes."""
    z = np.load(os.path.join(os.path.dirname(__file__), "golden", "edges_square.npz"))
    n = z["w"].shape[0]
    lin = rp.Linear(n, n, weights=z["w"].copy(), dtype=torch.float64)
    msk = rp.LinearMasked(n, n, mask=z["mask"].copy(), weights=z["w"].copy(), dtype=torch.float64)
    assert np.array_equal(lin.weights.numpy(), z["lin_weights"]) and np.array_equal(msk.mask.numpy(), z["masked_mask"])
    assert np.array_equal(lin.weights.numpy(), z["w"].T)
    for x, y_lin, y_msk in zip(z["xs"], z["lin_out"], z["masked_out"]):
        assert np.allclose(lin.forward(torch.tensor(x)).numpy(), y_lin, atol=1e-14)
        assert np.allclose(msk.forward(torch.tensor(x)).numpy(), y_msk, atol=1e-14)


@pytest.mark.parametrize("T,S,cutoff", [(100, 1, 0), (100, 2, 0), (600, 5, 7), (50, 4, 49), (20, 7, 3), (10, 3, 12)])
def test_window_mean_matches_reference_loop(T, S, cutoff):
    """Host-side windowing (used for non-linear output nodes) == the loop of Network.run (network.py:588-597)."""
    o = torch.randn(T, 2, 3, dtype=torch.float64)
    ref, buf = [], []
    for step in range(T):
        if step >= cutoff:
            buf.append(o[step])
            if step % S == 0:
                ref.append(torch.mean(torch.stack(buf, 0), 0)); buf = []
    got = _window_mean(o, S, cutoff)
    assert got.shape[0] == len(ref) == len(_record_steps(T, S, cutoff)) == abi.load().rp_num_records(T, S, cutoff)
    if ref:
        assert torch.allclose(got, torch.stack(ref), atol=1e-12)


def test_observer_block_recording():
    obs = rp.Observer(1e-2, record_output=True, record_loss=True, record_vars=[("rnn", "v", False), ("rnn", "s", True)])
    assert obs.recorded_state_variables == [("rnn", "v"), ("rnn", "s")] and obs.reduce_flags == [False, True]
    obs.record_block([0, 2, 4], torch.arange(6.).reshape(3, 2), 0.5, [torch.ones(3, 4), torch.zeros(3)])
    assert len(obs["out"]) == 3 and torch.stack(obs["out"]).shape == (3, 2) and obs["steps"] == [0, 2, 4]
    assert obs.to_numpy(("rnn", "v")).shape == (3, 4) and obs.to_numpy("loss").tolist() == [0.5, 0.5, 0.5]
    assert obs.to_dataframe("out").shape == (3, 2)
    obs.record(6, torch.zeros(2), 0.1, [torch.ones(4), torch.ones(4)])
    assert len(obs["out"]) == 4 and float(obs[("rnn", "s")][-1]) == 1.0
    obs.save("w", 3)
    assert obs["w"] == 3


def test_utilities():
    np.random.seed(0)
    C = rp.random_connectivity(20, 30, 0.2, normalize=True)
    assert C.shape == (20, 30) and np.allclose(C.sum(1), 1.0) and ((C > 0).sum(1) == 6).all()
    assert rp.normalize(np.array([1.0, 3.0, 5.0])).tolist() == [0.0, 0.5, 1.0]
    assert rp.wta_score(np.eye(3), np.eye(3)) == 1.0
    from scipy.stats import rv_discrete
    dist = rv_discrete(values=([1, 2, 3], [0.5, 0.3, 0.2]))
    Cc = rp.circular_connectivity(12, 0.5, dist)
    assert Cc.shape == (12, 12) and np.allclose(Cc.sum(1), 1.0)
    W = rp.input_connections(10, 8, 0.5, variance=2.0, zero_mean=True)
    assert W.shape == (10, 8) and np.allclose(W.sum(1), 0.0, atol=1e-12)


def test_multi_node_chain_is_recognised():
    """Feed-forward chains of several differential-equation nodes compile to a node path (no GPU needed for the matching)."""
    import rectipy_b200 as rp
    net = rp.Network(1e-3, device="cpu")
    n1, n2 = 6, 5
    net.add_diffeq_node("a", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=np.zeros((n1, n1)),
                        source_var="tanh_op/r", target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v")
    net.add_diffeq_node("b", QIF, weights=np.zeros((n2, n2)), source_var="s", target_var="s_in", input_var="I_ext",
                        output_var="s", spike_var="spike", reset_var="v", op="qif_op")
    net.add_func_node("inp", 2, "identity"); net.add_func_node("out", 3, "sigmoid")
    net.add_edge("inp", "a"); net.add_edge("a", "b"); net.add_edge("b", "out")
    net.compile()
    assert net._is_multi() and net._get_path() == ["inp", "a", "b", "out"]
    with pytest.raises(NotImplementedError):
        net._get_chain()                       # the fused single-node plan does not apply
    net.add_func_node("side", 2, "identity"); net.add_edge("side", "b")
    with pytest.raises((NotImplementedError, ValueError)):
        net.compile(); net._get_path()        # fan-in: neither the reference nor the engine executes it


def test_stateful_edges_match_reference_fixture():
    """LinearMemory / LinearFilter / LinearMemoryFilter (rectipy/edges.py:68-147): 40 steps of the reference classes' outputs
    (tests/golden/edges_stateful.npz, oracle/make_golden.py G8b), step by step and as a batched series."""
    import os
    from golden_util import GOLDEN
    from rectipy_b200.edges import LinearMemory, LinearFilter, LinearMemoryFilter
    z = np.load(os.path.join(GOLDEN, "edges_stateful.npz"))
    n_in, n_out = z["w"].shape[1], z["w"].shape[0]
    makers = {"memory": lambda: LinearMemory(n_in, n_out, delays=z["delays"], weights=z["w"], dtype=torch.float64),
              "filter": lambda: LinearFilter(n_in, n_out, filter_weights=z["filter"], weights=z["w"], dtype=torch.float64),
              "memory_filter": lambda: LinearMemoryFilter(n_in, n_out, delays=z["delays"], filter_weights=z["filter"], weights=z["w"],
                                                          dtype=torch.float64)}
    for name, mk in makers.items():
        e = mk()
        out = np.stack([e.forward(torch.tensor(x)).detach().numpy() for x in z["xs"]])
        assert np.abs(out - z[f"{name}_out"]).max() < 1e-13, name
        series = mk().apply_series(torch.tensor(z["xs"])[:, None, :].repeat(1, 3, 1)).numpy()
        for b in range(3):
            assert np.abs(series[:, b] - z[f"{name}_out"]).max() < 1e-13, (name, b)
    with pytest.raises(ValueError):
        LinearMemory(n_in, n_out, delays=np.asarray([0, 1]), weights=z["w"])
    with pytest.raises(ValueError):
        LinearFilter(n_in, n_out, filter_weights=np.zeros((n_in, n_in + 1)), weights=z["w"])


def test_padding_helpers_are_inert_and_differentiable():
    """Host side of the padded tensor-core route (network._pad_axis / _ceil128): zero padding for weights / inputs, replication for state
    and parameters, and gradients that flow back to the real entries only as what the real entries contributed."""
    from rectipy_b200.network import _pad_axis, _ceil128
    assert [_ceil128(v) for v in (1, 127, 128, 129, 600)] == [128, 128, 128, 256, 640]
    w = torch.arange(12.0).reshape(3, 4).requires_grad_(True)
    wp = _pad_axis(_pad_axis(w, 0, 5, False), 1, 6, False)
    assert wp.shape == (5, 6) and torch.equal(wp[:3, :4], w) and float(wp[3:].abs().sum()) == 0 and float(wp[:, 4:].abs().sum()) == 0
    wp.sum().backward()
    assert torch.equal(w.grad, torch.ones_like(w))
    st = torch.arange(2 * 3 * 4.0).reshape(2, 3, 4)
    sp = _pad_axis(_pad_axis(st, 2, 6, True), 1, 5, True)
    assert sp.shape == (2, 5, 6) and torch.equal(sp[:, :3, :4], st)
    assert torch.equal(sp[:, :, 4], sp[:, :, 3]) and torch.equal(sp[:, 4], sp[:, 2])          # copies of the last real neuron / trial
    p = torch.tensor([1.0, 2.0, 3.0], requires_grad=True)
    pp = _pad_axis(p, 0, 5, True)
    (pp * torch.tensor([1.0, 1.0, 1.0, 0.0, 0.0])).sum().backward()                          # padded entries carry zero adjoint in the engine
    assert torch.equal(p.grad, torch.ones(3))
    assert _pad_axis(p, 0, 3, True) is p


def test_multi_spike_binding_checks(tmp_path, monkeypatch):
    from rectipy_b200 import jit
    from test_gpu_jit import EI_YAML
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "twopop.yaml").write_text(EI_YAML)
    monkeypatch.chdir(tmp_path)
    spec = templates.resolve_template("mymodels.twopop.ei")
    with pytest.raises(ValueError):
        jit.bind_spec(spec, "s_e", "s_in", "I_ext", ["spike_e", "spike_i"], ["v_e"])
    with pytest.raises(KeyError):
        jit.bind_spec(spec, "s_e", "s_in", "I_ext", ["spike_e", "nope"], ["v_e", "v_i"])
    with pytest.raises(ValueError):                                  # one reset variable cannot serve two spike variables
        jit.bind_spec(spec, "s_e", "s_in", "I_ext", ["spike_e", "spike_i"], ["v_e", "v_e"])
    # a compiled template asked for MultiSpikeResetNet semantics goes through the generator as well
    q = templates.resolve_template("neuron_model_templates.spiking_neurons.qif.qif", force_jit=True)
    b = jit.bind_spec(q, "s", "s_in", "I_ext", ["spike"], ["v"])
    assert b.jit_program.post_out and b.jit_program.spiking == 1 and b.planes == {"qif_op/v": 0, "qif_op/s": 1}
