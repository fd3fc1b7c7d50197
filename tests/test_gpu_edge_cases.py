"""GPU: the less-travelled options of the path, each against the fp64 CPU oracle: wide input / readout edges (torch-side
projection + dense engine I/O incl. the dense input gradient), lif driven through s_ext, spiking node read out from v, a bare
diffeq node without edges, degenerate record grids (cutoff >= T, sampling_steps > T), scalar input broadcast."""
import numpy as np
import pytest
import torch

from golden_util import TEMPLATE_PATH, rel_err, orc

pytestmark = pytest.mark.gpu


def _engine(model, n, B, dt, W, params, prec="fp32", input_var="I_ext", output_var=None, train=("weights",), skw=None):
    import rectipy_b200 as rp
    path, op, svar, tvar = TEMPLATE_PATH[model]
    net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
    kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/{input_var}",
              node_vars={f"{op}/{p}": v for p, v in params.items()}, train_params=list(train))
    if model in orc.SPIKING:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/{output_var or 's'}", **(skw or {}))
    else:
        kw.update(output_var=f"{op}/{output_var or 'v'}")
    node = net.add_diffeq_node("rnn", path, **kw)
    return net, node


@pytest.mark.parametrize("B", [1, 20])
def test_wide_input_and_readout_edges(B):
    """m = 10 > RP_MAX_IN and k = 12 > RP_MAX_OUT: the projections run in torch, the engine sees dense currents / dense
    outputs, and dL/dW_in flows through the engine's dense input gradient."""
    n, m, k, T, dt, S = 48, 10, 12, 90, 1e-2, 3
    rng = np.random.default_rng(3 + B)
    W = rng.standard_normal((n, n)) * 1.5 / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)) / np.sqrt(m), rng.standard_normal((k, n)) / np.sqrt(n)
    params = dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1)
    x = rng.standard_normal((T, B, m))
    net, node = _engine("li_tanh", n, B, dt, W, params)
    net.add_func_node("inp", m, "tanh"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    obs = net.run(x if B > 1 else x[:, 0], sampling_steps=S, verbose=False, enable_grad=True)
    out = torch.stack(obs["out"]).reshape(-1, B, k)
    out.square().sum().backward()
    g_eng = [node["weights"].grad, net.get_edge("inp", "rnn").weights.grad, net.get_edge("rnn", "out").weights.grad]
    g_ref = None
    for b in range(B):
        onode = orc.make_node("li_tanh", n, W, dt, params=params, train_params=["weights"])
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, requires_grad=True), w_out=torch.tensor(w_out, requires_grad=True), in_act="tanh")
        r = onet.run(torch.tensor(x[:, b]), sampling_steps=S, enable_grad=True)
        ref = torch.stack(r["out"])
        ref.square().sum().backward()
        assert rel_err(out[:, b].detach().cpu().numpy(), ref.detach().numpy()) < 1e-5
        gs = [p.grad.numpy().copy() for p in onet.parameters()]
        g_ref = gs if g_ref is None else [a + c for a, c in zip(g_ref, gs)]
    for ge, gr in zip(g_eng, g_ref):
        assert rel_err(ge.cpu().numpy(), gr) < 2e-5


@pytest.mark.parametrize("B", [2, 20])
def test_lif_s_ext_input_and_v_output(B):
    """lif_op with the input routed to s_ext (documentation/interfaces/train_test.py:37) and the membrane potential as output."""
    n, T, dt = 40, 400, 5e-3
    rng = np.random.default_rng(11 + B)
    W = rng.standard_normal((n, n))
    params = dict(eta=10.0, tau=rng.uniform(10, 20, n), tau_s=5.0, k=2.0)
    skw = dict(spike_threshold=10.0, spike_reset=-10.0)
    x = np.abs(rng.standard_normal((T, B, n))) * 2.0
    net, node = _engine("lif", n, B, dt, W, params, input_var="s_ext", output_var="v", train=("weights", "lif_op/tau_s"), skw=skw)
    obs = net.run(x, sampling_steps=2, verbose=False, enable_grad=True, record_vars=[("rnn", "lif_op/s", False)])
    out = torch.stack(obs["out"])
    assert out.shape == (T // 2, B, n)
    (out * 1e-3).square().sum().backward()
    gW_ref, gts_ref = 0.0, 0.0
    for b in range(B):
        onode = orc.make_node("lif", n, W, dt, params=params, train_params=["weights", "tau_s"], input_var="s_ext", output_var="v", **skw)
        onet = orc.OracleNet(onode)
        r = onet.run(torch.tensor(x[:, b]), sampling_steps=2, record_vars=[("s", False)], enable_grad=True)
        ref = torch.stack(r["out"])
        (ref * 1e-3).square().sum().backward()
        assert rel_err(out[:, b].detach().cpu().numpy(), ref.detach().numpy()) < 1e-4
        s_ref = torch.stack([v.detach() for v in r["vars"]["s"]]).numpy()
        assert rel_err(obs.to_numpy(("rnn", "lif_op/s"))[:, b], s_ref) < 1e-4
        gW_ref = gW_ref + onode.get("weights").grad.numpy()
        gts_ref = gts_ref + onode.get("tau_s").grad.numpy()
    assert rel_err(node["weights"].grad.cpu().numpy(), gW_ref) < 2e-3
    assert rel_err(node["lif_op/tau_s"].grad.cpu().numpy(), gts_ref) < 2e-3


def test_bare_node_scalar_input_and_degenerate_record_grids():
    import rectipy_b200 as rp
    n, dt = 30, 1e-2
    rng = np.random.default_rng(0)
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    params = dict(tau=1.5, k=1.0, eta=0.0)
    # scalar input broadcast to all neurons of a node without input edge (nodes.py:76-77: n_in from the arg shape)
    net, node = _engine("li_tanh", n, 1, dt, W, params, train=())
    x = rng.standard_normal((50, 1))
    obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=False)
    onode = orc.make_node("li_tanh", n, W, dt, params=params)
    r = orc.OracleNet(onode).run(torch.tensor(np.repeat(x, n, axis=1)), enable_grad=False)
    assert rel_err(obs.to_numpy("out"), torch.stack(r["out"]).numpy()) < 1e-5
    # cutoff >= T: nothing is recorded, the state still advances
    net2, node2 = _engine("li_tanh", n, 1, dt, W, params, train=())
    obs2 = net2.run(np.repeat(x, n, axis=1), sampling_steps=3, cutoff=60, verbose=False, enable_grad=False)
    assert len(obs2["out"]) == 0 and len(obs2["steps"]) == 0
    assert rel_err(node2.y.cpu().numpy(), node.y.cpu().numpy()) < 1e-6
    # sampling_steps > T: only step 0 is recorded (window of one sample)
    net3, _ = _engine("li_tanh", n, 1, dt, W, params, train=())
    obs3 = net3.run(np.repeat(x, n, axis=1), sampling_steps=100, verbose=False, enable_grad=False)
    assert obs3["steps"] == [0] and rel_err(obs3.to_numpy("out")[0], np.zeros(n) + 0.0) == 0.0
    # T = 0
    net4, node4 = _engine("li_tanh", n, 1, dt, W, params, train=())
    y_before = node4.y.clone()
    obs4 = net4.run(np.zeros((0, n)), verbose=False, enable_grad=False)
    assert len(obs4["out"]) == 0 and torch.equal(node4.y, y_before)
    with pytest.raises(RuntimeError):
        net4.run(np.zeros((5, n + 1)), verbose=False)


def test_softmax_output_with_windows_batched():
    """non-linear output node + sampling_steps > 1 + trials: activation per step, then the Observer window mean."""
    n, m, k, T, dt, S, B = 24, 2, 3, 61, 1e-2, 4, 3
    rng = np.random.default_rng(9)
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n))
    params = dict(tau=1.0, k=1.0, eta=0.0)
    x = rng.standard_normal((T, B, m))
    net, node = _engine("li_tanh", n, B, dt, W, params)
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in)
    net.add_func_node("out", k, "softmax"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    obs = net.run(x, sampling_steps=S, cutoff=2, verbose=False, enable_grad=True, record_vars=[("rnn", "li_op/v", True)])
    out = torch.stack(obs["out"])
    out[:, :, 0].sum().backward()
    for b in range(B):
        onode = orc.make_node("li_tanh", n, W, dt, params=params, train_params=["weights"])
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out, requires_grad=True), out_act="softmax")
        r = onet.run(torch.tensor(x[:, b]), sampling_steps=S, cutoff=2, record_vars=[("v", True)], enable_grad=True)
        ref = torch.stack(r["out"]).detach().numpy()
        assert rel_err(out[:, b].detach().cpu().numpy(), ref) < 1e-5
        assert rel_err(obs.to_numpy(("rnn", "li_op/v"))[:, b], torch.stack([v.detach() for v in r["vars"]["v"]]).numpy()) < 1e-5
    assert torch.isfinite(node["weights"].grad).all() and float(node["weights"].grad.abs().max()) > 0


def test_binary16_range_guard_reports_exploding_adjoint():
    """RP_PREC_3XF16 keeps one power-of-two scale per weight-gradient K chunk (2^11 headroom + 2^3 absorbed by the
    source operand).  An adjoint that grows ~9x per reverse step (diagonal W = 8, tau = 10, dt = 1, v = 0) leaves that range within
    one chunk: the call must fail loudly with a pointer to RP_PREC_3XTF32 -- and the tf32 format must handle the same
    problem (finite gradients, equal to the FFMA path)."""
    n, B, T, dt, k = 128, 128, 24, 1.0, 2
    rng = np.random.default_rng(0)
    W = 8.0 * np.eye(n)
    w_out = rng.standard_normal((k, n)) / np.sqrt(n)
    x = np.zeros((T, B, n), dtype=np.float32)
    tgt = torch.tensor(rng.standard_normal((T, B, k)).astype(np.float32), device="cuda")
    grads = {}
    for prec in ("3xf16", "3xtf32", "fp32"):
        net, node = _engine("li_tanh", n, B, dt, W, dict(tau=10.0, k=1.0, eta=0.0), prec=prec)
        net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
        obs = net.run(x, verbose=False, enable_grad=True)
        loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt)
        if prec == "3xf16":
            with pytest.raises(RuntimeError, match="RP_PREC_3XTF32"):
                loss.backward()
                torch.cuda.synchronize()
        else:
            loss.backward()
            grads[prec] = net.get_edge("rnn", "out").weights.grad.cpu().numpy()
            assert np.isfinite(grads[prec]).all() and np.isfinite(node["weights"].grad.cpu().numpy()).all()
    assert rel_err(grads["3xtf32"], grads["fp32"]) < 1e-5
    # the guard is sticky only until it is read: the same plan works again on a tame problem
    net, node = _engine("li_tanh", n, B, 1e-2, rng.standard_normal((n, n)) / np.sqrt(n), dict(tau=10.0, k=1.0, eta=0.0), prec="3xf16")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    obs = net.run(rng.standard_normal((T, B, n)).astype(np.float32), verbose=False, enable_grad=True)
    torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt).backward()
    assert np.isfinite(node["weights"].grad.cpu().numpy()).all()
