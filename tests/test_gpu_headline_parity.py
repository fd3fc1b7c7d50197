"""GPU parity of the HEADLINE execution path at the headline shapes, against the CPU oracle.

The default precision ("auto" -> RP_PREC_3XF16: tcgen05 split-3 contractions on binary16 hi/lo words, fused forward epilogue,
fused reverse kernels) at N = 4096 / 8192 neurons and >= 256 trials is what `bench.py` times; these tests pin exactly that chain
to the oracle (fp64 truth and the reference's own fp32 arithmetic) on sampled trials:

  (a) QIF N=4096, 256 trials, T=1000 steps from the spread initial state: identical per-neuron spike counts, spike times within
      one step, readout records <= 1e-4 relative.  T = 1000 is the "stated horizon" of DESIGN.md section 4 (the oracle's own fp32
      and fp64 runs still agree spike for spike there).
  (b) gradients: dL/dout is zero for all but the sampled trials, so the engine's dW / dW_out must equal the SUM of the
      oracle's per-trial gradients (no transitive comparison through the FFMA path).
  (c) LI-tanh N=4096, 256 trials: trajectories AND dW, dW_out <= 1e-5 relative vs the fp64 oracle (BASELINE bar).
  (d) one short N=8192 case (BASELINE configs[4] shape).

Reference lines: rectipy/nodes.py:166-170,382-392,468-481; rectipy/network.py:588-599,1123-1130.
The oracle runs cost ~8 s per 1000 steps and trial on 8 host cores; the sampled-trial counts keep the file at a few minutes.
"""
import numpy as np
import pytest
import torch

from golden_util import rel_err, orc

pytestmark = pytest.mark.gpu

DT = 1e-3
QIF_PATH = "neuron_model_templates.spiking_neurons.qif.qif"
LI_PATH = "neuron_model_templates.rate_neurons.leaky_integrator.tanh"


def _qif_problem(n, B, T, m=2, k=3, seed=99):
    """Recipe of SURVEY 8(d) M-QIF / bench.py: W = 2 randn/sqrt(N), Lorentzian eta, sinusoidal input + 8, spread initial phases."""
    rng = np.random.default_rng(seed)
    W = (2.0 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    w_in = rng.standard_normal((n, m)).astype(np.float32)
    w_out = (rng.standard_normal((k, n)) / np.sqrt(n)).astype(np.float32)
    etas = orc.lorentzian_etas(n).astype(np.float32)
    t = np.arange(T, dtype=np.float32) * DT
    amp = rng.uniform(5, 15, (1, B, 1)).astype(np.float32)
    phase = rng.uniform(0, 2 * np.pi, (1, B, m)).astype(np.float32)
    omega = np.asarray([3.0, 5.0], dtype=np.float32)[None, None, :]
    x = (amp * np.sin(2 * np.pi * omega * t[:, None, None] + phase) + 8.0).astype(np.float32)
    y0 = np.concatenate([rng.uniform(-50.0, 99.0, (B, n)), np.zeros((B, n))], axis=1).astype(np.float32)
    return W, w_in, w_out, etas, x, y0


def _qif_engine(n, B, W, w_in, w_out, etas, y0, train=True):
    import rectipy_b200 as rp
    from rectipy_b200 import _cabi
    net = rp.Network(DT, device="cuda:0", batch=B)        # default precision: auto -> 3xf16 on tensor-core shapes
    node = net.add_diffeq_node("qif", QIF_PATH, weights=W, source_var="s", target_var="s_in", input_var="I_ext", output_var="s",
                               spike_var="spike", reset_var="v", op="qif_op", node_vars={"eta": etas},
                               train_params=["weights"] if train else None)
    net.add_func_node("inp", w_in.shape[1], "identity"); net.add_edge("inp", "qif", weights=w_in)
    net.add_func_node("out", w_out.shape[0], "identity"); net.add_edge("qif", "out", weights=w_out, train="gd" if train else None)
    net.compile()
    node.reset(y0)
    return net, node, _cabi


def _assert_headline_plan(_cabi):
    from rectipy_b200 import engine
    precs = {p.key.precision for p in engine._PLANS.values()}
    assert precs == {_cabi.RP_PREC_3XF16}, f"the headline path (3xf16 tcgen05) did not run: plan precisions {precs}"


def _oracle_qif(n, W, w_in, w_out, etas, y0_b, dtype, train=False):
    node = orc.make_node("qif", n, W, DT, params=dict(eta=etas), dtype=dtype, y0=y0_b, train_params=["weights"] if train else None)
    return orc.OracleNet(node, w_in=torch.tensor(w_in, dtype=dtype), w_out=torch.tensor(w_out, dtype=dtype, requires_grad=train))


def test_headline_qif_rasters_and_records_long_horizon():
    """(a) N=4096, 256 trials, T=1000 on the default tcgen05 path vs the fp32 AND fp64 oracle for 4 sampled trials."""
    from rectipy_b200 import engine
    engine.clear_plans()
    n, B, T = 4096, 256, 1000
    W, w_in, w_out, etas, x, y0 = _qif_problem(n, B, T)
    net, node, _cabi = _qif_engine(n, B, W, w_in, w_out, etas, y0, train=False)
    obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=False, record_vars=[("qif", "v", False)])
    _assert_headline_plan(_cabi)
    out = torch.stack(obs["out"])                      # [T, B, k]
    v = torch.stack(obs[("qif", "v")])                 # [T, B, n]
    assert torch.isfinite(out).all()
    report = {}
    for b in (0, 85, 170, 255):
        r_eng = (v[:, b, :] >= 100.0).cpu().numpy()
        o_eng = out[:, b, :].cpu().numpy()
        for dtype in (torch.float64, torch.float32):
            onet = _oracle_qif(n, W, w_in, w_out, etas, y0[b], dtype)
            r = onet.run(torch.tensor(x[:, b, :], dtype=dtype), sampling_steps=1, record_vars=[("v", False)], enable_grad=False)
            r_ref = orc.spike_raster(torch.stack(r["vars"]["v"]).numpy(), 100.0)
            o_ref = torch.stack(r["out"]).numpy()
            cmp_ = orc.compare_spikes(r_ref, r_eng)
            err = rel_err(o_eng, o_ref)
            report[(b, str(dtype))] = (cmp_, err)
            assert cmp_["total_ref"] > 2000, cmp_
            assert cmp_["neurons_count_mismatch"] == 0 and cmp_["unmatched"] == 0 and cmp_["max_shift"] <= 1, (b, dtype, cmp_)
            assert err <= 1e-4, (b, dtype, err)
    print("headline rasters/records:", {k_: (c["total_ref"], c["max_shift"], f"{e:.2e}") for k_, (c, e) in report.items()})
    engine.clear_plans()


def _masked_gradient_check(n, B, T, trials, seed, bar):
    """(b)/(d): engine gradients with dL/dout masked to `trials` vs the summed per-trial fp64 oracle gradients."""
    from rectipy_b200 import engine
    engine.clear_plans()
    W, w_in, w_out, etas, x, y0 = _qif_problem(n, B, T, seed=seed)
    rng = np.random.default_rng(seed + 1)
    g = np.zeros((T, B, w_out.shape[0]), dtype=np.float32)
    for b in trials:
        g[:, b, :] = rng.standard_normal((T, w_out.shape[0]))
    net, node, _cabi = _qif_engine(n, B, W, w_in, w_out, etas, y0, train=True)
    obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=True, record_vars=[("qif", "v", False)])
    _assert_headline_plan(_cabi)
    out = torch.stack(obs["out"])
    (out * torch.tensor(g, device="cuda:0")).sum().backward()
    gW = node["weights"].grad.cpu().numpy()
    gWo = net.get_edge("qif", "out").weights.grad.cpu().numpy()
    v = torch.stack(obs[("qif", "v")])
    gW_ref, gWo_ref = np.zeros((n, n)), np.zeros(w_out.shape)
    for b in trials:
        onet = _oracle_qif(n, W, w_in, w_out, etas, y0[b], torch.float64, train=True)
        r = onet.run(torch.tensor(x[:, b, :], dtype=torch.float64), sampling_steps=1, record_vars=[("v", False)], enable_grad=True)
        pred = torch.stack(r["out"])
        (pred * torch.tensor(g[:, b, :], dtype=torch.float64)).sum().backward()
        gW_ref += onet.node.get("weights").grad.numpy()
        gWo_ref += onet.w_out.grad.numpy()
        r_ref = orc.spike_raster(torch.stack([q.detach() for q in r["vars"]["v"]]).numpy(), 100.0)
        cmp_ = orc.compare_spikes(r_ref, (v[:, b, :] >= 100.0).cpu().numpy())
        assert cmp_["total_ref"] > 0 and cmp_["neurons_count_mismatch"] == 0 and cmp_["max_shift"] <= 1, (b, cmp_)
        assert rel_err(out[:, b, :].detach().cpu().numpy(), pred.detach().numpy()) <= 1e-4
    errs = dict(dW=rel_err(gW, gW_ref), dW_out=rel_err(gWo, gWo_ref))
    print(f"headline gradients N={n} B={B} T={T} trials={trials}: realised rel err {errs} (bar {bar:.0e})")
    assert np.abs(gW_ref).max() > 0 and np.abs(gWo_ref).max() > 0
    assert all(e <= bar for e in errs.values()), errs
    engine.clear_plans()


def test_headline_qif_gradients_equal_summed_oracle_gradients():
    """(b) N=4096, 256 trials, T=300: realised error printed; bar 1e-3 (spiking gradients: the surrogate 1/(1+slope|v-theta|)^2
    and the threshold gate amplify fp32 rounding of v; the reference's own fp32 run sits at ~1e-4 of its fp64 run)."""
    _masked_gradient_check(4096, 256, 300, (3, 130, 254), seed=17, bar=1e-3)


def test_headline_qif_n8192_short():
    """(d) BASELINE configs[4] shape (N=8192), 128 trials, 120 steps: rasters, records and masked gradients vs the fp64 oracle."""
    _masked_gradient_check(8192, 128, 120, (5, 127), seed=23, bar=1e-3)


def test_headline_rate_network_trajectories_and_gradients_1e5():
    """(c) LI-tanh N=4096, 256 trials on the default tcgen05 path: records, dW and dW_out <= 1e-5 relative (max-norm) vs the
    fp64 oracle on sampled trials (dL/dout masked to them)."""
    import rectipy_b200 as rp
    from rectipy_b200 import engine, _cabi
    engine.clear_plans()
    n, B, T, dt, m, k = 4096, 256, 200, 1e-2, 2, 2
    rng = np.random.default_rng(31)
    W = (1.5 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    w_in = rng.standard_normal((n, m)).astype(np.float32)
    w_out = (rng.standard_normal((k, n)) / np.sqrt(n)).astype(np.float32)
    tau = rng.uniform(1.0, 2.0, n).astype(np.float32)
    t = np.arange(T, dtype=np.float32) * dt
    x = (1.5 * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m)))).astype(np.float32)
    trials = (0, 101, 255)
    g = np.zeros((T, B, k), dtype=np.float32)
    for b in trials:
        g[:, b, :] = rng.standard_normal((T, k))
    net = rp.Network(dt, device="cuda:0", batch=B)
    node = net.add_diffeq_node("rnn", LI_PATH, weights=W, source_var="tanh_op/r", target_var="li_op/r_in", input_var="li_op/I_ext",
                               output_var="li_op/v", node_vars={"li_op/tau": tau, "li_op/k": 1.2, "li_op/eta": 0.1},
                               train_params=["weights"])
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in)
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=True)
    assert {p.key.precision for p in engine._PLANS.values()} == {_cabi.RP_PREC_3XF16}
    out = torch.stack(obs["out"])
    (out * torch.tensor(g, device="cuda:0")).sum().backward()
    gW, gWo = node["weights"].grad.cpu().numpy(), net.get_edge("rnn", "out").weights.grad.cpu().numpy()
    gW_ref, gWo_ref = np.zeros((n, n)), np.zeros((k, n))
    errs = {}
    for b in trials:
        onode = orc.make_node("li_tanh", n, W, dt, params=dict(tau=tau, k=1.2, eta=0.1), dtype=torch.float64, train_params=["weights"])
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, dtype=torch.float64), w_out=torch.tensor(w_out, dtype=torch.float64, requires_grad=True))
        r = onet.run(torch.tensor(x[:, b, :], dtype=torch.float64), sampling_steps=1, enable_grad=True)
        pred = torch.stack(r["out"])
        (pred * torch.tensor(g[:, b, :], dtype=torch.float64)).sum().backward()
        gW_ref += onode.get("weights").grad.numpy()
        gWo_ref += onet.w_out.grad.numpy()
        errs[f"out[{b}]"] = rel_err(out[:, b, :].detach().cpu().numpy(), pred.detach().numpy())
    errs["dW"], errs["dW_out"] = rel_err(gW, gW_ref), rel_err(gWo, gWo_ref)
    print("headline rate network, realised rel err vs fp64 oracle:", {k_: f"{e:.2e}" for k_, e in errs.items()})
    assert all(e <= 1e-5 for e in errs.values()), errs
    engine.clear_plans()


def test_persistent_forward_kernel_is_bit_identical_to_per_step_launches(monkeypatch):
    """Round 2: the batched forward runs as ONE persistent cooperative launch per horizon (k_gemm_fwd_persist: per-trial-group
    dependency counters instead of a kernel boundary per step, double-buffered source operand).  RP_NO_FWD_PERSIST=1 forces one
    launch per step.  Same kernels' arithmetic, same operand scales -> records, final state and gradients must be bit-identical,
    with and without checkpoints, over a horizon long enough for every barrier phase to wrap many times."""
    from rectipy_b200 import engine
    n, B, T = 1024, 512, 300
    W, w_in, w_out, etas, x, y0 = _qif_problem(n, B, T, seed=5)
    g = torch.tensor(np.random.default_rng(6).standard_normal((T, B, w_out.shape[0])).astype(np.float32), device="cuda:0")
    res = {}
    for mode in ("persistent", "persistent_1cta", "per_step"):       # default: the CTA-pair kernel (k_gemm_fwd_persist_cg2)
        monkeypatch.delenv("RP_NO_FWD_PERSIST", raising=False)
        monkeypatch.delenv("RP_NO_FWD_CG2", raising=False)
        if mode == "per_step":
            monkeypatch.setenv("RP_NO_FWD_PERSIST", "1")
        elif mode == "persistent_1cta":
            monkeypatch.setenv("RP_NO_FWD_CG2", "1")
        engine.clear_plans()
        net, node, _cabi = _qif_engine(n, B, W, w_in, w_out, etas, y0, train=True)
        obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=True)
        launches_fwd = engine.total_launches()
        out = torch.stack(obs["out"])
        (out * g).sum().backward()
        rec = dict(out=out.detach().clone(), y=node.y.detach().clone(), gW=node["weights"].grad.clone(),
                   gWo=net.get_edge("qif", "out").weights.grad.clone(), launches_fwd=launches_fwd)
        node.reset(y0)
        obs2 = net.run(x, sampling_steps=7, cutoff=11, verbose=False, enable_grad=False)      # no checkpoints, windowed records
        rec["out_nograd"] = torch.stack(obs2["out"]).clone()
        rec["y_nograd"] = node.y.detach().clone()
        res[mode] = rec
    monkeypatch.delenv("RP_NO_FWD_PERSIST", raising=False)
    monkeypatch.delenv("RP_NO_FWD_CG2", raising=False)
    engine.clear_plans()
    assert res["persistent"]["launches_fwd"] < 20 < res["per_step"]["launches_fwd"], (res["persistent"]["launches_fwd"], res["per_step"]["launches_fwd"])
    assert res["persistent_1cta"]["launches_fwd"] < 20
    assert float(res["per_step"]["out"].abs().max()) > 0 and float(res["per_step"]["gW"].abs().max()) > 0
    for key in ("out", "y", "gW", "gWo", "out_nograd", "y_nograd"):
        assert torch.equal(res["persistent"][key], res["per_step"][key]), key
        assert torch.equal(res["persistent_1cta"][key], res["per_step"][key]), key


@pytest.mark.parametrize("prec", ["3xtf32", "3xf16"])
def test_per_step_tensor_core_forward_is_bit_reproducible(prec):
    """Rate networks with a readout run one tensor-core launch per step.  Round 2 found a write-after-read hazard there: the epilogue of a
    tile that had finished its K loop wrote src_{t+1} into the operand buffer other CTAs of the launch were still loading src_t from -- on a
    box whose CTAs drift apart (power throttling) three runs of the same forward gave three different results, 1e-4 apart in the summed
    output (tools/exp_determinism_v1.py).  The source operand is now double buffered by step parity on every tensor-core path; the
    forward must be bit-reproducible and within rounding of the fp32 path."""
    import rectipy_b200 as rp
    n, B, T, dt = 4096, 1024, 25, 1e-2
    rng = np.random.default_rng(5)
    W = (1.5 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    w_out = (rng.standard_normal((2, n)) / np.sqrt(n)).astype(np.float32)
    x = torch.tensor(rng.standard_normal((T, B, n)).astype(np.float32), device="cuda")

    def run(precision):
        net = rp.Network(dt, device="cuda:0", batch=B, precision=precision)
        net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=W, source_var="tanh_op/r",
                            target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v", node_vars={"li_op/tau": 0.5, "li_op/k": 1.3})
        net.add_func_node("out", 2, "identity"); net.add_edge("rnn", "out", weights=w_out)
        return torch.stack(net.run(x, verbose=False)["out"])
    outs = [run(prec) for _ in range(4)]
    for o in outs[1:]:
        assert torch.equal(o, outs[0])
    ref = run("fp32")
    err = float((outs[0] - ref).abs().max() / ref.abs().max())
    print(prec, "vs fp32 path:", err)
    assert err <= 1e-5
