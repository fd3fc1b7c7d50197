"""GPU parity: the CUDA engine, driven through the drop-in API and the C ABI, against the golden fixtures minted from
the unmodified reference and against the CPU oracle on seeded inputs.

Bars (BASELINE.json north_star):
  * rate-neuron trajectories and gradients: <= 1e-5 relative error (max-norm) vs the fp64 reference run;
  * spiking networks: identical spike counts, spike times within one step; continuous quantities as close to the fp64
    truth as the reference's own fp32 run (x10 slack), because threshold dynamics amplify fp32 rounding.
"""
import numpy as np
import pytest
import torch

from golden_util import Case, RUN_CASES, engine_run_case, engine_net, oracle_net, rel_err, orc

pytestmark = pytest.mark.gpu

RATE_TOL = 1e-5


def _bar(case, key, rate):
    """Allowed relative error vs fp64 truth for result `key`."""
    ref32 = rel_err(case.ref("float32", key), case.ref("float64", key))
    if rate:
        return max(RATE_TOL, 3.0 * ref32)
    return max(2e-5, 10.0 * ref32)


@pytest.mark.parametrize("name", RUN_CASES)
def test_engine_matches_reference_fixture(name):
    case = Case(name)
    res = engine_run_case(case, precision="fp32")
    rate = case.meta["model"] in ("li_tanh", "li_sigmoid")
    assert np.array_equal(res["steps"], case.ref("float64", "steps"))
    report = {}
    for key, val in res.items():
        if key == "steps":
            continue
        ref = case.ref("float64", key)
        assert val.shape == ref.shape, (key, val.shape, ref.shape)
        err, bar = rel_err(val, ref), _bar(case, key, rate)
        report[key] = (err, bar)
    print(name, {k: f"{e:.2e}/{b:.1e}" for k, (e, b) in report.items()})
    bad = {k: v for k, v in report.items() if not v[0] <= v[1]}
    assert not bad, bad


@pytest.mark.parametrize("model,n,B,m,k", [("li_tanh", 37, 3, 2, 2), ("qif", 50, 5, 3, 1), ("qif_sfa", 64, 20, 2, 3),
                                            ("lif", 33, 2, 1, 2), ("li_sigmoid", 130, 17, 4, 8)])
def test_batched_trials_match_per_trial_oracle(model, n, B, m, k):
    """The trial axis is an engine extension: every trial must equal the unbatched reference path run on its own."""
    import rectipy_b200 as rp
    from golden_util import TEMPLATE_PATH
    rng = np.random.default_rng(1000 * n + 10 * B + len(model))      # deterministic (str hashes are salted per process)
    dt, T, S = (1e-2, 120, 3) if model.startswith("li_") else (1e-3, 400, 4)
    W = rng.standard_normal((n, n)) * (1.5 if model.startswith("li_") else 2.0) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    params = {"li_tanh": dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1),
              "li_sigmoid": dict(tau=2.0, k=rng.uniform(0.5, 1.5, n), eta=0.0, r_max=1.5, s=2.0, v0=0.1),
              "qif": dict(eta=orc.lorentzian_etas(n), k=1.5, tau_s=0.7),
              "qif_sfa": dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.4, tau_x=1.2),
              "lif": dict(eta=10.0, tau=rng.uniform(10, 20, n), tau_s=5.0, k=2.0)}[model]
    spike_kwargs = dict(spike_threshold=10.0, spike_reset=-10.0) if model == "lif" else {}
    amp, off = (1.5, 0.0) if model.startswith("li_") else ((40.0, 0.0) if model == "lif" else (10.0, 14.0))
    t = np.arange(T) * dt
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] * (50 if model == "lif" else 1) +
                     rng.uniform(0, 6.28, (1, B, m))) + off
    targets = rng.standard_normal((len(range(0, T, S)), B, k))
    train = ["weights", "eta", "tau"]

    # engine, all trials at once
    path, op, svar, tvar = TEMPLATE_PATH[model]
    net = rp.Network(dt, device="cuda:0", batch=B, precision="fp32")
    sig = {"r_max", "s", "v0"}
    node_vars = {(f"sigmoid_op/{p}" if p in sig and model == "li_sigmoid" else f"{op}/{p}"): v for p, v in params.items()}
    kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext", node_vars=node_vars,
              train_params=["weights", f"{op}/eta", f"{op}/tau"])
    if model in orc.SPIKING:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **spike_kwargs)
    else:
        kw.update(output_var=f"{op}/v")
    node = net.add_diffeq_node("rnn", path, **kw)
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True)
    out = torch.stack(obs["out"])
    assert out.shape == (len(range(0, T, S)), B, k)
    loss = torch.nn.functional.mse_loss(out, torch.tensor(targets, dtype=torch.float32, device="cuda:0"))
    loss.backward()

    # oracle, one trial at a time in fp64; gradients of the mean loss add over trials
    g_ref = None
    out_ref = np.zeros(out.shape)
    for b in range(B):
        onode = orc.make_node(model, n, W, dt, params=params, dtype=torch.float64, train_params=train, **spike_kwargs)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, requires_grad=True), w_out=torch.tensor(w_out, requires_grad=True))
        r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=True)
        pred = torch.stack(r["out"])
        out_ref[:, b, :] = pred.detach().numpy()
        lb = torch.nn.functional.mse_loss(pred, torch.tensor(targets[:, b, :]), reduction="sum") / out.numel()
        lb.backward()
        gs = [p.grad.numpy().copy() for p in onet.parameters()]
        g_ref = gs if g_ref is None else [a + c for a, c in zip(g_ref, gs)]
    rate = model.startswith("li_")
    tol = RATE_TOL if rate else 1e-4
    assert rel_err(out.detach().cpu().numpy(), out_ref) <= tol
    g_eng = [node["weights"].grad, node[f"{op}/eta"].grad, node[f"{op}/tau"].grad,
             net.get_edge("inp", "rnn").weights.grad, net.get_edge("rnn", "out").weights.grad]
    names = ["weights", "eta", "tau", "w_in", "w_out"]
    errs = {nm: rel_err(ge.detach().cpu().numpy().reshape(gr.shape), gr) for nm, ge, gr in zip(names, g_eng, g_ref)}
    print(model, errs)
    assert all(e <= (RATE_TOL if rate else 2e-3) for e in errs.values()), errs


def test_spike_raster_parity_config1_style():
    """QIF-SFA forward run in the regime of documentation/qif_example.py (scaled): spike counts identical and spike
    times within one step of the fp32 CPU oracle over the stated horizon of 4000 steps."""
    import rectipy_b200 as rp
    n, T, dt = 256, 4000, 1e-3
    rng = np.random.default_rng(5)
    np.random.seed(5)
    W = rp.random_connectivity(n, n, 0.2, normalize=True)
    etas = orc.lorentzian_etas(n)
    inp = np.zeros((T, 1)); inp[500:3500, 0] = 6.0
    w_in = np.ones((n, 1))
    params = dict(eta=etas, k=15.0, alpha=0.3, tau_x=2.0)
    net = rp.Network(dt, device="cuda:0", precision="fp32")
    net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif_sfa", weights=W, source_var="s",
                        target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v",
                        op="qif_sfa_op", node_vars={"eta": etas, "k": 15.0, "alpha": 0.3, "tau_x": 2.0})
    net.add_func_node("inp", 1, "identity"); net.add_edge("inp", "qif", weights=w_in)
    obs = net.run(inp, sampling_steps=1, verbose=False, enable_grad=False, record_output=False,
                  record_vars=[("qif", "v", False)])
    v_eng = obs.to_numpy(("qif", "v"))
    onode = orc.make_node("qif_sfa", n, W, dt, params=params, dtype=torch.float32)
    onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, dtype=torch.float32))
    r = onet.run(torch.tensor(inp, dtype=torch.float32), sampling_steps=1, record_vars=[("v", False)], enable_grad=False)
    v_ref = torch.stack(r["vars"]["v"]).numpy()
    cmp_ = orc.compare_spikes(orc.spike_raster(v_ref, 100.0), orc.spike_raster(v_eng, 100.0))
    print(cmp_)
    assert cmp_["total_ref"] > 200
    assert cmp_["neurons_count_mismatch"] == 0 and cmp_["unmatched"] == 0 and cmp_["max_shift"] <= 1


@pytest.mark.parametrize("P,Q,K", [(128, 128, 32), (256, 512, 1024), (37, 5, 19), (130, 70, 257), (1000, 1, 1000), (64, 16, 64)])
def test_fp32_contraction_kernels(P, Q, K):
    from rectipy_b200 import engine
    g = torch.Generator(device="cuda").manual_seed(P * 7 + Q)
    A = torch.randn(P, K, device="cuda", generator=g)
    B = torch.randn(Q, K, device="cuda", generator=g)
    C = engine.gemm_tn(A, B)
    ref = (B.double() @ A.double().T)
    assert rel_err(C.cpu().numpy(), ref.cpu().numpy()) < 2e-6


TC_PRECS = ("3xtf32", "3xf16")      # tensor-core operand formats: two tf32 words | two binary16 words + power-of-two scales


@pytest.mark.parametrize("prec", TC_PRECS)
@pytest.mark.parametrize("P,Q,K", [(128, 128, 32), (128, 256, 64), (256, 128, 96), (512, 1024, 2048), (384, 384, 500)])
def test_tcgen05_split3_contraction(P, Q, K, prec):
    """tcgen05 split-3 kernel (tf32 and binary16 operands) vs fp64 matmul: error must be at fp32 level (a single-pass
    TF32 / FP16 product would be ~1e-3)."""
    from rectipy_b200 import engine, _cabi
    code = _cabi.RP_PREC_3XTF32 if prec == "3xtf32" else _cabi.RP_PREC_3XF16
    g = torch.Generator(device="cuda").manual_seed(P + Q + K)
    A = torch.randn(P, K, device="cuda", generator=g)
    B = torch.randn(Q, K, device="cuda", generator=g)
    C = engine.gemm_tn(A, B, precision=code)
    torch.cuda.synchronize()
    ref = (B.double() @ A.double().T)
    err = rel_err(C.cpu().numpy(), ref.cpu().numpy())
    fp32_err = rel_err((B @ A.T).cpu().numpy(), ref.cpu().numpy())
    print(f"{prec} rel err {err:.2e}  (torch fp32 matmul {fp32_err:.2e})")
    assert err < max(3e-6, 2.0 * fp32_err)     # chunked fp32 re-accumulation keeps the tcgen05 path at FFMA accuracy
    C2 = engine.gemm_tn(A, B, precision=code, out=C.clone(), accumulate=True)
    assert rel_err(C2.cpu().numpy(), 2 * ref.cpu().numpy()) < 5e-6


@pytest.mark.parametrize("case", ["tiny", "huge", "positive", "wide_range", "sparse_spikes"])
def test_binary16_split_operand_scaling(case):
    """RP_PREC_3XF16 carries the split in binary16 words; its 5-bit exponent is handled by exact power-of-two operand
    scales.  Operands far outside the binary16 range, one-signed sums (accumulator truncation) and operands spanning many
    orders of magnitude must all come out at fp32 accuracy relative to the result's norm, like the tf32 split."""
    from rectipy_b200 import engine, _cabi
    P, Q, K = 256, 256, 1024
    g = torch.Generator(device="cuda").manual_seed(7)
    A = torch.randn(P, K, device="cuda", generator=g)
    B = torch.randn(Q, K, device="cuda", generator=g)
    if case == "tiny":
        A, B = A * 1e-20, B * 3e-9
    elif case == "huge":
        A, B = A * 1e15, B * 7e11
    elif case == "positive":
        A, B = A.abs() + 0.5, B.abs() + 0.25
    elif case == "wide_range":
        A = A * torch.exp(6.0 * torch.randn(P, K, device="cuda", generator=g))
        B = B * torch.exp(6.0 * torch.randn(Q, K, device="cuda", generator=g))
    else:
        B = (torch.rand(Q, K, device="cuda", generator=g) < 0.02).float() + 0.3 * torch.rand(Q, K, device="cuda", generator=g) ** 8
    ref = (B.double() @ A.double().T).cpu().numpy()
    errs = {}
    for prec, code in (("fp32", _cabi.RP_PREC_FP32), ("3xtf32", _cabi.RP_PREC_3XTF32), ("3xf16", _cabi.RP_PREC_3XF16)):
        C = engine.gemm_tn(A, B, precision=code)
        torch.cuda.synchronize()
        assert torch.isfinite(C).all(), (case, prec)
        errs[prec] = rel_err(C.cpu().numpy(), ref)
    print(case, errs)
    assert errs["3xf16"] < max(3e-6, 2.0 * errs["fp32"], 2.0 * errs["3xtf32"]), errs


@pytest.mark.parametrize("tc_prec", TC_PRECS)
@pytest.mark.parametrize("model", ["li_tanh", "qif"])
def test_tensor_core_path_matches_fp32_path(model, tc_prec):
    """N=256, B=256: the tensor-core engine path (both operand formats) vs the FFMA path vs the oracle (sample of trials)."""
    import rectipy_b200 as rp
    from golden_util import TEMPLATE_PATH
    n, B, m, k = 256, 256, 2, 3
    rng = np.random.default_rng(11)
    dt, T, S = (1e-2, 60, 2) if model == "li_tanh" else (1e-3, 300, 2)
    W = rng.standard_normal((n, n)) * (1.5 if model == "li_tanh" else 2.0) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    params = dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1) if model == "li_tanh" else dict(eta=orc.lorentzian_etas(n), k=1.5, tau_s=0.7)
    t = np.arange(T) * dt
    amp, off = (1.5, 0.0) if model == "li_tanh" else (10.0, 14.0)
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + off
    targets = torch.tensor(rng.standard_normal((len(range(0, T, S)), B, k)), dtype=torch.float32, device="cuda")
    path, op, svar, tvar = TEMPLATE_PATH[model]
    y_init = np.concatenate([rng.uniform(-50.0, 99.0, (B, n)), np.zeros((B, n))], axis=1).astype(np.float32)
    results = {}
    for prec in ("fp32", tc_prec):
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext",
                  node_vars={f"{op}/{p}": v for p, v in params.items()}, train_params=["weights", f"{op}/eta"])
        if model == "qif":
            kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s")
        else:
            kw.update(output_var=f"{op}/v")
        node = net.add_diffeq_node("rnn", path, **kw)
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
        net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
        if model == "qif":
            node.reset(y_init)          # membrane potentials spread over the cycle -> threshold crossings within the horizon
        obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True)
        out = torch.stack(obs["out"])
        torch.nn.functional.mse_loss(out, targets).backward()
        results[prec] = dict(out=out.detach().cpu().numpy(), gW=node["weights"].grad.cpu().numpy(),
                             geta=node[f"{op}/eta"].grad.cpu().numpy(),
                             gin=net.get_edge("inp", "rnn").weights.grad.cpu().numpy(),
                             gout=net.get_edge("rnn", "out").weights.grad.cpu().numpy(), y=node.y.detach().cpu().numpy())
    errs = {key: rel_err(results[tc_prec][key], results["fp32"][key]) for key in results["fp32"]}
    print(model, tc_prec, errs)
    assert np.abs(results["fp32"]["out"]).max() > 0 and np.abs(results["fp32"]["gW"]).max() > 0      # spikes happened, gradient is live
    tol = 1e-5 if model == "li_tanh" else 1e-3
    assert all(e <= tol for e in errs.values()), errs
    # and against the oracle for two trials
    for b in (0, B - 1):
        onode = orc.make_node(model, n, W, dt, params=params, dtype=torch.float64, y0=y_init[b] if model == "qif" else None)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out))
        r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=False)
        ref = torch.stack(r["out"]).numpy()
        assert rel_err(results[tc_prec]["out"][:, b, :], ref) <= (1e-5 if model == "li_tanh" else 1e-4)


@pytest.mark.parametrize("model", ["qif_sfa", "lif", "li_sigmoid"])
def test_binary16_path_remaining_templates(model):
    """The default tensor-core format (binary16 split operands) on the templates the headline tests do not touch: records of
    two trials vs the fp64 oracle, every gradient (weights, template parameters, both edges) vs the FFMA path."""
    import rectipy_b200 as rp
    from golden_util import TEMPLATE_PATH
    n, B, m, k = 128, 128, 2, 3
    rng = np.random.default_rng(21)
    rate = model.startswith("li_")
    dt, T, S = (1e-2, 90, 3) if rate else ((5e-3, 600, 2) if model == "lif" else (1e-3, 300, 2))     # lif: tau ~ 15 -> first spikes after ~200 steps
    W = rng.standard_normal((n, n)) * (1.5 if rate else 2.0) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    params = {"li_sigmoid": dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1),
              "qif_sfa": dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.4, tau_x=1.2),
              "lif": dict(eta=10.0, tau=rng.uniform(10, 20, n), tau_s=5.0, k=2.0)}[model]
    skw = dict(spike_threshold=10.0, spike_reset=-10.0) if model == "lif" else {}
    t = np.arange(T) * dt
    amp, off = (1.5, 0.0) if rate else ((40.0, 0.0) if model == "lif" else (10.0, 14.0))
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + off
    targets = torch.tensor(rng.standard_normal((len(range(0, T, S)), B, k)), dtype=torch.float32, device="cuda")
    path, op, svar, tvar = TEMPLATE_PATH[model]
    res = {}
    for prec in ("fp32", "3xf16"):
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext",
                  node_vars={f"{op}/{p}": v for p, v in params.items()}, train_params=["weights", f"{op}/eta", f"{op}/tau"])
        if rate:
            kw.update(output_var=f"{op}/v")
        else:
            kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **skw)
        node = net.add_diffeq_node("rnn", path, **kw)
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
        net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
        obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True)
        out = torch.stack(obs["out"])
        torch.nn.functional.mse_loss(out, targets).backward()
        res[prec] = dict(out=out.detach().cpu().numpy(), gW=node["weights"].grad.cpu().numpy(), geta=node[f"{op}/eta"].grad.cpu().numpy(),
                         gtau=node[f"{op}/tau"].grad.cpu().numpy(), gin=net.get_edge("inp", "rnn").weights.grad.cpu().numpy(),
                         gout=net.get_edge("rnn", "out").weights.grad.cpu().numpy())
    errs = {key: rel_err(res["3xf16"][key], res["fp32"][key]) for key in res["fp32"]}
    print(model, errs)
    assert np.abs(res["fp32"]["gW"]).max() > 0
    tol = 1e-5 if rate else 1e-3
    assert all(e <= tol for e in errs.values()), errs
    for b in (0, B - 1):
        onode = orc.make_node(model, n, W, dt, params=params, dtype=torch.float64, **skw)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out))
        ref = torch.stack(onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=False)["out"]).numpy()
        assert rel_err(res["3xf16"]["out"][:, b, :], ref) <= (1e-5 if rate else 1e-4)


@pytest.mark.parametrize("model,n,B", [("qif_sfa", 1000, 1), ("li_tanh", 203, 2), ("lif", 64, 4), ("qif", 1536, 1),
                                       ("qif", 1000, 32), ("li_tanh", 203, 7), ("qif_sfa", 500, 20), ("lif", 130, 64)])
def test_persistent_kernels_match_per_step_path(model, n, B, monkeypatch):
    """Few-trial shapes run as ONE cooperative persistent launch per pass (rp_persistent.cuh); RP_NO_PERSISTENT=1 forces
    the per-step launch sequence.  Both must agree (same fp32 arithmetic up to summation order).  Round 2: trial counts above 4
    run on the two-dimensional (row block x trial block) grid, with register-tiled products when a CTA holds more than 4 trials."""
    import rectipy_b200 as rp
    from rectipy_b200 import engine
    from golden_util import TEMPLATE_PATH
    rng = np.random.default_rng(n + B)
    rate = model.startswith("li_")
    dt, T, S, cutoff = (1e-2, 150, 3, 4) if rate else (1e-3, 400, 2, 0)
    m, k = 2, 3
    W = rng.standard_normal((n, n)) * (1.5 if rate else 2.0) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    path, op, svar, tvar = TEMPLATE_PATH[model]
    params = {"li_tanh": dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1), "qif": dict(eta=orc.lorentzian_etas(n), k=1.5),
              "qif_sfa": dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.4, tau_x=1.2),
              "lif": dict(eta=10.0, tau=rng.uniform(10, 20, n), tau_s=5.0, k=2.0)}[model]
    spike_kwargs = dict(spike_threshold=10.0, spike_reset=-10.0) if model == "lif" else {}
    t = np.arange(T) * dt
    amp, off = (1.5, 0.0) if rate else ((40.0, 0.0) if model == "lif" else (10.0, 14.0))
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] * (50 if model == "lif" else 1)
                     + rng.uniform(0, 6.28, (1, B, m))) + off
    n_rec = len([s for s in range(T) if s >= cutoff and s % S == 0])
    targets = torch.tensor(rng.standard_normal((n_rec, B, k)), dtype=torch.float32, device="cuda")
    res = {}
    for mode in ("persistent", "per_step"):
        if mode == "per_step":
            monkeypatch.setenv("RP_NO_PERSISTENT", "1")
        else:
            monkeypatch.delenv("RP_NO_PERSISTENT", raising=False)
        engine.clear_plans()
        net = rp.Network(dt, device="cuda:0", batch=B, precision="fp32")
        kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext",
                  node_vars={f"{op}/{p}": v for p, v in params.items()}, train_params=["weights", f"{op}/eta", f"{op}/tau", f"{op}/k"])
        if not rate:
            kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **spike_kwargs)
        else:
            kw.update(output_var=f"{op}/v")
        node = net.add_diffeq_node("rnn", path, **kw)
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
        net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
        var = "v"
        rec = [("rnn", f"{op}/v", False)] + ([] if rate else [("rnn", f"{op}/s", True)])
        obs = net.run(x, sampling_steps=S, cutoff=cutoff, verbose=False, enable_grad=True, record_vars=rec, truncate_steps=170)
        out = torch.stack(obs["out"])
        torch.nn.functional.mse_loss(out.reshape(n_rec, B, k), targets).backward()
        launches = engine.total_launches()
        res[mode] = dict(out=out.detach().cpu().numpy(), var=obs.to_numpy(("rnn", f"{op}/{var}")), y=node.y.detach().cpu().numpy(),
                         var2=obs.to_numpy(("rnn", f"{op}/v")) if rate else obs.to_numpy(("rnn", f"{op}/s")),
                         gW=node["weights"].grad.cpu().numpy(), geta=node[f"{op}/eta"].grad.cpu().numpy(),
                         gtau=node[f"{op}/tau"].grad.cpu().numpy(), gk=node[f"{op}/k"].grad.cpu().numpy(),
                         gin=net.get_edge("inp", "rnn").weights.grad.cpu().numpy(),
                         gout=net.get_edge("rnn", "out").weights.grad.cpu().numpy(), launches=launches)
    monkeypatch.delenv("RP_NO_PERSISTENT", raising=False)
    engine.clear_plans()
    assert res["persistent"]["launches"] < 12 < res["per_step"]["launches"]
    errs = {key: rel_err(res["persistent"][key], res["per_step"][key]) for key in res["persistent"] if key != "launches"}
    print(model, n, B, errs)
    tol = 2e-5 if rate else 2e-3
    assert all(e <= tol for e in errs.values()), errs


@pytest.mark.parametrize("tc_prec", TC_PRECS)
def test_full_size_properties_headline_config(tc_prec):
    """BASELINE full size (QIF N=4096, 1024 trials): properties that need no CPU reference run.
    (1) trials are independent: permuting the trial axis of the inputs permutes records and leaves dW unchanged (up to
        summation order); (2) the adjoint is linear in the output gradient; (3) the 3xTF32 path reproduces the FFMA path's
        spike rasters (counts identical, times within one step) on this size."""
    import rectipy_b200 as rp
    n, B, m, k, T, dt = 4096, 1024, 2, 3, 40, 1e-3
    rng = np.random.default_rng(99)
    W = (2.0 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    w_in, w_out = rng.standard_normal((n, m)).astype(np.float32), (rng.standard_normal((k, n)) / np.sqrt(n)).astype(np.float32)
    etas = orc.lorentzian_etas(n).astype(np.float32)
    t = np.arange(T, dtype=np.float32) * dt
    x = (rng.uniform(5, 15, (1, B, 1)) * np.sin(2 * np.pi * np.array([3.0, 5.0])[None, None, :] * t[:, None, None]
                                               + rng.uniform(0, 6.28, (1, B, m))) + 30.0).astype(np.float32)
    y0 = np.concatenate([rng.uniform(-50, 99, (B, n)), np.zeros((B, n))], axis=1).astype(np.float32)   # spread phases -> spikes within T
    gout = torch.tensor(rng.standard_normal((T, B, k)).astype(np.float32), device="cuda")

    def run(prec, xin, y_init, g, vars_=False):
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        node = net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=W, source_var="s", target_var="s_in",
                                   input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op",
                                   node_vars={"eta": etas}, train_params=["weights"])
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "qif", weights=w_in)
        net.add_func_node("out", k, "identity"); net.add_edge("qif", "out", weights=w_out, train="gd")
        node.reset(y_init)
        rec = [("qif", "v", False)] if vars_ else []
        obs = net.run(xin, verbose=False, enable_grad=True, record_vars=rec)
        out = torch.stack(obs["out"])
        (out * g).sum().backward()
        res = dict(out=out.detach(), gW=node["weights"].grad.clone(), gout=net.get_edge("qif", "out").weights.grad.clone())
        if vars_:
            res["v"] = torch.stack(obs[("qif", "v")])
        return res

    base = run(tc_prec, x, y0, gout, vars_=True)
    assert torch.isfinite(base["out"]).all() and torch.isfinite(base["gW"]).all()
    n_spikes = int((base["v"] >= 100.0).sum())
    assert n_spikes > 10000, n_spikes
    # (1) permutation of trials
    perm = torch.randperm(B, generator=torch.Generator().manual_seed(0))
    pr = run(tc_prec, x[:, perm.numpy(), :], y0[perm.numpy()], gout[:, perm.cuda(), :])
    assert torch.equal(pr["out"], base["out"][:, perm.cuda(), :])            # per-trial arithmetic is identical
    assert rel_err(pr["gW"].cpu().numpy(), base["gW"].cpu().numpy()) < 1e-4
    assert rel_err(pr["gout"].cpu().numpy(), base["gout"].cpu().numpy()) < 1e-4
    # (2) linearity of the adjoint in dL/dout
    lin = run(tc_prec, x, y0, 2.5 * gout)
    assert rel_err(lin["gW"].cpu().numpy(), 2.5 * base["gW"].cpu().numpy()) < 1e-5
    assert rel_err(lin["gout"].cpu().numpy(), 2.5 * base["gout"].cpu().numpy()) < 1e-5
    # (3) tensor-core path vs FFMA path: spike rasters of a sample of trials
    ff = run("fp32", x, y0, gout, vars_=True)
    for b in (0, 511, 1023):
        cmp_ = orc.compare_spikes((ff["v"][:, b, :] >= 100.0).cpu().numpy(), (base["v"][:, b, :] >= 100.0).cpu().numpy())
        assert cmp_["total_ref"] > 0 and cmp_["neurons_count_mismatch"] == 0 and cmp_["max_shift"] <= 1, cmp_
    assert rel_err(base["out"].cpu().numpy(), ff["out"].cpu().numpy()) < 1e-3


@pytest.mark.parametrize("tc_prec", TC_PRECS)
def test_full_size_rate_network_directional_derivative(tc_prec):
    """LI-tanh N=4096, 1024 trials on the tensor-core path: <dL/dW, D> from the adjoint equals the finite-difference
    directional derivative of the loss (fp32 central difference, so 2 % tolerance)."""
    import rectipy_b200 as rp
    n, B, T, dt = 4096, 1024, 25, 1e-2
    rng = np.random.default_rng(5)
    W = (1.5 * rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    D = (rng.standard_normal((n, n)) / np.sqrt(n)).astype(np.float32)
    w_out = (rng.standard_normal((2, n)) / np.sqrt(n)).astype(np.float32)
    x = torch.tensor(rng.standard_normal((T, B, n)).astype(np.float32), device="cuda")
    tgt = torch.tensor(rng.standard_normal((T, B, 2)).astype(np.float32), device="cuda")

    def loss_of(Wm, grad):
        net = rp.Network(dt, device="cuda:0", batch=B, precision=tc_prec)
        node = net.add_diffeq_node("rnn", "neuron_model_templates.rate_neurons.leaky_integrator.tanh", weights=Wm, source_var="tanh_op/r",
                                   target_var="li_op/r_in", input_var="li_op/I_ext", output_var="li_op/v",
                                   node_vars={"li_op/tau": 0.5, "li_op/k": 1.3}, train_params=["weights"] if grad else None)
        net.add_func_node("out", 2, "identity"); net.add_edge("rnn", "out", weights=w_out)
        obs = net.run(x, verbose=False, enable_grad=grad)
        loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]).double(), tgt.double(), reduction="sum")
        if grad:
            loss.backward()
            return float(loss), node["weights"].grad.double()
        return float(loss), None

    l0, gW = loss_of(W, True)
    analytic = float((gW * torch.tensor(D, device="cuda").double()).sum())
    eps = 2e-2
    lp, _ = loss_of(W + eps * D, False)
    lm, _ = loss_of(W - eps * D, False)
    numeric = (lp - lm) / (2 * eps)
    print(f"directional derivative: adjoint {analytic:.6e}  finite difference {numeric:.6e}")
    assert abs(analytic - numeric) <= 2e-2 * abs(numeric)


@pytest.mark.parametrize("tmpl", ["ik", "iku", "ik_biexp"])
def test_izhikevich_batched_paths_agree(tmpl):
    """ik_op / iku_op / ik_biexp_op on all execution paths (persistent B=2 -- per-step for iku and ik_biexp, whose recovery variable
    needs the population means of every step --, per-step FFMA B=20, tcgen05 B=128 N=128 in both operand formats) vs the fp64 oracle."""
    import rectipy_b200 as rp
    n, m, k, T, dt = 128, 2, 2, 300, 1e-1
    rng = np.random.default_rng(12)
    W = np.abs(rng.standard_normal((n, n))) * 4.0 / n
    w_in, w_out = rng.standard_normal((n, m)) * 10.0, rng.standard_normal((k, n)) / np.sqrt(n)
    etas = rng.uniform(60.0, 160.0, n)
    params = dict(eta=etas, g=1.5)
    xs = {B: (3.0 * np.sin(2 * np.pi * rng.uniform(5, 30, (1, B, m)) * (np.arange(T) * dt * 1e-3)[:, None, None]) + 1.0) for B in (2, 20, 128)}
    grads128 = {}
    for B, prec in ((2, "fp32"), (20, "fp32"), (128, "fp32"), (128, "3xtf32"), (128, "3xf16")):
        x = xs[B]
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        node = net.add_diffeq_node("ik", f"neuron_model_templates.spiking_neurons.ik.{tmpl}", weights=W, source_var="s", target_var="s_in",
                                   input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op=f"{tmpl}_op",
                                   node_vars={"eta": etas, "g": 1.5}, spike_threshold=40.0, spike_reset=-60.0,
                                   train_params=["weights", "eta", "g", "kappa"])
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "ik", weights=w_in)
        net.add_func_node("out", k, "identity"); net.add_edge("ik", "out", weights=w_out, train="gd")
        obs = net.run(x, sampling_steps=3, verbose=False, enable_grad=True, record_vars=[("ik", "u", False)])
        out = torch.stack(obs["out"])
        out.square().sum().backward()
        for b in (0, B - 1):
            onode = orc.make_node(tmpl, n, W, dt, params=params, dtype=torch.float64, train_params=["weights", "eta", "g", "kappa"],
                                  spike_threshold=40.0, spike_reset=-60.0)
            onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out, requires_grad=True))
            r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=3, record_vars=[("u", False)], enable_grad=False)
            ref = torch.stack(r["out"]).numpy()
            got = out.detach().cpu().numpy().reshape(ref.shape[0], B, k)[:, b, :]
            assert rel_err(got, ref) < 1e-4, (B, prec, b, rel_err(got, ref))
            u_ref = torch.stack(r["vars"]["u"]).numpy()
            u_got = obs.to_numpy(("ik", "u")).reshape(ref.shape[0], B, n)[:, b, :]
            assert rel_err(u_got, u_ref) < 1e-4
        assert all(torch.isfinite(p.grad).all() for p in net.parameters())
        if B == 128:
            grads128[prec] = [p.grad.detach().cpu().numpy().copy() for p in net.parameters()]
    for prec in ("3xtf32", "3xf16"):            # every gradient of the tensor-core paths vs the FFMA path (same trials)
        for g_tc, g_ff in zip(grads128[prec], grads128["fp32"]):
            assert rel_err(g_tc, g_ff) < 1e-3, (tmpl, prec, rel_err(g_tc, g_ff))


@pytest.mark.parametrize("model,n,B,prec", [("qif", 128, 128, "3xtf32"), ("qif", 128, 128, "3xf16"), ("li_tanh", 64, 1, "fp32"), ("qif_sfa", 96, 20, "fp32"), ("ik", 64, 2, "fp32"),
                                               ("ik_biexp", 64, 20, "fp32")])
def test_checkpoint_recompute_segments_match_full_history(model, n, B, prec, monkeypatch):
    """Long horizons run in segments (boundary checkpoints + recompute, engine.plan_segments).  With a tiny history budget the
    same run must give the same records and the same gradients as the single-segment run, on all three execution paths."""
    import rectipy_b200 as rp
    from rectipy_b200 import engine
    from golden_util import TEMPLATE_PATH
    rng = np.random.default_rng(n * 3 + B)
    rate = model.startswith("li_")
    ik = model.startswith("ik")          # ik, ik_biexp
    dt, T, S, cutoff, trunc = (1e-2, 230, 4, 6, 90) if rate else ((1e-1, 230, 4, 6, 90) if ik else (1e-3, 230, 3, 5, 100))
    m, k = 2, 2
    W = rng.standard_normal((n, n)) * (1.5 if rate else 2.0) / np.sqrt(n)
    if ik:
        W = np.abs(W) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)) * (10.0 if ik else 1.0), rng.standard_normal((k, n)) / np.sqrt(n)
    path, op, svar, tvar = TEMPLATE_PATH[model]
    params = {"li_tanh": dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1), "qif": dict(eta=orc.lorentzian_etas(n), k=1.5),
              "qif_sfa": dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.4, tau_x=1.2), "ik": dict(eta=rng.uniform(60, 160, n), g=1.5)}["ik" if ik else model]
    skw = dict(spike_threshold=40.0, spike_reset=-60.0) if ik else {}
    t = np.arange(T) * dt
    amp, off = (1.5, 0.0) if rate else ((3.0, 1.0) if ik else (10.0, 14.0))
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + off
    n_rec = len([s for s in range(T) if s >= cutoff and s % S == 0])
    targets = torch.tensor(rng.standard_normal((n_rec, B, k)), dtype=torch.float32, device="cuda")
    res = {}
    for mode in ("full", "segmented"):
        if mode == "segmented":
            nh = (5 if model == "ik_biexp" else 4) if ik else (3 if model == "qif_sfa" else (1 if rate else 2))
            monkeypatch.setenv("RECTIPY_B200_HISTORY_GB", repr(60 * nh * B * n * 4 / 2**30))      # ~57 steps per segment
        else:
            monkeypatch.delenv("RECTIPY_B200_HISTORY_GB", raising=False)
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext",
                  node_vars={f"{op}/{p}": v for p, v in params.items()}, train_params=["weights", f"{op}/eta"])
        if not rate:
            kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **skw)
        else:
            kw.update(output_var=f"{op}/v")
        node = net.add_diffeq_node("rnn", path, **kw)
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
        net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
        obs = net.run(x, sampling_steps=S, cutoff=cutoff, verbose=False, enable_grad=True,
                      record_vars=[("rnn", f"{op}/v", False), ("rnn", f"{op}/v" if rate else f"{op}/s", True)][:1 if rate else 2], truncate_steps=trunc)
        out = torch.stack(obs["out"])
        torch.nn.functional.mse_loss(out.reshape(n_rec, B, k), targets).backward()
        res[mode] = dict(out=out.detach().cpu().numpy(), var=obs.to_numpy(("rnn", f"{op}/v")), y=node.y.detach().cpu().numpy(),
                         gW=node["weights"].grad.cpu().numpy(), geta=node[f"{op}/eta"].grad.cpu().numpy(),
                         gin=net.get_edge("inp", "rnn").weights.grad.cpu().numpy(), gout=net.get_edge("rnn", "out").weights.grad.cpu().numpy())
    segs = engine.plan_segments(T, S, (4 if ik else 3) * B * n * 4, int(60 * 3 * B * n * 4))
    assert len(segs) >= 3
    assert np.array_equal(res["full"]["y"], res["segmented"]["y"]) and np.array_equal(res["full"]["var"], res["segmented"]["var"])
    errs = {key: rel_err(res["segmented"][key], res["full"][key]) for key in res["full"]}
    print(model, errs)
    assert all(e <= 2e-5 for e in errs.values()), errs


@pytest.mark.parametrize("model,n,B,prec", [("qif", 40, 3, "fp32"), ("li_tanh", 50, 20, "fp32"), ("qif", 128, 128, "3xtf32"), ("li_sigmoid", 128, 128, "3xtf32"),
                                            ("qif", 128, 128, "3xf16"), ("li_sigmoid", 128, 128, "3xf16")])
def test_parameter_sweep_per_trial_values(model, n, B, prec):
    """Parameter sweep: every trial carries its own parameter values ([B,1] and [B,n] tensors).  Each trial must equal the
    reference path run on its own with that trial's parameters; the recurrent weights (and their gradient) are shared."""
    import rectipy_b200 as rp
    from golden_util import TEMPLATE_PATH
    rng = np.random.default_rng(n + 7 * B)
    rate = model.startswith("li_")
    dt, T, S = (1e-2, 100, 2) if rate else (1e-3, 300, 3)
    m, k = 2, 2
    W = rng.standard_normal((n, n)) * (1.5 if rate else 2.0) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    path, op, svar, tvar = TEMPLATE_PATH[model]
    eta_sweep = (rng.uniform(-1, 1, (B, 1)) if rate else rng.uniform(-8, 2, (B, 1)))                 # one value per trial
    tau_sweep = rng.uniform(0.8, 2.0, (B, n))                                                       # per trial and neuron
    node_vars = {f"{op}/eta": eta_sweep, f"{op}/tau": tau_sweep, f"{op}/k": 1.3}
    if model == "li_sigmoid":
        node_vars["sigmoid_op/r_max"] = rng.uniform(0.5, 2.0, (B, 1))
    t = np.arange(T) * dt
    amp, off = (1.5, 0.0) if rate else (10.0, 14.0)
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + off
    net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
    kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext", node_vars=node_vars, train_params=["weights"])
    if not rate:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s")
    else:
        kw.update(output_var=f"{op}/v")
    node = net.add_diffeq_node("rnn", path, **kw)
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in)
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out)
    obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True)
    out = torch.stack(obs["out"])
    out.square().sum().backward()
    gW = node["weights"].grad.cpu().numpy()
    gW_ref = np.zeros_like(gW, dtype=np.float64)
    sample = range(B) if B <= 20 else (0, 1, B // 2, B - 1)
    for b in sample:
        params = dict(eta=float(eta_sweep[b, 0]), tau=tau_sweep[b], k=1.3)
        if model == "li_sigmoid":
            params["r_max"] = float(node_vars["sigmoid_op/r_max"][b, 0])
        onode = orc.make_node(model, n, W, dt, params=params, dtype=torch.float64, train_params=["weights"])
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out))
        r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=True)
        ref = torch.stack(r["out"])
        ref.square().sum().backward()
        gW_ref += onode.get("weights").grad.numpy()
        assert rel_err(out[:, b, :].detach().cpu().numpy(), ref.detach().numpy()) < (1e-5 if rate else 1e-4), (b,)
    if B <= 20:
        assert rel_err(gW, gW_ref) < (1e-5 if rate else 2e-3)
    with pytest.raises(NotImplementedError):
        rp.Network(dt, device="cuda:0", batch=B).add_diffeq_node("rnn", path, **{**kw, "train_params": [f"{op}/eta"]})
    with pytest.raises(RuntimeError):      # the coupling constant is folded into the shared weights
        bad = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        bad.add_diffeq_node("rnn", path, **{**kw, "node_vars": {f"{op}/k": np.ones((B, 1))}, "train_params": None})
        bad.run(x, verbose=False, enable_grad=False)
