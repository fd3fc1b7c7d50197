"""Shapes the tcgen05 kernels do not take as they are (trial count or n not a multiple of 128) and the persistent few-trial kernels
cannot hold: precision="auto" pads the trial / neuron axes to multiples of 128 and runs the tensor-core pass (network._padded_shape)
instead of one fp32 contraction launch per step.  The padded entries are inert; every real trial must still equal the unbatched
fp64 oracle run on its own (rate <= 1e-5; spiking: identical spike raster at the record steps, outputs <= 1e-4, gradients <= 2e-3)."""
import numpy as np
import pytest
import torch

from golden_util import TEMPLATE_PATH, rel_err, orc

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("model,n,B,dense", [("li_tanh", 500, 100, False), ("qif", 600, 96, False), ("qif_sfa", 520, 130, True)])
def test_padded_tensor_core_pass_matches_per_trial_oracle(model, n, B, dense):
    import rectipy_b200 as rp
    from rectipy_b200 import engine, _cabi as abi
    rng = np.random.default_rng(n + B)
    rate = model.startswith("li_")
    m, k = 2, 3
    dt, T, S = (1e-2, 120, 3) if rate else (1e-3, 300, 4)
    W = rng.standard_normal((n, n)) * (1.5 if rate else 2.0) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    params = {"li_tanh": dict(tau=rng.uniform(1, 2, n), k=1.2, eta=0.1), "qif": dict(eta=orc.lorentzian_etas(n), k=1.5, tau_s=0.7),
              "qif_sfa": dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.4, tau_x=1.2)}[model]
    amp, off = (1.5, 0.0) if rate else (10.0, 14.0)
    t = np.arange(T) * dt
    x = amp * np.sin(2 * np.pi * rng.uniform(0.5, 3, (1, B, m)) * t[:, None, None] + rng.uniform(0, 6.28, (1, B, m))) + off
    n_rec = len(range(0, T, S))
    path, op, svar, tvar = TEMPLATE_PATH[model]
    engine.clear_plans()
    net = rp.Network(dt, device="cuda:0", batch=B)                       # precision="auto"
    kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext", node_vars={f"{op}/{p}": v for p, v in params.items()},
              train_params=["weights", f"{op}/eta"])
    if rate:
        kw.update(output_var=f"{op}/v")
    else:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s")
    node = net.add_diffeq_node("rnn", path, **kw)
    rec_name = "v"
    if dense:       # no projection nodes: dense input current, dense output record
        xin = np.einsum("tbm,nm->tbn", x, w_in)
        obs = net.run(xin, sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", rec_name, True)])
        targets = rng.standard_normal((n_rec, B, n))
    else:
        net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
        net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
        obs = net.run(x, sampling_steps=S, verbose=False, enable_grad=True, record_vars=[("rnn", rec_name, True)])
        targets = rng.standard_normal((n_rec, B, k))
    keys = list(engine._PLANS.keys())
    assert len(keys) == 1 and keys[0].n % 128 == 0 and keys[0].batch % 128 == 0 and keys[0].precision == abi.RP_PREC_3XF16, keys
    assert node.state.shape == (node.spec.n_sv, B, n)
    out = torch.stack(obs["out"])
    assert out.shape == targets.shape
    rec = obs.to_numpy(("rnn", rec_name))
    torch.nn.functional.mse_loss(out, torch.tensor(targets, dtype=torch.float32, device="cuda:0")).backward()

    sample = sorted(set([0, 1, B // 2, B - 1]))
    gW_eng = node["weights"].grad.cpu().numpy()
    for b in sample:
        onode = orc.make_node(model, n, W, dt, params=params, dtype=torch.float64, train_params=["weights", "eta"])
        if dense:
            onet = orc.OracleNet(onode)
            r = onet.run(torch.tensor(np.einsum("tm,nm->tn", x[:, b, :], w_in)), sampling_steps=S, enable_grad=False, record_vars=[("v", False)])
        else:
            onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out))
            r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=False, record_vars=[("v", False)])
        ref = torch.stack(r["out"]).numpy()
        v_ref = torch.stack(r["vars"]["v"]).numpy()
        e_out = rel_err(out[:, b].detach().cpu().numpy(), ref)
        e_mean = rel_err(rec.reshape(n_rec, B)[:, b], v_ref.mean(axis=-1))      # neuron means over the REAL neurons only
        assert e_out <= (1e-5 if rate else 1e-4) and e_mean <= (1e-5 if rate else 1e-4), (b, e_out, e_mean)
    print(model, "padded plan", (keys[0].n, keys[0].batch), "out err", e_out, "mean-v err", e_mean)

    # gradients: all trials, per-trial oracle (few trials only when the oracle is slow)
    if B <= 100 and not dense:
        g_ref = np.zeros_like(gW_eng, dtype=np.float64)
        for b in range(B):
            onode = orc.make_node(model, n, W, dt, params=params, dtype=torch.float64, train_params=["weights"])
            onet = orc.OracleNet(onode, w_in=torch.tensor(w_in), w_out=torch.tensor(w_out))
            r = onet.run(torch.tensor(x[:, b, :]), sampling_steps=S, enable_grad=True)
            pred = torch.stack(r["out"])
            (torch.nn.functional.mse_loss(pred, torch.tensor(targets[:, b, :]), reduction="sum") / out.numel()).backward()
            g_ref += onode.get("weights").grad.numpy()
        e_g = rel_err(gW_eng, g_ref)
        print(model, "dW err", e_g)
        assert e_g <= (1e-5 if rate else 2e-3)


def test_recorded_neuron_means_ignore_padded_neurons():
    import rectipy_b200 as rp
    from rectipy_b200 import engine
    n, B, T, dt = 450, 100, 64, 1e-2
    rng = np.random.default_rng(5)
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    path, op, svar, tvar = TEMPLATE_PATH["li_tanh"]
    x = rng.standard_normal((T, B, n)).astype(np.float32)
    res = {}
    for prec in ("auto", "fp32"):
        engine.clear_plans()
        net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
        net.add_diffeq_node("rnn", path, weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext", output_var=f"{op}/v",
                            node_vars={f"{op}/eta": rng.standard_normal(n) * 0 + 0.3})
        obs = net.run(x, sampling_steps=2, verbose=False, record_vars=[("rnn", "v", True)])
        res[prec] = (obs.to_numpy("out"), [obs.to_numpy(("rnn", "v"))], list(engine._PLANS.keys())[0])
    assert res["auto"][2].n == 512 and res["auto"][2].batch == 128 and res["fp32"][2].n == n
    assert rel_err(res["auto"][0], res["fp32"][0]) <= 1e-5
    assert rel_err(res["auto"][1][0], res["fp32"][1][0]) <= 1e-5
