"""Mint the golden fixtures under tests/golden/ from the UNMODIFIED reference classes.  TEST INFRASTRUCTURE ONLY.

Run in the build container (needs /root/reference):   python oracle/make_golden.py

For every case the reference's own `Network` / `RateNet` / `SpikeResetNet` / `Linear` / `RLS` / `Observer`
(imported via oracle/ref_shim.py) are driven exactly like a user would (`add_node`, `add_func_node`, `add_edge`,
`run`, `loss.backward()`), with the hand-written vector fields of oracle/rectipy_oracle.py supplying the part
PyRates would have generated.  Inputs AND outputs are stored, so tests never need the reference at run time.
All cases are seeded; fixtures are fp64 truth plus the fp32 reference run of the same inputs.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
import rectipy_oracle as orc  # noqa: E402

OUT = os.path.join(os.path.dirname(HERE), "tests", "golden")
TD = {"float64": torch.float64, "float32": torch.float32}


def ref_network(ref, model, n, W, dt, dtype, params=None, train_params=None, w_in=None, w_out=None,
                train_in=False, train_out=False, in_act="identity", out_act="identity", spike_kwargs=None,
                input_var="I_ext", output_var=None, in_mask=None):
    """Assemble a reference Network around a reference RateNet/SpikeResetNet built on an oracle vector field."""
    func, args, var_map, param_map = orc.build_node_args(model, n, W, params, dtype, input_var, output_var)
    net = ref.Network(dt, device="cpu", dtype=dtype)
    if model in orc.SPIKING:
        node = ref.nodes.SpikeResetNet(func, args, var_map, param_map, dt=dt, dtype=dtype, train_params=train_params,
                                       device="cpu", **(spike_kwargs or {}))
    else:
        node = ref.nodes.RateNet(func, args, var_map, param_map, dt=dt, dtype=dtype, train_params=train_params,
                                 device="cpu")
    net.add_node("rnn", node, node_type="diff_eq")
    if w_in is not None:
        net.add_func_node("inp", w_in.shape[1], in_act)
        kw = {"mask": in_mask} if in_mask is not None else {}
        net.add_edge("inp", "rnn", weights=np.asarray(w_in), train="gd" if train_in else None, **kw)
    if w_out is not None:
        net.add_func_node("out", w_out.shape[0], out_act)
        net.add_edge("rnn", "out", weights=np.asarray(w_out), train="gd" if train_out else None)
    return net, node


def run_case(ref, spec, dtype_name):
    dtype = TD[dtype_name]
    n, T = spec["n"], spec["T"]
    params = {k: v for k, v in spec.get("params", {}).items()}
    net, node = ref_network(ref, spec["model"], n, spec["W"], spec["dt"], dtype, params=params,
                            train_params=spec.get("train_params"), w_in=spec.get("w_in"), w_out=spec.get("w_out"),
                            train_in=spec.get("train_in", False), train_out=spec.get("train_out", False),
                            in_act=spec.get("in_act", "identity"), out_act=spec.get("out_act", "identity"),
                            spike_kwargs=spec.get("spike_kwargs"), in_mask=spec.get("in_mask"))
    inputs = torch.tensor(spec["inputs"], dtype=dtype)
    rec = [("rnn", v, red) for v, red in spec.get("record_vars", [])]
    kw = {}
    if "truncate_steps" in spec:
        kw["truncate_steps"] = spec["truncate_steps"]
    obs = net.run(inputs, sampling_steps=spec.get("S", 1), cutoff=spec.get("cutoff", 0), verbose=False,
                  enable_grad=spec.get("grad", False), record_vars=rec, **kw)
    res = {"out": obs.to_numpy("out"), "steps": np.asarray(obs["steps"])}
    for v, red in spec.get("record_vars", []):
        res[f"var_{v}"] = obs.to_numpy(("rnn", v))
    if spec["model"] in orc.SPIKING:
        res["y_final"] = torch.cat((node._y_start, node._y_spike, node._y_stop), 0).detach().numpy()
    else:
        res["y_final"] = node.y.detach().numpy()
    if spec.get("grad", False):
        target = torch.tensor(spec["targets"], dtype=dtype)
        loss = torch.nn.MSELoss()(torch.stack(obs["out"]), target)
        loss.backward()
        res["loss"] = loss.detach().numpy()
        for name in spec.get("train_params", []) or []:
            res[f"grad_{name}"] = node[name].grad.detach().numpy()
        if spec.get("train_in"):
            res["grad_w_in"] = net.get_edge("inp", "rnn").weights.grad.detach().numpy()
        if spec.get("train_out"):
            res["grad_w_out"] = net.get_edge("rnn", "out").weights.grad.detach().numpy()
    return res


def save_case(ref, name, spec):
    blob = {}
    for k, v in spec.items():
        if isinstance(v, np.ndarray):
            blob[f"in_{k}"] = v
    for k, v in spec.get("params", {}).items():
        blob[f"param_{k}"] = np.asarray(v, dtype=np.float64)
    meta = {k: v for k, v in spec.items() if not isinstance(v, np.ndarray) and k != "params"}
    blob["meta"] = np.asarray(repr(meta))
    for dn in ("float64", "float32"):
        res = run_case(ref, spec, dn)
        for k, v in res.items():
            blob[f"{dn}_{k}"] = v
    path = os.path.join(OUT, f"{name}.npz")
    np.savez_compressed(path, **blob)
    print(f"{name}: wrote {os.path.getsize(path)/1024:.1f} KiB;  out {blob['float64_out'].shape}")


def save_two_node_chain(ref):
    """G9: a feed-forward chain of TWO differential-equation nodes, the general topology the reference's graph walk can
    execute (rectipy/network.py:962-973; documentation/rnn_tryout.py):  inp -> Linear -> LI-tanh (RateNet, output v) ->
    Linear -> QIF (SpikeResetNet, output s) -> Linear -> out.  BPTT through both nodes and all three edges."""
    rng = np.random.default_rng(4242)
    n1, n2, m, k, T, dt, S, cutoff = 10, 8, 2, 3, 900, 1e-3, 3, 4
    W1 = rng.standard_normal((n1, n1)) * 1.5 / np.sqrt(n1)
    W2 = rng.standard_normal((n2, n2)) * 2.0 / np.sqrt(n2)
    p1 = dict(tau=rng.uniform(0.02, 0.05, n1), k=1.2, eta=0.3)
    p2 = dict(eta=orc.lorentzian_etas(n2) + 20.0, k=1.5, tau_s=0.7)
    w_in, w12, w_out = rng.standard_normal((n1, m)), rng.standard_normal((n2, n1)) * 8.0, rng.standard_normal((k, n2)) / np.sqrt(n2)
    inputs = sin_inputs(rng, T, m, dt, amp=3.0, offset=1.0)
    targets = rng.standard_normal((len([t for t in range(T) if t >= cutoff and t % S == 0]), k))
    blob = dict(in_W1=W1, in_W2=W2, in_w_in=w_in, in_w12=w12, in_w_out=w_out, in_inputs=inputs, in_targets=targets,
                p1_tau=np.asarray(p1["tau"]), p2_eta=np.asarray(p2["eta"]),
                meta=np.asarray(repr(dict(n1=n1, n2=n2, m=m, k=k, T=T, dt=dt, S=S, cutoff=cutoff, p1_k=1.2, p1_eta=0.3, p2_k=1.5, p2_tau_s=0.7))))
    for dn in ("float64", "float32"):
        dtype = TD[dn]
        net = ref.Network(dt, device="cpu", dtype=dtype)
        f1, a1, vm1, pm1 = orc.build_node_args("li_tanh", n1, W1, p1, dtype, "I_ext", None)
        node1 = ref.nodes.RateNet(f1, a1, vm1, pm1, dt=dt, dtype=dtype, train_params=["weights", "tau"], device="cpu")
        f2, a2, vm2, pm2 = orc.build_node_args("qif", n2, W2, p2, dtype, "I_ext", None)
        node2 = ref.nodes.SpikeResetNet(f2, a2, vm2, pm2, dt=dt, dtype=dtype, train_params=["weights", "eta"], device="cpu")
        net.add_node("rate", node1, node_type="diff_eq")
        net.add_node("spk", node2, node_type="diff_eq")
        net.add_func_node("inp", m, "identity"); net.add_func_node("out", k, "identity")
        net.add_edge("inp", "rate", weights=w_in, train="gd")
        net.add_edge("rate", "spk", weights=w12, train="gd")
        net.add_edge("spk", "out", weights=w_out, train="gd")
        obs = net.run(torch.tensor(inputs, dtype=dtype), sampling_steps=S, cutoff=cutoff, verbose=False, enable_grad=True,
                      record_vars=[("rate", "v", False), ("spk", "s", True)])
        out = torch.stack(obs["out"])
        loss = torch.nn.MSELoss()(out, torch.tensor(targets, dtype=dtype))
        loss.backward()
        res = dict(out=out.detach().numpy(), steps=np.asarray(obs["steps"]), var_rate_v=obs.to_numpy(("rate", "v")),
                   var_spk_s=obs.to_numpy(("spk", "s")), loss=loss.detach().numpy(),
                   grad_W1=node1["weights"].grad.numpy(), grad_tau1=node1["tau"].grad.numpy(),
                   grad_W2=node2["weights"].grad.numpy(), grad_eta2=node2["eta"].grad.numpy(),
                   grad_w_in=net.get_edge("inp", "rate").weights.grad.numpy(), grad_w12=net.get_edge("rate", "spk").weights.grad.numpy(),
                   grad_w_out=net.get_edge("spk", "out").weights.grad.numpy(),
                   n_spikes=np.asarray(float(obs.to_numpy(("spk", "s")).max())))
        for kk, v in res.items():
            blob[f"{dn}_{kk}"] = v
    np.savez_compressed(os.path.join(OUT, "two_node_chain.npz"), **blob)
    print("two_node_chain: out", blob["float64_out"].shape, "max mean-s", float(blob["float64_var_spk_s"].max()))


def save_feedback_net(ref):
    """G10: the reference's own `FeedbackNetwork` (rectipy/network.py:1196-1357; documentation/rnn_tryout.py):
    inp -> Linear -> p1 -> Linear -> p2 -> Linear -> out  plus a FEEDBACK edge p2 -> p1, BPTT through both nodes and all four edges.
    Variant A: p1 LI-tanh (RateNet, output v), p2 QIF (SpikeResetNet, output s) -- the feedback reads the spiking node's `y`, which
    the reference leaves at the state BEFORE that node's last step (nodes.py:387): one step of delay.
    Variant B: p1 QIF, p2 LI-tanh -- the feedback reads the rate node's `y`, which is its current state (nodes.py:169): no delay."""
    rng = np.random.default_rng(5151)
    n1, n2, m, k, T, dt, S, cutoff = 10, 8, 2, 3, 700, 1e-3, 3, 4
    blob = {}
    for variant in ("A", "B"):
        nr, nq = (n1, n2) if variant == "A" else (n2, n1)          # rate / spiking population sizes
        Wr = rng.standard_normal((nr, nr)) * 1.5 / np.sqrt(nr)
        Wq = rng.standard_normal((nq, nq)) * 2.0 / np.sqrt(nq)
        pr = dict(tau=rng.uniform(0.02, 0.05, nr), k=1.2, eta=0.3)
        pq = dict(eta=orc.lorentzian_etas(nq) + 20.0, k=1.5, tau_s=0.7)
        if variant == "A":
            w_in, w12, wfb = rng.standard_normal((n1, m)), rng.standard_normal((n2, n1)) * 8.0, rng.standard_normal((n1, n2)) * 0.5
        else:
            w_in, w12, wfb = rng.standard_normal((n1, m)) * 8.0, rng.standard_normal((n2, n1)) * 0.3, rng.standard_normal((n1, n2)) * 6.0
        w_out = rng.standard_normal((k, n2)) / np.sqrt(n2)
        inputs = sin_inputs(rng, T, m, dt, amp=3.0, offset=1.0)
        targets = rng.standard_normal((len([t for t in range(T) if t >= cutoff and t % S == 0]), k))
        pre = f"{variant}_"
        blob.update({pre + "in_Wr": Wr, pre + "in_Wq": Wq, pre + "in_w_in": w_in, pre + "in_w12": w12, pre + "in_wfb": wfb, pre + "in_w_out": w_out,
                     pre + "in_inputs": inputs, pre + "in_targets": targets, pre + "pr_tau": np.asarray(pr["tau"]), pre + "pq_eta": np.asarray(pq["eta"])})
        for dn in ("float64", "float32"):
            dtype = TD[dn]
            net = ref.FeedbackNetwork(dt, device="cpu")
            net.dtype = dtype                      # FeedbackNetwork.__init__ does not forward a dtype (network.py:1198-1200)
            fr, ar, vmr, pmr = orc.build_node_args("li_tanh", nr, Wr, pr, dtype, "I_ext", None)
            rate = ref.nodes.RateNet(fr, ar, vmr, pmr, dt=dt, dtype=dtype, train_params=["weights", "tau"], device="cpu")
            fq, aq, vmq, pmq = orc.build_node_args("qif", nq, Wq, pq, dtype, "I_ext", None)
            spk = ref.nodes.SpikeResetNet(fq, aq, vmq, pmq, dt=dt, dtype=dtype, train_params=["weights", "eta"], device="cpu")
            first, second = (("rate", rate), ("spk", spk)) if variant == "A" else (("spk", spk), ("rate", rate))
            net.add_node(first[0], first[1], node_type="diff_eq")
            net.add_node(second[0], second[1], node_type="diff_eq")
            net.add_func_node("inp", m, "identity"); net.add_func_node("out", k, "identity")
            net.add_edge("inp", first[0], weights=w_in, train="gd")
            net.add_edge(first[0], second[0], weights=w12, train="gd")
            net.add_edge(second[0], "out", weights=w_out, train="gd")
            net.add_edge(second[0], first[0], weights=wfb, train="gd", feedback=True)
            obs = net.run(torch.tensor(inputs, dtype=dtype), sampling_steps=S, cutoff=cutoff, verbose=False, enable_grad=True,
                          record_vars=[("rate", "v", False), ("spk", "s", True)])
            out = torch.stack(obs["out"])
            loss = torch.nn.MSELoss()(out, torch.tensor(targets, dtype=dtype))
            loss.backward()
            res = dict(out=out.detach().numpy(), steps=np.asarray(obs["steps"]), var_rate_v=obs.to_numpy(("rate", "v")),
                       var_spk_s=obs.to_numpy(("spk", "s")), loss=loss.detach().numpy(),
                       grad_Wr=rate["weights"].grad.numpy(), grad_tau=rate["tau"].grad.numpy(),
                       grad_Wq=spk["weights"].grad.numpy(), grad_eta=spk["eta"].grad.numpy(),
                       grad_w_in=net.get_edge("inp", first[0]).weights.grad.numpy(), grad_w12=net.get_edge(first[0], second[0]).weights.grad.numpy(),
                       grad_wfb=net.get_edge(second[0], first[0]).weights.grad.numpy(),
                       grad_w_out=net.get_edge(second[0], "out").weights.grad.numpy())
            for kk, v in res.items():
                blob[f"{pre}{dn}_{kk}"] = v
        print(f"feedback_net {variant}: out", blob[pre + "float64_out"].shape, "max mean-s", float(blob[pre + "float64_var_spk_s"].max()),
              "|grad_wfb|", float(np.abs(blob[pre + "float64_grad_wfb"]).max()),
              "fp32 vs fp64 out", float(np.abs(blob[pre + "float32_out"] - blob[pre + "float64_out"]).max()))
    blob["meta"] = np.asarray(repr(dict(n1=n1, n2=n2, m=m, k=k, T=T, dt=dt, S=S, cutoff=cutoff, pr_k=1.2, pr_eta=0.3, pq_k=1.5, pq_tau_s=0.7)))
    np.savez_compressed(os.path.join(OUT, "feedback_net.npz"), **blob)


def save_multispike(ref):
    """G12: the reference's MultiSpikeResetNet (rectipy/nodes.py:404-465) around a two-population QIF field with two spike /
    reset variable pairs, inside a reference Network (inp -> Linear -> node -> Linear -> out), BPTT.  Outputs are post-update."""
    rng = np.random.default_rng(1212)
    n, m, k, T, dt, S = 16, 2, 2, 1200, 1e-3, 3
    W = rng.standard_normal((n, n)) * 2.0 / np.sqrt(n)
    params = dict(eta_e=orc.lorentzian_etas(n) + 8.0, eta_i=rng.uniform(-2.0, 6.0, n), tau_e=1.0, J_ee=1.3, J_ei=2.0, tau_i=0.5, J_ie=3.0, tau_s=0.8)
    w_in, w_out = rng.standard_normal((n, m)), rng.standard_normal((k, n)) / np.sqrt(n)
    inputs = sin_inputs(rng, T, m, dt, amp=10.0, offset=12.0)
    targets = rng.standard_normal((len(range(0, T, S)), k))
    train = ["weights", "eta_e", "J_ei", "tau_i"]
    blob = dict(in_W=W, in_w_in=w_in, in_w_out=w_out, in_inputs=inputs, in_targets=targets, param_eta_e=params["eta_e"], param_eta_i=params["eta_i"],
                meta=np.asarray(repr(dict(n=n, m=m, k=k, T=T, dt=dt, S=S, thresh=100.0, reset=-100.0, train=train,
                                          **{q: v for q, v in params.items() if np.ndim(v) == 0}))))
    for dn in ("float64", "float32"):
        dtype = TD[dn]
        func, args, var_map, param_map = orc.build_ei_node_args(n, W, params, dtype)
        node = ref.nodes.MultiSpikeResetNet(func, args, var_map, param_map, dt=dt, dtype=dtype, train_params=train, device="cpu",
                                            spike_threshold=100.0, spike_reset=-100.0)
        net = ref.Network(dt, device="cpu", dtype=dtype)
        net.add_node("rnn", node, node_type="diff_eq")
        net.add_func_node("inp", m, "identity"); net.add_func_node("out", k, "identity")
        net.add_edge("inp", "rnn", weights=w_in, train="gd")
        net.add_edge("rnn", "out", weights=w_out, train="gd")
        obs = net.run(torch.tensor(inputs, dtype=dtype), sampling_steps=S, verbose=False, enable_grad=True,
                      record_vars=[("rnn", "v_e", False), ("rnn", "v_i", False)])
        out = torch.stack(obs["out"])
        loss = torch.nn.MSELoss()(out, torch.tensor(targets, dtype=dtype))
        loss.backward()
        res = dict(out=out.detach().numpy(), steps=np.asarray(obs["steps"]), var_v_e=obs.to_numpy(("rnn", "v_e")), var_v_i=obs.to_numpy(("rnn", "v_i")),
                   y_final=node.y.detach().numpy(), loss=loss.detach().numpy(),
                   grad_w_in=net.get_edge("inp", "rnn").weights.grad.numpy(), grad_w_out=net.get_edge("rnn", "out").weights.grad.numpy())
        for name in train:
            res[f"grad_{name}"] = node[name].grad.detach().numpy()
        # the restatement against the class it restates
        onode = orc.OracleMultiSpikeResetNode(*orc.build_ei_node_args(n, W, params, dtype), dt, dtype, train, spike_threshold=100.0, spike_reset=-100.0)
        onet = orc.OracleNet(onode, w_in=torch.tensor(w_in, dtype=dtype, requires_grad=True), w_out=torch.tensor(w_out, dtype=dtype, requires_grad=True))
        r = onet.run(torch.tensor(inputs, dtype=dtype), sampling_steps=S, enable_grad=True)
        o_out = torch.stack(r["out"])
        torch.nn.MSELoss()(o_out, torch.tensor(targets, dtype=dtype)).backward()
        tol = 1e-12 if dn == "float64" else 1e-5
        assert np.max(np.abs(o_out.detach().numpy() - res["out"])) <= tol * max(1.0, np.abs(res["out"]).max()), "oracle restatement deviates (outputs)"
        assert np.max(np.abs(onode.get("weights").grad.numpy() - res["grad_weights"])) <= tol * np.abs(res["grad_weights"]).max(), "oracle restatement deviates (dW)"
        for kk, v in res.items():
            blob[f"{dn}_{kk}"] = v
    np.savez_compressed(os.path.join(OUT, "multispike_ei.npz"), **blob)
    n_spk = int((blob["float64_var_v_e"] == -100.0).sum() + (blob["float64_var_v_i"] == -100.0).sum())
    print("multispike_ei: out", blob["float64_out"].shape, "resets seen at record steps:", n_spk)


def save_edges_square(ref):
    """Square, non-symmetric edge weights: the reference transposes whenever shape == (n_in, n_out), which is always
    true for an N x N matrix (edges.py:22-23,160-161) -- a user's W therefore acts as W.T.  Outputs of the unmodified classes."""
    rng = np.random.default_rng(4242)
    n, steps = 6, 5
    w = rng.standard_normal((n, n))
    mask = (rng.uniform(size=(n, n)) < 0.5).astype(np.float64)
    xs = rng.standard_normal((steps, n))
    lin = ref.edges.Linear(n, n, weights=w.copy(), dtype=torch.float64)
    msk = ref.edges.LinearMasked(n, n, mask=mask.copy(), weights=w.copy(), dtype=torch.float64)
    np.savez_compressed(os.path.join(OUT, "edges_square.npz"), w=w, mask=mask, xs=xs,
                        lin_out=np.stack([lin.forward(torch.tensor(x)).numpy() for x in xs]),
                        masked_out=np.stack([msk.forward(torch.tensor(x)).numpy() for x in xs]),
                        lin_weights=lin.weights.numpy().copy(), masked_mask=msk.mask.numpy().copy())
    print("edges_square: ok")


def sin_inputs(rng, T, m, dt, amp=1.0, offset=0.0):
    t = np.arange(T) * dt
    freqs = rng.uniform(0.5, 3.0, size=m)
    ph = rng.uniform(0, 2 * np.pi, size=m)
    return amp * np.sin(2 * np.pi * freqs[None, :] * t[:, None] + ph[None, :]) + offset


def main(only=None):
    """`only`: regenerate a single case that has its own random stream (currently: ik_bptt, iku_bptt, ik_biexp_bptt, two_node_chain, feedback_net)."""
    ref = ref_shim.import_reference()
    if only is not None:
        global save_case
        _real_save = save_case

        def save_case(ref_, name, spec):      # noqa: F811  -- skip every other case
            if name == only:
                _real_save(ref_, name, spec)
    os.makedirs(OUT, exist_ok=True)
    rng = np.random.default_rng(20261018)

    # ---- G1: LI-tanh rate net, BPTT with windowed readout (SURVEY Appendix A.2 set-up) ------------
    n, T, m, k, dt = 32, 600, 3, 2, 1e-2
    spec = dict(model="li_tanh", n=n, T=T, dt=dt, S=5, cutoff=7, grad=True,
                W=rng.standard_normal((n, n)) / np.sqrt(n) * 1.5,
                params=dict(tau=rng.uniform(1.0, 2.0, n), k=1.3, eta=0.2),
                train_params=["weights", "tau", "k", "eta"], train_in=True, train_out=True,
                w_in=rng.standard_normal((n, m)), w_out=rng.standard_normal((k, n)) / np.sqrt(n),
                inputs=sin_inputs(rng, T, m, dt, amp=2.0), record_vars=[("v", False)])
    n_rec = len([s for s in range(T) if s >= 7 and s % 5 == 0])
    spec["targets"] = rng.standard_normal((n_rec, k))
    save_case(ref, "li_tanh_bptt", spec)

    # ---- G2: LI-sigmoid, forward only, direct input on the node, no readout (ridge-style X) -------
    n, T, dt = 20, 300, 5e-2
    spec = dict(model="li_sigmoid", n=n, T=T, dt=dt, S=3, cutoff=0, grad=False,
                W=rng.standard_normal((n, n)) / np.sqrt(n) * 2.0,
                params=dict(tau=2.0, k=1.5, eta=rng.standard_normal(n) * 0.3, r_max=2.0, s=1.5, v0=0.2),
                inputs=sin_inputs(rng, T, n, dt, amp=1.0), record_vars=[("v", True)])
    save_case(ref, "li_sigmoid_fwd", spec)

    # ---- G3: QIF spiking net, BPTT through surrogate + reset gate ---------------------------------
    n, T, m, k, dt = 24, 1500, 2, 3, 1e-3
    spec = dict(model="qif", n=n, T=T, dt=dt, S=4, cutoff=0, grad=True,
                W=rng.standard_normal((n, n)) * 2.0 / np.sqrt(n),
                params=dict(eta=orc.lorentzian_etas(n), tau=1.0, k=1.2, tau_s=0.8),
                train_params=["weights", "eta", "tau", "k", "tau_s"], train_in=True, train_out=True,
                w_in=rng.standard_normal((n, m)), w_out=rng.standard_normal((k, n)) / np.sqrt(n),
                inputs=sin_inputs(rng, T, m, dt, amp=10.0, offset=14.0), record_vars=[("v", False), ("s", True)])
    spec["targets"] = rng.standard_normal((len(range(0, T, 4)), k))
    save_case(ref, "qif_bptt", spec)

    # ---- G4: QIF-SFA, config-1 style (random_connectivity p, k=15, step input through tanh node) ---
    n, T, dt = 100, 4000, 1e-3
    np.random.seed(7)
    W = ref.random_connectivity(n, n, 0.2, normalize=True)
    inp = np.zeros((T, 1))
    inp[1000:3000, 0] = 3.0
    spec = dict(model="qif_sfa", n=n, T=T, dt=dt, S=100, cutoff=0, grad=False, W=W,
                params=dict(eta=orc.lorentzian_etas(n), k=15.0, alpha=0.3, tau_x=2.0),
                w_in=rng.standard_normal((n, 1)), in_act="tanh", inputs=inp,
                spike_kwargs=dict(spike_threshold=100.0, spike_reset=-100.0),
                record_vars=[("s", True), ("v", False)])
    save_case(ref, "qif_sfa_fwd", spec)

    # ---- G5: QIF-SFA BPTT with alpha, truncated BPTT ----------------------------------------------
    n, T, m, k, dt = 16, 800, 2, 2, 1e-3
    spec = dict(model="qif_sfa", n=n, T=T, dt=dt, S=2, cutoff=10, grad=True, truncate_steps=300,
                W=rng.standard_normal((n, n)) * 2.0 / np.sqrt(n),
                params=dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.5, tau_x=1.5, k=2.0),
                train_params=["weights", "eta"], train_in=False, train_out=True,
                w_in=rng.standard_normal((n, m)), w_out=rng.standard_normal((k, n)) / np.sqrt(n),
                inputs=sin_inputs(rng, T, m, dt, amp=10.0, offset=16.0))
    spec["targets"] = rng.standard_normal((len([s for s in range(T) if s >= 10 and s % 2 == 0]), k))
    save_case(ref, "qif_sfa_bptt_trunc", spec)

    # ---- G6: LIF (documentation/bptt_spiking_neurons_recurrent.py set-up, N=10) -------------------
    n, T, m, k, dt = 10, 2000, 2, 3, 5e-3
    t = np.arange(T) * dt
    inp = np.stack([np.sin(t * 2 * np.pi * w) * 40.0 for w in (0.3, 0.5)], axis=1)
    spec = dict(model="lif", n=n, T=T, dt=dt, S=1, cutoff=0, grad=True,
                W=rng.standard_normal((n, n)),
                params=dict(eta=10.0, tau=rng.uniform(10.0, 20.0, n), tau_s=5.0, k=2.0),
                train_params=["weights"], train_in=False, train_out=True,
                w_in=rng.standard_normal((n, m)), w_out=rng.standard_normal((k, n)),
                spike_kwargs=dict(spike_threshold=10.0, spike_reset=-10.0), inputs=inp)
    spec["targets"] = rng.standard_normal((T, k))
    save_case(ref, "lif_bptt", spec)

    # ---- G6b: Izhikevich neurons (spiking_neurons/ik.yaml:8-30; regime of documentation/models/ik.py), BPTT -----------
    n, T, m, k, dt = 20, 1200, 2, 2, 1e-1          # ms time scale: C=100, tau_u=33, spikes at 40 mV, reset to -60 mV
    rng_main, rng = rng, np.random.default_rng(777)    # own stream: the case was added after the others were minted
    spec = dict(model="ik", n=n, T=T, dt=dt, S=2, cutoff=0, grad=True,
                W=np.abs(rng.standard_normal((n, n))) * 4.0 / n,
                params=dict(eta=rng.uniform(60.0, 160.0, n), g=1.5, kappa=10.0, tau_s=6.0, E_r=0.0, b=-2.0, tau_u=33.33, k=0.7, C=100.0),
                train_params=["weights", "eta", "g", "kappa", "tau_s", "b", "tau_u", "C", "k", "E_r"], train_in=True, train_out=True,
                w_in=rng.standard_normal((n, m)) * 10.0, w_out=rng.standard_normal((k, n)) / np.sqrt(n),
                inputs=sin_inputs(rng, T, m, dt * 1e-2, amp=3.0, offset=1.0),
                spike_kwargs=dict(spike_threshold=40.0, spike_reset=-60.0), record_vars=[("v", False), ("u", True), ("s", False)])
    spec["targets"] = rng.standard_normal((len(range(0, T, 2)), k))
    if only in (None, "ik_bptt"):
        save_case(ref, "ik_bptt", spec)
    # ---- G6c: Izhikevich neurons with a global recovery variable (ik.yaml:33-39, iku_op: mean(v), mean(spike)), BPTT ---------
    rng = np.random.default_rng(778)
    n, T, m, k, dt = 16, 1000, 2, 2, 1e-1
    spec = dict(model="iku", n=n, T=T, dt=dt, S=2, cutoff=0, grad=True,
                W=np.abs(rng.standard_normal((n, n))) * 4.0 / n,
                params=dict(eta=rng.uniform(60.0, 160.0, n), g=1.5, kappa=rng.uniform(5.0, 15.0, n), tau_s=6.0, E_r=0.0, b=rng.uniform(-3.0, -1.0, n),
                            tau_u=33.33, k=0.7, C=100.0),
                train_params=["weights", "eta", "g", "kappa", "tau_s", "b", "tau_u", "C", "k", "E_r"], train_in=True, train_out=True,
                w_in=rng.standard_normal((n, m)) * 10.0, w_out=rng.standard_normal((k, n)) / np.sqrt(n),
                inputs=sin_inputs(rng, T, m, dt * 1e-2, amp=3.0, offset=1.0),
                spike_kwargs=dict(spike_threshold=40.0, spike_reset=-60.0), record_vars=[("v", False), ("u", True), ("s", False)])
    spec["targets"] = rng.standard_normal((len(range(0, T, 2)), k))
    if only in (None, "iku_bptt"):
        save_case(ref, "iku_bptt", spec)
    # ---- G6d: Izhikevich neurons with a bi-exponential synapse (ik.yaml:42-70, ik_biexp_op: 4 state variables v, u, s, x), BPTT ----
    rng = np.random.default_rng(779)
    n, T, m, k, dt = 16, 1000, 2, 2, 1e-1
    spec = dict(model="ik_biexp", n=n, T=T, dt=dt, S=2, cutoff=0, grad=True,
                W=np.abs(rng.standard_normal((n, n))) * 4.0 / n,
                params=dict(eta=rng.uniform(60.0, 160.0, n), g=1.5, kappa=rng.uniform(5.0, 15.0, n), tau_d=rng.uniform(5.0, 7.0, n), tau_r=2.0,
                            E_r=0.0, b=rng.uniform(-3.0, -1.0, n), tau_u=33.33, k=0.7, C=100.0),
                train_params=["weights", "eta", "g", "kappa", "tau_d", "tau_r", "b", "tau_u", "C", "k", "E_r"], train_in=True, train_out=True,
                w_in=rng.standard_normal((n, m)) * 10.0, w_out=rng.standard_normal((k, n)) / np.sqrt(n),
                inputs=sin_inputs(rng, T, m, dt * 1e-2, amp=3.0, offset=1.0),
                spike_kwargs=dict(spike_threshold=40.0, spike_reset=-60.0),
                record_vars=[("v", False), ("u", True), ("s", False), ("x", False)])
    spec["targets"] = rng.standard_normal((len(range(0, T, 2)), k))
    if only in (None, "ik_biexp_bptt"):
        save_case(ref, "ik_biexp_bptt", spec)
    if only in (None, "two_node_chain"):
        save_two_node_chain(ref)
    if only in (None, "feedback_net"):
        save_feedback_net(ref)
    if only in (None, "edges_square"):
        save_edges_square(ref)
    if only in (None, "multispike_ei"):
        save_multispike(ref)
    rng = rng_main

    # ---- G7: output activation + masked input edge (LinearMasked, edges.py:150-174) ---------------
    n, T, m, k, dt = 12, 200, 4, 3, 2e-2
    spec = dict(model="li_tanh", n=n, T=T, dt=dt, S=2, cutoff=0, grad=True,
                W=rng.standard_normal((n, n)) / np.sqrt(n),
                params=dict(tau=1.5, k=1.0, eta=0.0), train_params=["weights"], train_in=True, train_out=True,
                w_in=rng.standard_normal((n, m)), in_mask=(rng.uniform(size=(n, m)) < 0.6).astype(np.float64),
                w_out=rng.standard_normal((k, n)), out_act="softmax", in_act="sigmoid",
                inputs=sin_inputs(rng, T, m, dt, amp=1.5))
    spec["targets"] = rng.uniform(size=(len(range(0, T, 2)), k))
    save_case(ref, "li_tanh_masked_softmax", spec)

    # ---- G8: edges known-answers: Linear vs torch.nn.Linear convention; RLS.update sequence --------
    torch.manual_seed(3)
    n_in, n_out, steps = 9, 4, 25
    lin = ref.edges.Linear(n_in, n_out, weights=rng.standard_normal((n_out, n_in)), dtype=torch.float64)
    xs = rng.standard_normal((steps, n_in))
    ys = rng.standard_normal((steps, n_out))
    lin_out = np.stack([lin.forward(torch.tensor(x)).numpy() for x in xs])
    rls = ref.edges.RLS(n_in, n_out, dtype=torch.float64, beta=0.98, alpha=2.0)
    w_hist, p_hist, loss_hist = [], [], []
    for x, y in zip(xs, ys):
        xt, yt = torch.tensor(x), torch.tensor(y)
        y_hat = rls.forward(xt)
        rls.update(xt, yt, y_hat)
        w_hist.append(rls.weights.numpy().copy())
        p_hist.append(rls.P.numpy().copy())
        loss_hist.append(float(rls.loss))
    np.savez_compressed(os.path.join(OUT, "edges.npz"), lin_w=lin.weights.numpy(), xs=xs, ys=ys, lin_out=lin_out,
                        rls_w=np.stack(w_hist), rls_p=np.stack(p_hist), rls_loss=np.asarray(loss_hist),
                        rls_beta=0.98, rls_alpha=2.0)
    print("edges: ok")

    # ---- G8b: stateful edges (edges.py:68-147): delay buffer, linear filter, both -- outputs of the reference classes --------
    rng_e = np.random.default_rng(99)
    n_in, n_out, Ts = 4, 3, 40
    delays = np.asarray([2, 0, 1, 2])
    filt = rng_e.standard_normal((n_in, n_in)) * 0.3
    w = rng_e.standard_normal((n_out, n_in))
    xs = rng_e.standard_normal((Ts, n_in))
    blob = dict(delays=delays, filter=filt, w=w, xs=xs)
    mem = ref.edges.LinearMemory(n_in, n_out, delays=delays, weights=w, dtype=torch.float64)
    flt = ref.edges.LinearFilter(n_in, n_out, filter_weights=filt, weights=w, dtype=torch.float64)
    mf = ref.edges.LinearMemoryFilter(n_in, n_out, delays=delays, filter_weights=filt, weights=w, dtype=torch.float64)
    for name, e in (("memory", mem), ("filter", flt), ("memory_filter", mf)):
        blob[f"{name}_out"] = np.stack([e.forward(torch.tensor(x)).detach().numpy().copy() for x in xs])
    blob["memory_buffer"] = mem.buffer.numpy(); blob["filter_y"] = flt.y.numpy(); blob["memory_filter_buffer"] = mf.buffer.numpy()
    # and inside a reference Network: delayed input edge -> LI-tanh node -> filtered readout edge
    n, m, k, T, dt = 9, 3, 3, 150, 2e-2
    W = rng_e.standard_normal((n, n)) / np.sqrt(n)
    w_in, w_out = rng_e.standard_normal((n, m)), rng_e.standard_normal((k, n))
    d_in = np.asarray([1, 3, 0]); f_out = rng_e.standard_normal((n, n)) * 0.2 / np.sqrt(n)
    inputs = sin_inputs(rng_e, T, m, dt, amp=1.5)
    func, args, var_map, param_map = orc.build_node_args("li_tanh", n, W, dict(tau=1.5, k=1.0, eta=0.1), torch.float64, "I_ext", None)
    net = ref.Network(dt, device="cpu", dtype=torch.float64)
    node = ref.nodes.RateNet(func, args, var_map, param_map, dt=dt, dtype=torch.float64, device="cpu")
    net.add_node("rnn", node, node_type="diff_eq")
    net.add_func_node("inp", m, "identity"); net.add_func_node("out", k, "identity")
    net.add_edge("inp", "rnn", weights=w_in, delays=d_in)
    net.add_edge("rnn", "out", weights=w_out, filter_weights=f_out)
    obs = net.run(torch.tensor(inputs), sampling_steps=2, cutoff=3, verbose=False, enable_grad=False)
    blob.update(net_W=W, net_w_in=w_in, net_w_out=w_out, net_d_in=d_in, net_f_out=f_out, net_inputs=inputs,
                net_out=obs.to_numpy("out"), net_meta=np.asarray(repr(dict(n=n, m=m, k=k, T=T, dt=dt, S=2, cutoff=3, tau=1.5, k_c=1.0, eta=0.1))))
    np.savez_compressed(os.path.join(OUT, "edges_stateful.npz"), **blob)
    print("edges_stateful: ok", blob["memory_out"][:4, 0])

    # ---- G9: fit_ridge on a small reservoir (network.py:709-784) ----------------------------------
    n, T, m, k, dt = 15, 400, 3, 2, 1e-2
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    w_in = rng.standard_normal((n, m))
    inputs = sin_inputs(rng, T, m, dt, amp=1.0)
    targets = rng.standard_normal((T, k))   # fit_ridge demands len(inputs)==len(targets) (network.py:744-746) => S=1
    blob = dict(W=W, w_in=w_in, inputs=inputs, targets=targets, dt=dt, alpha=1e-3, S=1)
    net, node = ref_network(ref, "li_tanh", n, W, dt, torch.float64, params=dict(tau=1.0), w_in=w_in)
    obs = net.fit_ridge(torch.tensor(inputs).numpy(), targets, sampling_steps=1, alpha=1e-3, verbose=False,
                        add_readout_node=False)
    blob["X"] = obs.to_numpy("out")
    blob["w_out"] = obs["w_out"].detach().numpy()
    blob["y"] = obs["y"].detach().numpy()
    np.savez_compressed(os.path.join(OUT, "ridge.npz"), **blob)
    print("ridge: ok")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else None)
