"""Helpers shared by the parity tests: load a golden case and rebuild it on the oracle / on the engine."""
from __future__ import annotations

import ast
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))

import rectipy_oracle as orc  # noqa: E402

TD = {"float64": torch.float64, "float32": torch.float32}
RUN_CASES = ["li_tanh_bptt", "li_sigmoid_fwd", "qif_bptt", "qif_sfa_fwd", "qif_sfa_bptt_trunc", "lif_bptt", "ik_bptt", "iku_bptt", "ik_biexp_bptt",
             "li_tanh_masked_softmax"]


class Case:
    def __init__(self, name):
        self.name = name
        z = np.load(os.path.join(GOLDEN, f"{name}.npz"))
        self.z = z
        self.meta = ast.literal_eval(str(z["meta"]))
        self.params = {k[len("param_"):]: z[k] for k in z.files if k.startswith("param_")}
        self.inp = {k[len("in_"):]: z[k] for k in z.files if k.startswith("in_")}

    def ref(self, dtype_name, key):
        return self.z[f"{dtype_name}_{key}"]

    def has(self, dtype_name, key):
        return f"{dtype_name}_{key}" in self.z.files

    def params_py(self):
        return {k: (float(v) if v.ndim == 0 else v) for k, v in self.params.items()}


def oracle_net(case: Case, dtype_name: str) -> orc.OracleNet:
    dtype = TD[dtype_name]
    m = case.meta
    node = orc.make_node(m["model"], m["n"], case.inp["W"], m["dt"], params=case.params_py(), dtype=dtype,
                         train_params=m.get("train_params"), **(m.get("spike_kwargs") or {}))
    w_in = w_out = mask = None
    if "w_in" in case.inp:
        w_in = torch.tensor(case.inp["w_in"], dtype=dtype, requires_grad=bool(m.get("train_in")))
    if "w_out" in case.inp:
        w_out = torch.tensor(case.inp["w_out"], dtype=dtype, requires_grad=bool(m.get("train_out")))
    if "in_mask" in case.inp:
        mask = torch.tensor(case.inp["in_mask"], dtype=dtype)
    return orc.OracleNet(node, w_in=w_in, w_out=w_out, in_act=m.get("in_act"), out_act=m.get("out_act"),
                         w_in_mask=mask)


def oracle_run_case(case: Case, dtype_name: str):
    """Re-run a golden case on the oracle port; returns dict with the same keys as the golden blob."""
    dtype = TD[dtype_name]
    m = case.meta
    net = oracle_net(case, dtype_name)
    inputs = torch.tensor(case.inp["inputs"], dtype=dtype)
    rec = [(v, red) for v, red in m.get("record_vars", [])]
    res = {}
    if m.get("grad"):
        targets = torch.tensor(case.inp["targets"], dtype=dtype)
        run = None
        # bptt_grads does run + mse + backward; re-run pieces here to also collect recorded vars
        r = net.run(inputs, sampling_steps=m.get("S", 1), cutoff=m.get("cutoff", 0), record_vars=rec,
                    enable_grad=True, truncate_steps=m.get("truncate_steps"))
        pred = torch.stack(r["out"])
        loss = torch.nn.functional.mse_loss(pred, targets)
        loss.backward()
        res["loss"] = loss.detach().numpy()
        for name in m.get("train_params") or []:
            res[f"grad_{name}"] = net.node.get(name).grad.numpy()
        if m.get("train_in"):
            res["grad_w_in"] = net.w_in.grad.numpy()
        if m.get("train_out"):
            res["grad_w_out"] = net.w_out.grad.numpy()
    else:
        r = net.run(inputs, sampling_steps=m.get("S", 1), cutoff=m.get("cutoff", 0), record_vars=rec,
                    enable_grad=False, truncate_steps=m.get("truncate_steps"))
    res["out"] = torch.stack(r["out"]).detach().numpy()
    res["steps"] = np.asarray(r["steps"])
    for v, red in rec:
        res[f"var_{v}"] = torch.stack([x.detach() for x in r["vars"][v]]).numpy()
    node = net.node
    res["y_final"] = (node.full_state() if node.spiking else node.y).detach().numpy()
    return res


def rel_err(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.max(np.abs(a - b)) / (np.max(np.abs(b)) + 1e-300))


# ---- the same cases on the engine (rectipy_b200 public API) -------------------------------------------------
TEMPLATE_PATH = {
    "li_tanh": ("neuron_model_templates.rate_neurons.leaky_integrator.tanh", "li_op", "tanh_op/r", "li_op/r_in"),
    "li_sigmoid": ("neuron_model_templates.rate_neurons.leaky_integrator.sigmoid", "li_op", "sigmoid_op/r", "li_op/r_in"),
    "qif": ("neuron_model_templates.spiking_neurons.qif.qif", "qif_op", "s", "s_in"),
    "qif_sfa": ("neuron_model_templates.spiking_neurons.qif.qif_sfa", "qif_sfa_op", "s", "s_in"),
    "lif": ("neuron_model_templates.spiking_neurons.lif.lif", "lif_op", "s", "s_in"),
    "ik": ("neuron_model_templates.spiking_neurons.ik.ik", "ik_op", "s", "s_in"),
    "iku": ("neuron_model_templates.spiking_neurons.ik.iku", "iku_op", "s", "s_in"),
    "ik_biexp": ("neuron_model_templates.spiking_neurons.ik.ik_biexp", "ik_biexp_op", "s", "s_in"),
}


def engine_net(case: "Case", device="cuda:0", batch=1, precision="auto", params_override=None):
    """Build the golden case through the drop-in API (add_diffeq_node / add_func_node / add_edge)."""
    import rectipy_b200 as rp
    m = case.meta
    model = m["model"]
    path, op, svar, tvar = TEMPLATE_PATH[model]
    net = rp.Network(m["dt"], device=device, batch=batch, precision=precision)
    params = case.params_py()
    if params_override:
        params.update(params_override)
    sigmoid_keys = {"r_max", "s", "v0"}
    node_vars = {(f"sigmoid_op/{k}" if (model == "li_sigmoid" and k in sigmoid_keys) else f"{op}/{k}"): v
                 for k, v in params.items()}
    kw = dict(weights=case.inp["W"], source_var=svar, target_var=tvar, input_var=f"{op}/I_ext",
              node_vars=node_vars, train_params=[("weights" if p == "weights" else f"{op}/{p}") for p in (m.get("train_params") or [])])
    if model in orc.SPIKING:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **(m.get("spike_kwargs") or {}))
    else:
        kw.update(output_var=f"{op}/v")
    node = net.add_diffeq_node("rnn", path, **kw)
    if "w_in" in case.inp:
        net.add_func_node("inp", case.inp["w_in"].shape[1], m.get("in_act", "identity"))
        ekw = {"mask": case.inp["in_mask"]} if "in_mask" in case.inp else {}
        net.add_edge("inp", "rnn", weights=case.inp["w_in"], train="gd" if m.get("train_in") else None, **ekw)
    if "w_out" in case.inp:
        net.add_func_node("out", case.inp["w_out"].shape[0], m.get("out_act", "identity"))
        net.add_edge("rnn", "out", weights=case.inp["w_out"], train="gd" if m.get("train_out") else None)
    return net, node


def engine_run_case(case: "Case", device="cuda:0", precision="auto"):
    """Run a golden case on the engine; same result keys as the golden blob."""
    m = case.meta
    net, node = engine_net(case, device=device, precision=precision)
    rec = [("rnn", v, red) for v, red in m.get("record_vars", [])]
    kw = {}
    if "truncate_steps" in m:
        kw["truncate_steps"] = m["truncate_steps"]
    obs = net.run(case.inp["inputs"], sampling_steps=m.get("S", 1), cutoff=m.get("cutoff", 0), verbose=False,
                  enable_grad=bool(m.get("grad")), record_vars=rec, **kw)
    res = {"out": obs.to_numpy("out"), "steps": np.asarray(obs["steps"])}
    for v, red in m.get("record_vars", []):
        res[f"var_{v}"] = obs.to_numpy(("rnn", v))
    res["y_final"] = node.y.detach().cpu().numpy()
    if m.get("grad"):
        target = torch.tensor(case.inp["targets"], dtype=torch.float32, device=device)
        loss = torch.nn.MSELoss()(torch.stack(obs["out"]), target)
        loss.backward()
        res["loss"] = loss.detach().cpu().numpy()
        for name in m.get("train_params") or []:
            res[f"grad_{name}"] = node[name].grad.detach().cpu().numpy()
        if m.get("train_in"):
            res["grad_w_in"] = net.get_edge("inp", "rnn").weights.grad.detach().cpu().numpy()
        if m.get("train_out"):
            res["grad_w_out"] = net.get_edge("rnn", "out").weights.grad.detach().cpu().numpy()
    return res
