"""CPU oracle for the RectiPy time-stepping hot path.  TEST INFRASTRUCTURE ONLY.

This file is the checker, never the product: only ``tests/``, ``__graft_entry__.smoke()`` and the
``cpu_baseline`` / ``--impl reference`` legs of ``bench.py`` may import it.  The shipped package
``rectipy_b200`` never imports anything under ``oracle/``.

It restates, in plain torch-on-CPU, what the reference computes along the path

    Network.run / fit_bptt  ->  Network.forward  ->  Linear.forward  ->  RateNet/SpikeResetNet.forward
                            ->  (PyRates-generated) vector field  ->  Spike (heaviside + surrogate)

Every function cites the reference lines it follows (paths relative to /root/reference).

Parity pin
----------
* Everything except the vector-field arithmetic is pinned against the *unmodified* reference classes
  (imported through ``oracle/ref_shim.py`` in the build container) by the golden vectors under
  ``tests/golden/`` (generator: ``oracle/make_golden.py``).
* The vector-field arithmetic itself is produced at run time by the third-party package PyRates, which
  the reference does not pin (requirements.txt:2) and which is not installable offline.  The fields
  below restate the published YAML equations (neuron_model_templates/**.yaml).  For that part:
  "parity unpinned" by any reference-owned golden vector; the YAML text is the ground truth.

The reference has no trial/batch axis (nodes.py:90, network.py:549).  ``run_trials`` simply loops the
unbatched path over independent trials, which is what a reference user would have to do.
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch

# --------------------------------------------------------------------------------------------------
# vector fields:  f(t, y, *args) -> dy      (operator boundary: nodes.py:58,169,388)
# --------------------------------------------------------------------------------------------------
# Each field returns (func, arg_names).  ``args`` handed to the node is ``[y0] + [value of each name]``.
# State order inside y follows the order of the differential equations in the YAML operator.


def field_li(activation: str = "tanh") -> Tuple[Callable, List[str]]:
    """li_op + tanh_op / sigmoid_op   (rate_neurons/leaky_integrator.yaml:8-36).

    v' = -v/tau + k*r_in + I_ext + eta ,  r_in = W @ r ,  r = tanh(v) | r_max/(1+exp(s*(v0-v)))
    (hand-written precedent in the reference's own test: rectipy_tests/test_nodes.py:32-33)
    """
    if activation == "tanh":
        names = ["weights", "tau", "k", "eta", "I_ext"]

        def f(t, y, weights, tau, k, eta, I_ext):
            r = torch.tanh(y)
            return -y / tau + k * (weights @ r) + I_ext + eta
        return f, names
    if activation == "sigmoid":
        names = ["weights", "tau", "k", "eta", "I_ext", "r_max", "s", "v0"]

        def f(t, y, weights, tau, k, eta, I_ext, r_max, s, v0):
            r = r_max / (1.0 + torch.exp(s * (v0 - y)))
            return -y / tau + k * (weights @ r) + I_ext + eta
        return f, names
    raise ValueError(activation)


def field_qif(n: int) -> Tuple[Callable, List[str]]:
    """qif_op   (spiking_neurons/qif.yaml:8-22):  v' = (v^2+eta+I_ext)/tau + k*s_in ;  s' = -s/tau_s + spike."""
    names = ["weights", "tau", "k", "tau_s", "eta", "I_ext", "spike"]

    def f(t, y, weights, tau, k, tau_s, eta, I_ext, spike):
        v, s = y[:n], y[n:2 * n]
        dv = (v * v + eta + I_ext) / tau + k * (weights @ s)
        ds = -s / tau_s + spike
        return torch.cat((dv, ds), 0)
    return f, names


def field_qif_sfa(n: int) -> Tuple[Callable, List[str]]:
    """qif_sfa_op (qif.yaml:25-35): qif_op with eta -> eta - x and  x' = -x/tau_x + alpha*spike."""
    names = ["weights", "tau", "k", "tau_s", "eta", "I_ext", "spike", "alpha", "tau_x"]

    def f(t, y, weights, tau, k, tau_s, eta, I_ext, spike, alpha, tau_x):
        v, s, x = y[:n], y[n:2 * n], y[2 * n:3 * n]
        dv = (v * v + eta - x + I_ext) / tau + k * (weights @ s)
        ds = -s / tau_s + spike
        dx = -x / tau_x + alpha * spike
        return torch.cat((dv, ds, dx), 0)
    return f, names


def field_lif(n: int) -> Tuple[Callable, List[str]]:
    """lif_op (spiking_neurons/lif.yaml:10-23):  v' = -v/tau + k*s_in + I_ext + eta ;  s' = -s/tau_s + spike + s_ext."""
    names = ["weights", "tau", "k", "tau_s", "eta", "I_ext", "spike", "s_ext"]

    def f(t, y, weights, tau, k, tau_s, eta, I_ext, spike, s_ext):
        v, s = y[:n], y[n:2 * n]
        dv = -v / tau + k * (weights @ s) + I_ext + eta
        ds = -s / tau_s + spike + s_ext
        return torch.cat((dv, ds), 0)
    return f, names


def field_ik(n: int) -> Tuple[Callable, List[str]]:
    """ik_op (spiking_neurons/ik.yaml:8-30), Izhikevich neuron with a conductance synapse:
    v' = (k*(v-v_r)*(v-v_theta) - u + I_ext + eta + g*s_in*(E_r - v)) / C ;  u' = (b*(v-v_r) - u)/tau_u + kappa*spike ;
    s' = -s/tau_s + spike.   State order = equation order: v, u, s."""
    names = ["weights", "C", "k", "v_r", "v_theta", "eta", "g", "E_r", "b", "tau_u", "kappa", "tau_s", "I_ext", "spike"]

    def f(t, y, weights, C, k, v_r, v_theta, eta, g, E_r, b, tau_u, kappa, tau_s, I_ext, spike):
        v, u, s = y[:n], y[n:2 * n], y[2 * n:3 * n]
        dv = (k * (v - v_r) * (v - v_theta) - u + I_ext + eta + g * (weights @ s) * (E_r - v)) / C
        du = (b * (v - v_r) - u) / tau_u + kappa * spike
        ds = -s / tau_s + spike
        return torch.cat((dv, du, ds), 0)
    return f, names


def field_iku(n: int) -> Tuple[Callable, List[str]]:
    """iku_op (spiking_neurons/ik.yaml:33-39): ik_op whose recovery variable is driven by the population means,
    u' = (b*(mean(v)-v_r) - u)/tau_u + kappa*mean(spike)  (`equations: replace` of ik_op).  State order v, u, s."""
    names = ["weights", "C", "k", "v_r", "v_theta", "eta", "g", "E_r", "b", "tau_u", "kappa", "tau_s", "I_ext", "spike"]

    def f(t, y, weights, C, k, v_r, v_theta, eta, g, E_r, b, tau_u, kappa, tau_s, I_ext, spike):
        v, u, s = y[:n], y[n:2 * n], y[2 * n:3 * n]
        dv = (k * (v - v_r) * (v - v_theta) - u + I_ext + eta + g * (weights @ s) * (E_r - v)) / C
        du = (b * (torch.mean(v) - v_r) - u) / tau_u + kappa * torch.mean(spike)
        ds = -s / tau_s + spike
        return torch.cat((dv, du, ds), 0)
    return f, names


def field_ik_biexp(n: int) -> Tuple[Callable, List[str]]:
    """ik_biexp_op (spiking_neurons/ik.yaml:42-70): iku_op with a bi-exponential synapse,
    s' = -s/tau_d + x ;  x' = -x/tau_r + spike.   State order = equation order: v, u, s, x."""
    names = ["weights", "C", "k", "v_r", "v_theta", "eta", "g", "E_r", "b", "tau_u", "kappa", "tau_d", "tau_r", "I_ext", "spike"]

    def f(t, y, weights, C, k, v_r, v_theta, eta, g, E_r, b, tau_u, kappa, tau_d, tau_r, I_ext, spike):
        v, u, s, x = y[:n], y[n:2 * n], y[2 * n:3 * n], y[3 * n:4 * n]
        dv = (k * (v - v_r) * (v - v_theta) - u + I_ext + eta + g * (weights @ s) * (E_r - v)) / C
        du = (b * (torch.mean(v) - v_r) - u) / tau_u + kappa * torch.mean(spike)
        ds = -s / tau_d + x
        dx = -x / tau_r + spike
        return torch.cat((dv, du, ds, dx), 0)
    return f, names


#: template defaults (leaky_integrator.yaml:12-17,24-28; qif.yaml:13-22,33-35; lif.yaml:17-23)
DEFAULTS = {
    "li_tanh":    dict(tau=10.0, k=1.0, eta=0.0),
    "li_sigmoid": dict(tau=10.0, k=1.0, eta=0.0, r_max=1.0, s=1.0, v0=0.0),
    "qif":        dict(tau=1.0, k=1.0, tau_s=1.0, eta=-5.0),
    "qif_sfa":    dict(tau=1.0, k=1.0, tau_s=1.0, eta=-5.0, alpha=1.0, tau_x=10.0),
    "lif":        dict(tau=10.0, k=1.0, tau_s=0.5, eta=0.0),
    "ik":         dict(C=100.0, k=0.7, v_r=-60.0, v_theta=-40.0, eta=0.0, g=1.0, E_r=0.0, b=-2.0, tau_u=33.33, kappa=10.0, tau_s=6.0),
    "iku":        dict(C=100.0, k=0.7, v_r=-60.0, v_theta=-40.0, eta=0.0, g=1.0, E_r=0.0, b=-2.0, tau_u=33.33, kappa=10.0, tau_s=6.0),
    "ik_biexp":   dict(C=100.0, k=0.7, v_r=-60.0, v_theta=-40.0, eta=0.0, g=1.0, E_r=0.0, b=-2.0, tau_u=33.33, kappa=10.0, tau_d=6.0, tau_r=2.0),
}
#: initial values of the state variables, in state order
INIT = {
    "li_tanh": [("v", 0.0)], "li_sigmoid": [("v", 0.0)],
    "qif": [("v", -2.0), ("s", 0.0)], "qif_sfa": [("v", -2.0), ("s", 0.0), ("x", 0.0)],
    "lif": [("v", 0.0), ("s", 0.0)],
    "ik": [("v", -60.0), ("u", 0.0), ("s", 0.0)],
    "iku": [("v", -60.0), ("u", 0.0), ("s", 0.0)],
    "ik_biexp": [("v", -60.0), ("u", 0.0), ("s", 0.0), ("x", 0.0)],
}
SPIKING = {"qif", "qif_sfa", "lif", "ik", "iku", "ik_biexp"}


def build_field(model: str, n: int):
    if model == "li_tanh":
        return field_li("tanh")
    if model == "li_sigmoid":
        return field_li("sigmoid")
    if model == "qif":
        return field_qif(n)
    if model == "qif_sfa":
        return field_qif_sfa(n)
    if model == "lif":
        return field_lif(n)
    if model == "ik":
        return field_ik(n)
    if model == "iku":
        return field_iku(n)
    if model == "ik_biexp":
        return field_ik_biexp(n)
    raise ValueError(model)


def build_node_args(model: str, n: int, weights, params: Optional[dict] = None, dtype=torch.float64,
                    input_var: str = "I_ext", output_var: Optional[str] = None, y0=None):
    """Assemble (func, args, var_map, param_map) with the contract of nodes.py:58-90,146-159.

    ``args[0]`` is the initial state, ``args[1:]`` the parameter list; ``param_map`` maps names to
    indices in ``args[1:]`` ("in" -> input slot, "weights" -> recurrent matrix, "spike_var" -> spike slot);
    ``var_map`` maps names to (start, stop) slices of y ("out", and "reset_var" for spiking nodes).
    """
    func, names = build_field(model, n)
    p = dict(DEFAULTS[model])
    if params:
        p.update(params)
    svars = INIT[model]
    if y0 is None:
        y0 = torch.cat([torch.full((n,), val, dtype=dtype) for _, val in svars])
    else:
        y0 = torch.as_tensor(y0, dtype=dtype).clone()
    args = [y0]
    for name in names:
        if name == "weights":
            args.append(torch.as_tensor(weights, dtype=dtype).clone())
        elif name in ("I_ext", "spike", "s_ext"):
            args.append(torch.zeros(n, dtype=dtype))
        else:
            args.append(torch.as_tensor(p[name], dtype=dtype).clone().reshape(-1)
                        if np.ndim(p[name]) > 0 else torch.tensor([float(p[name])], dtype=dtype))
    param_map = {name: i for i, name in enumerate(names)}
    param_map["in"] = param_map[input_var]
    if model in SPIKING:
        param_map["spike_var"] = param_map["spike"]
    var_map = {name: (i * n, (i + 1) * n) for i, (name, _) in enumerate(svars)}
    if output_var is None:
        output_var = "s" if model in SPIKING else "v"
    var_map["out"] = var_map[output_var]
    if model in SPIKING:
        var_map["reset_var"] = var_map["v"]
    return func, args, var_map, param_map


# --------------------------------------------------------------------------------------------------
# Spike  (nodes.py:468-481)
# --------------------------------------------------------------------------------------------------

class OracleSpike(torch.autograd.Function):
    """heaviside forward (value at 0 = ``center``), 1/(1+slope*|x|)^2 surrogate backward."""

    @staticmethod
    def forward(ctx, x, slope, center):
        ctx.save_for_backward(x, slope)
        return torch.heaviside(x, center)

    @staticmethod
    def backward(ctx, g):
        x, slope = ctx.saved_tensors
        return g / (1.0 + slope * torch.abs(x)) ** 2, None, None


# --------------------------------------------------------------------------------------------------
# nodes
# --------------------------------------------------------------------------------------------------

class OracleRateNode:
    """RateNet (nodes.py:54-211): y <- y + dt*f(0,y,*args); forward returns the PRE-update out slice."""

    spiking = False

    def __init__(self, func, args, var_map, param_map, dt, dtype=torch.float64, train_params=None):
        self.args = [a for a in args[1:]]
        self.var_map, self.param_map = var_map, param_map
        self.start, self.stop = var_map["out"]
        self.inp = param_map["in"]
        self.dt, self.func, self.dtype = dt, func, dtype
        for a in self.args:
            a.requires_grad = False                                       # nodes.py:80-82
        self.train_idx = [param_map[p] for p in (train_params or [])]      # nodes.py:83-86
        for i in self.train_idx:
            self.args[i].requires_grad = True
        self.y = args[0].detach().clone().to(dtype)
        self.y.requires_grad = len(self.train_idx) > 0                     # nodes.py:89-90
        in_arg = self.args[self.inp]
        self.n_in = int(in_arg.shape[0])
        self.n_out = self.stop - self.start

    def parameters(self):
        return [self.args[i] for i in self.train_idx]

    def get(self, name):
        if name in self.param_map:
            return self.args[self.param_map[name]]
        a, b = self.var_map[name]
        return self.y[a:b]

    def forward(self, x):                                                  # nodes.py:166-170
        self.args[self.inp] = x
        y_old = self.y
        self.y = y_old + self.dt * self.func(0, y_old, *self.args)
        return y_old[self.start:self.stop]

    def detach(self):                                                      # nodes.py:176-196 (requires_grad=True)
        self.y = self.y.detach()
        self.y.requires_grad = True

    def reset(self, y=None):                                               # nodes.py:198-211
        req = self.y.requires_grad
        y = torch.zeros_like(self.y) if y is None else torch.as_tensor(y, dtype=self.dtype).clone().detach()
        y.requires_grad = req
        self.y = y


class OracleSpikeResetNode(OracleRateNode):
    """SpikeResetNet (nodes.py:333-401): threshold -> spike -> Euler -> reset blend."""

    spiking = True

    def __init__(self, func, args, var_map, param_map, dt, dtype=torch.float64, train_params=None,
                 spike_threshold=1e2, spike_reset=-1e2, spike_slope=None, spike_center=1.0):
        super().__init__(func, args, var_map, param_map, dt, dtype, train_params)
        if spike_slope is None:
            spike_slope = 100.0 / abs(spike_threshold - spike_reset)       # nodes.py:346
        self.slope = torch.tensor(spike_slope, dtype=dtype)
        self.center = torch.tensor(spike_center, dtype=dtype)
        self.spike_idx = param_map["spike_var"]
        self.v_reset = torch.tensor(spike_reset, dtype=dtype)
        self.thresh = torch.tensor(spike_threshold, dtype=dtype)
        self.r0, self.r1 = var_map["reset_var"]
        self._split()

    def _split(self):                                                      # nodes.py:398-401
        self.y_a = self.y[:self.r0].clone()
        self.y_v = self.y[self.r0:self.r1].clone()
        self.y_b = self.y[self.r1:].clone()

    def forward(self, x):                                                  # nodes.py:382-392
        spikes = OracleSpike.apply(self.y_v - self.thresh, self.slope, self.center)
        reset = spikes.detach()
        self.args[self.spike_idx] = spikes / self.dt
        self.args[self.inp] = x
        self.y = torch.cat((self.y_a, self.y_v, self.y_b), 0)
        y_new = self.y + self.dt * self.func(0, self.y, *self.args)
        self.y_a = y_new[:self.r0]
        self.y_v = y_new[self.r0:self.r1] * (1.0 - reset) + reset * self.v_reset
        self.y_b = y_new[self.r1:]
        return self.y[self.start:self.stop]

    def detach(self):
        self.y_a, self.y_v, self.y_b = (t.detach() for t in (self.y_a, self.y_v, self.y_b))
        for t in (self.y_a, self.y_v, self.y_b):
            t.requires_grad = True

    def reset(self, y=None):                                               # nodes.py:394-396
        super().reset(y)
        self._split()

    def full_state(self):
        """State *after* the last step (the reference leaves ``y`` one step behind, nodes.py:387)."""
        return torch.cat((self.y_a, self.y_v, self.y_b), 0)


class OracleMultiSpikeResetNode(OracleRateNode):
    """MultiSpikeResetNet (nodes.py:404-465): several (spike variable, reset variable) pairs `spike_var_i` / `spike_reset_i`,
    one shared threshold and reset value.  Differences from SpikeResetNet that the restatement keeps: the spikes of a step are
    taken from the buffers `_y_reset[i]`, which start as ZEROS (nodes.py:430-436) and afterwards alias the reset slices of the
    post-update state; `forward` returns the POST-update output slice and `y` is the post-update state (nodes.py:457-465)."""

    spiking = True

    def __init__(self, func, args, var_map, param_map, dt, dtype=torch.float64, train_params=None,
                 spike_threshold=1e2, spike_reset=-1e2, spike_slope=None, spike_center=1.0):
        super().__init__(func, args, var_map, param_map, dt, dtype, train_params)
        if spike_slope is None:
            spike_slope = 100.0 / abs(spike_threshold - spike_reset)       # nodes.py:417-418
        self.slope = torch.tensor(spike_slope, dtype=dtype)
        self.center = torch.tensor(spike_center, dtype=dtype)
        n_spk = 1
        while f"spike_var_{n_spk}" in param_map:                           # nodes.py:422-425
            n_spk += 1
        self.spike_idx = [param_map[f"spike_var_{i}"] for i in range(n_spk)]
        self.slices = [var_map[f"spike_reset_{i}"] for i in range(n_spk)]
        self.v_reset = torch.tensor(spike_reset, dtype=dtype)
        self.thresh = torch.tensor(spike_threshold, dtype=dtype)
        self.y_reset = [torch.zeros(b - a, dtype=dtype) for a, b in self.slices]

    def forward(self, x):                                                  # nodes.py:451-465
        fired = []
        for idx, buf in zip(self.spike_idx, self.y_reset):
            spikes = OracleSpike.apply(buf - self.thresh, self.slope, self.center)
            fired.append(spikes.detach() > 0.0)
            self.args[idx] = spikes / self.dt
        self.args[self.inp] = x
        y_new = self.y + self.dt * self.func(0, self.y, *self.args)
        pieces, pos = [], 0
        for i, ((a, b), hit) in enumerate(zip(self.slices, fired)):        # masked assignment of the reset value = no gradient there
            pieces.append(y_new[pos:a])
            self.y_reset[i] = torch.where(hit, self.v_reset, y_new[a:b])
            pieces.append(self.y_reset[i])
            pos = b
        pieces.append(y_new[pos:])
        self.y = torch.cat(pieces, 0)
        return self.y[self.start:self.stop]


def field_ei_qif(n: int) -> Tuple[Callable, List[str]]:
    """Two coupled QIF populations in one node, each with its own spike variable (a MultiSpikeResetNet use case; not a shipped
    template -- the engine side is tests/test_gpu_jit.py::EI_YAML):
        v_e' = (v_e^2 + eta_e + I_ext)/tau_e + J_ee*s_in - J_ei*s_i ;  s_e' = -s_e/tau_s + spike_e
        v_i' = (v_i^2 + eta_i)/tau_i + J_ie*s_e                     ;  s_i' = -s_i/tau_s + spike_i      (s_in = weights @ s_e)"""
    names = ["weights", "eta_e", "tau_e", "J_ee", "J_ei", "eta_i", "tau_i", "J_ie", "tau_s", "I_ext", "spike_e", "spike_i"]

    def f(t, y, weights, eta_e, tau_e, J_ee, J_ei, eta_i, tau_i, J_ie, tau_s, I_ext, spike_e, spike_i):
        v_e, s_e, v_i, s_i = y[:n], y[n:2 * n], y[2 * n:3 * n], y[3 * n:]
        dv_e = (v_e * v_e + eta_e + I_ext) / tau_e + J_ee * (weights @ s_e) - J_ei * s_i
        dv_i = (v_i * v_i + eta_i) / tau_i + J_ie * s_e
        return torch.cat((dv_e, -s_e / tau_s + spike_e, dv_i, -s_i / tau_s + spike_i), 0)
    return f, names


EI_DEFAULTS = dict(eta_e=-5.0, tau_e=1.0, J_ee=1.0, J_ei=2.0, eta_i=-5.0, tau_i=0.5, J_ie=3.0, tau_s=0.8)


def build_ei_node_args(n: int, weights, params=None, dtype=torch.float64, y0=None):
    """(func, args, var_map, param_map) of the two-population node with the MultiSpikeResetNet index maps (nodes.py:438-449)."""
    func, names = field_ei_qif(n)
    p = dict(EI_DEFAULTS)
    p.update(params or {})
    if y0 is None:
        y0 = torch.cat([torch.full((n,), v, dtype=dtype) for v in (-2.0, 0.0, -2.0, 0.0)])
    args = [torch.as_tensor(y0, dtype=dtype).clone(), torch.as_tensor(weights, dtype=dtype).clone()]
    for name in names[1:]:
        if name in ("I_ext", "spike_e", "spike_i"):
            args.append(torch.zeros(n, dtype=dtype))
        else:
            args.append(torch.as_tensor(np.atleast_1d(p[name]), dtype=dtype).clone())
    param_map = {name: i for i, name in enumerate(names)}
    param_map.update({"in": param_map["I_ext"], "spike_var_0": param_map["spike_e"], "spike_var_1": param_map["spike_i"]})
    var_map = {"v_e": (0, n), "s_e": (n, 2 * n), "v_i": (2 * n, 3 * n), "s_i": (3 * n, 4 * n)}
    var_map.update({"out": var_map["s_e"], "spike_reset_0": var_map["v_e"], "spike_reset_1": var_map["v_i"]})
    return func, args, var_map, param_map


def make_node(model: str, n: int, weights, dt: float, params=None, dtype=torch.float64, train_params=None,
              input_var="I_ext", output_var=None, y0=None, **spike_kwargs):
    func, args, var_map, param_map = build_node_args(model, n, weights, params, dtype, input_var, output_var, y0)
    if model in SPIKING:
        return OracleSpikeResetNode(func, args, var_map, param_map, dt, dtype, train_params, **spike_kwargs)
    return OracleRateNode(func, args, var_map, param_map, dt, dtype, train_params)


# --------------------------------------------------------------------------------------------------
# edges (edges.py:8-65,150-174,177-234)
# --------------------------------------------------------------------------------------------------

def linear_forward(weights, x, mask=None):
    """Linear.forward = weights @ x (edges.py:48-49);  LinearMasked = (weights*mask) @ x (edges.py:173-174)."""
    return (weights * mask) @ x if mask is not None else weights @ x


class OracleRLS:
    """RLS edge (edges.py:177-234)."""

    def __init__(self, n_in, n_out, weights=None, dtype=torch.float64, beta=1.0, alpha=1.0):
        self.weights = torch.zeros((n_out, n_in), dtype=dtype) if weights is None else torch.as_tensor(weights, dtype=dtype).clone()
        self.beta = beta ** (-1)
        self.P = alpha * torch.eye(n_in, dtype=dtype)
        self.loss = 0.0

    def forward(self, x):
        return self.weights @ x

    def update(self, x, y, y_hat):                                         # edges.py:227-234
        z = self.beta * self.P @ x
        k = (1.0 + x @ z) ** (-1)
        err = y - y_hat
        self.weights = self.weights + torch.outer((y - k * x @ (self.weights + torch.outer(y, z)).T), z)
        self.P = self.P - k * torch.outer(z, z)
        self.loss = torch.inner(err, err)


# --------------------------------------------------------------------------------------------------
# network loop  (network.py:462-478,542-601,962-981)
# --------------------------------------------------------------------------------------------------

_ACT = {
    None: lambda x: x, "identity": lambda x: x, "tanh": torch.tanh, "sigmoid": torch.sigmoid,
    "softmax": lambda x: torch.softmax(x, 0), "softmin": lambda x: torch.softmax(-x, 0),
    "log_softmax": lambda x: torch.log_softmax(x, 0),
}


class OracleNet:
    """Chain  [in func] -> [W_in] -> diffeq node -> [W_out] -> [out func]  (the only graphs that work)."""

    def __init__(self, node, w_in=None, w_out=None, in_act=None, out_act=None, w_in_mask=None, w_out_mask=None,
                 has_in_node=None, has_out_node=None):
        self.node, self.w_in, self.w_out = node, w_in, w_out
        self.in_act, self.out_act = in_act, out_act
        self.w_in_mask, self.w_out_mask = w_in_mask, w_out_mask
        self.has_in_node = (w_in is not None) if has_in_node is None else has_in_node
        self.has_out_node = (w_out is not None) if has_out_node is None else has_out_node

    def parameters(self):
        ps = list(self.node.parameters())
        for w in (self.w_in, self.w_out):
            if w is not None and w.requires_grad:
                ps.append(w)
        return ps

    def forward(self, x):                                                  # network.py:462-478, 962-977
        if self.has_in_node:
            x = _ACT[self.in_act](x)
        if self.w_in is not None:
            x = linear_forward(self.w_in, x, self.w_in_mask)
        o = self.node.forward(x)
        if self.w_out is not None:
            o = linear_forward(self.w_out, o, self.w_out_mask)
        if self.has_out_node:
            o = _ACT[self.out_act](o)
        return o

    def get_var(self, name):
        return self.node.get(name)

    def run(self, inputs, sampling_steps=1, cutoff=0, record_vars: Sequence[Tuple[str, bool]] = (),
            enable_grad=True, truncate_steps=None):
        """Network.run (network.py:542-601) + Observer.record (observer.py:79-105).

        Returns dict(out=[...], steps=[...], vars={name: [...]}); window mean of outputs buffered since the
        previous record; recorded vars are instantaneous ``get_var`` at the record step (neuron-mean if flagged).
        """
        steps = inputs.shape[0]
        truncate_steps = steps if truncate_steps is None else truncate_steps
        out, rec_steps, buf = [], [], []
        rec = {name: [] for name, _ in record_vars}
        grad = torch.enable_grad if enable_grad else torch.no_grad
        with grad():
            for step in range(steps):
                o = self.forward(inputs[step, :])
                if step >= cutoff:
                    buf.append(o)
                    if step % sampling_steps == 0:
                        rec_steps.append(step)
                        out.append(torch.mean(torch.stack(buf, dim=0), dim=0))
                        for name, reduce in record_vars:
                            val = self.get_var(name)
                            rec[name].append(torch.mean(val) if reduce else val)
                        buf = []
                if truncate_steps < steps and step % truncate_steps == truncate_steps - 1:
                    self.node.detach()
        return dict(out=out, steps=rec_steps, vars=rec)


def run_trials(make_net: Callable[[], OracleNet], inputs: torch.Tensor, **run_kwargs):
    """Loop the (unbatched) reference path over independent trials: inputs [T, B, m] -> list of run() dicts."""
    results = []
    for b in range(inputs.shape[1]):
        net = make_net()
        results.append(net.run(inputs[:, b, :], **run_kwargs))
    return results


def bptt_grads(net: OracleNet, inputs, targets, sampling_steps=1, cutoff=0, loss="mse", truncate_steps=None):
    """One epoch of Network._bptt_epochs up to (not including) optimizer.step (network.py:993-999,1123-1130)."""
    res = net.run(inputs, sampling_steps=sampling_steps, cutoff=cutoff, enable_grad=True, truncate_steps=truncate_steps)
    pred = torch.stack(res["out"])
    if loss == "mse":
        err = torch.nn.functional.mse_loss(pred, targets)
    elif loss == "l1":
        err = torch.nn.functional.l1_loss(pred, targets)
    else:
        raise ValueError(loss)
    params = net.parameters()
    for p in params:
        p.grad = None
    err.backward()
    return err.detach(), pred.detach(), [p.grad.detach().clone() for p in params]


def ridge_fit(X: torch.Tensor, targets: torch.Tensor, alpha: float):
    """fit_ridge normal equations (network.py:765-768)."""
    Xt = X.T
    w = torch.inverse(Xt @ X + alpha * torch.eye(X.shape[1], dtype=X.dtype)) @ Xt @ targets
    return w, X @ w


# --------------------------------------------------------------------------------------------------
# helpers shared by tests / bench
# --------------------------------------------------------------------------------------------------

def lorentzian_etas(n: int, eta: float = -5.0, delta: float = 1.0) -> np.ndarray:
    """Recipe of documentation/qif_example.py:13."""
    return eta + delta * np.tan((np.pi / 2) * (2.0 * np.arange(1, n + 1) - n - 1) / (n + 1))


def spike_raster(v_hist: np.ndarray, thresh: float) -> np.ndarray:
    """p_t = 1[v_t >= theta]   (nodes.py:383,476 with center=1.0)."""
    return (v_hist >= thresh)


def compare_spikes(r_ref: np.ndarray, r_new: np.ndarray) -> Dict[str, float]:
    """Spike-train parity metrics for rasters [T, N]: count equality and spike-time offsets (steps)."""
    c_ref, c_new = r_ref.sum(0), r_new.sum(0)
    max_shift = 0
    unmatched = 0
    for n in range(r_ref.shape[1]):
        t_ref, t_new = np.flatnonzero(r_ref[:, n]), np.flatnonzero(r_new[:, n])
        k = min(len(t_ref), len(t_new))
        unmatched += abs(len(t_ref) - len(t_new))
        if k:
            max_shift = max(max_shift, int(np.abs(t_ref[:k] - t_new[:k]).max()))
    return dict(total_ref=int(c_ref.sum()), total_new=int(c_new.sum()),
                neurons_count_mismatch=int((c_ref != c_new).sum()), unmatched=int(unmatched), max_shift=max_shift)
