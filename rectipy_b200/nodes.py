"""Differential-equation and function nodes -- host-side mirror of rectipy/nodes.py.

The node objects keep the reference's duck-type protocol (`forward`, `parameters`, `detach`, `reset`, `set_param`,
`__getitem__`, `y`, `n_in`, `n_out`, `parameter_names`, `variable_names`; used at rectipy/network.py:95,113,148,
171-174,204-211,513-514,530-534,948,971) but hold no vector-field code: a node is a template id, its parameter
tensors and a state tensor in the engine's `[var][trial][neuron]` layout.  `forward(x)` advances one Euler step on the
device through the same C-ABI entry the whole-horizon run uses (T = 1).
"""
from __future__ import annotations

import warnings
from typing import Dict, Iterator, List, Optional, Sequence, Tuple, Union

import numpy as np
import torch

from . import _cabi as abi
from .templates import TemplateSpec, resolve_template


class InstantNode:
    """Activation-function node (rectipy/nodes.py:14-51)."""

    _FUNCS = {"tanh": torch.nn.Tanh, "softmax": torch.nn.Softmax, "softmin": torch.nn.Softmin,
              "log_softmax": torch.nn.LogSoftmax, "sigmoid": torch.nn.Sigmoid, "identity": torch.nn.Identity}

    def __init__(self, n: int, func: str, **kwargs):
        if func not in self._FUNCS:
            raise ValueError(f"Invalid keyword argument `func`: {func} is not a valid option. See the docstring for "
                             f"`Network.add_func_node` for valid options.")
        if func in ("softmax", "softmin", "log_softmax") and "dim" not in kwargs:
            kwargs["dim"] = 0
        self.name = func
        self.n_in = n
        self.n_out = n
        self.func = self._FUNCS[func](**kwargs)
        self.kwargs = kwargs

    def __getitem__(self, item):
        return None

    def __call__(self, x):
        return self.forward(x)

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        return self.func.forward(x)

    def apply_batched(self, x: torch.Tensor) -> torch.Tensor:
        """Apply to [..., n]: the reference applies dim-0 reductions to a single [n] vector => last axis here."""
        if self.name in ("softmax", "softmin", "log_softmax"):
            fn = {"softmax": torch.softmax, "softmin": lambda t, dim: torch.softmax(-t, dim),
                  "log_softmax": torch.log_softmax}[self.name]
            return fn(x, dim=-1)
        return self.func.forward(x)

    def parameters(self, **kwargs):
        return self.func.parameters(**kwargs)


def _as_param(val, n: int, device, batch: int = 1) -> torch.Tensor:
    """scalar -> [1]; one value per neuron -> [n]; parameter sweeps: [B, 1] (one value per trial) or [B, n]."""
    t = torch.as_tensor(np.asarray(val, dtype=np.float32) if not isinstance(val, torch.Tensor) else val.detach().to(torch.float32))
    if t.dim() == 2 and batch > 1 and t.shape[0] == batch and t.shape[1] in (1, n):
        return t.to(device).contiguous().clone()
    t = t.reshape(-1).to(device).clone()
    if t.numel() not in (1, n):
        raise ValueError(f"node parameter must be a scalar, have one value per neuron ({n}), or be a [batch, 1] / [batch, {n}] "
                         f"sweep; got shape {tuple(np.shape(val))}")
    return t


class RateNet:
    """Rate-neuron node: y <- y + dt*f(y); output is the PRE-update slice (rectipy/nodes.py:54-211)."""

    spiking = False

    def __init__(self, spec: TemplateSpec, n: int, weights, dt: float = 1e-3, dtype: torch.dtype = torch.float32,
                 train_params: Optional[list] = None, device: str = "cuda:0", input_var: str = "I_ext",
                 output_var: Optional[str] = None, node_vars: Optional[dict] = None, batch: int = 1,
                 precision: str = "auto", **kwargs):
        if dtype not in (torch.float32, None):
            warnings.warn("rectipy_b200 integrates in float32 on the device; the requested node dtype "
                          f"{dtype} is ignored.", stacklevel=3)
        self.spec = spec
        self.n = int(n)
        self.batch = int(batch)
        self.dt = float(dt)
        self.device = torch.device(device)
        self.dtype = torch.float32
        self.precision = precision
        # parameters (rnn_args[1:] of the reference): recurrent weights + template parameters
        W = torch.as_tensor(np.asarray(weights, dtype=np.float32) if not isinstance(weights, torch.Tensor) else weights.detach().to(torch.float32))
        if W.shape != (self.n, self.n):
            raise ValueError(f"recurrent weights must be {self.n} x {self.n}, got {tuple(W.shape)}")
        self._params: Dict[str, torch.Tensor] = {"weights": W.to(self.device).contiguous().clone()}
        for key, (slot, default) in spec.params.items():
            self._params[key] = _as_param(default, self.n, self.device)
        for key, val in (node_vars or {}).items():
            try:
                pkey = spec.resolve(key, spec.params)
                self._params[pkey] = _as_param(val, self.n, self.device, self.batch)
            except KeyError:
                try:
                    vkey = spec.resolve(key, dict(spec.state_vars))
                except KeyError:
                    raise KeyError(f"Variable {key} was not found on the node template.")
                self._init_override = getattr(self, "_init_override", {})
                self._init_override[vkey] = val
        # variable slices inside the reference-style flat state vector
        self._var_map: Dict[str, Tuple[int, int]] = {}
        for i, (name, _) in enumerate(spec.state_vars):
            self._var_map[name] = (i * self.n, (i + 1) * self.n)
        self._in_key = spec.resolve(input_var, spec.input_vars)
        self.in_target = spec.input_vars[self._in_key]
        if output_var is None:
            output_var = spec.source_var if spec.spiking else spec.state_vars[0][0]
        try:
            self._out_key = spec.resolve(output_var, spec.out_vars)
        except KeyError:
            if spec.jit_program is not None:
                raise NotImplementedError(f"rectipy_b200.jit: output_var {output_var!r} must be one of the state variables "
                                          f"{sorted(spec.out_vars)} of a run-time compiled template")
            raise
        self.out_var = spec.out_vars[self._out_key]
        if self.out_var != abi.RP_VAR_R:
            self._var_map["out"] = self._var_map[self._out_key]
        self.n_out = self.n
        self.n_in = self.n
        # trainable parameters (nodes.py:79-86)
        self._train_keys: List[str] = []
        for p in (train_params or []):
            self._train_keys.append(self._param_key(p))
        for k in self._train_keys:
            if k != "weights" and self._params[k].dim() == 2:
                raise NotImplementedError(f"rectipy_b200: per-trial parameter {k} (a sweep) cannot be trained")
            self._params[k].requires_grad_(True)
        # state: [n_sv, B, n]
        st = torch.empty((spec.n_sv, self.batch, self.n), dtype=torch.float32)
        over = getattr(self, "_init_override", {})
        self._planes = spec.plane_order                       # engine plane of the i-th variable of the reference's y
        for i, (name, init) in enumerate(spec.state_vars):
            st[self._planes[i]] = torch.as_tensor(np.asarray(over.get(name, init), dtype=np.float32))
        self._state = st.to(self.device)
        self._y0_template = self._state.clone()

    @classmethod
    def from_pyrates(cls, node, input_var: str, output_var: str, weights=None, source_var: str = None,
                     target_var: str = None, train_params: list = None, **kwargs):
        """Same entry point as the reference (rectipy/nodes.py:112-164,363-380); no PyRates involved: the template is
        recognised by `rectipy_b200.templates` and mapped onto a compiled vector field."""
        if cls is RateNet:
            kwargs.pop("spike_var", None)
            kwargs.pop("reset_var", None)
            return node_from_template(node, input_var, output_var, weights=weights, source_var=source_var,
                                      target_var=target_var, train_params=train_params, **kwargs)
        kwargs.setdefault("spike_var", "spike")
        kwargs.setdefault("reset_var", "v")
        return node_from_template(node, input_var, output_var, weights=weights, source_var=source_var,
                                  target_var=target_var, train_params=train_params, **kwargs)

    # ---- reference-style views ---------------------------------------------------------------------------
    @property
    def train_params(self) -> List[torch.Tensor]:
        return [self._params[k] for k in self._train_keys]

    @property
    def y(self) -> torch.Tensor:
        """Flat state vector `[n_sv*n]` (one trial) or `[B, n_sv*n]`, ordered like the reference's `y` (nodes.py:90)."""
        st = self._state if self._planes == sorted(self._planes) else self._state[self._planes]
        if self.batch == 1:
            return st.reshape(-1)
        return st.permute(1, 0, 2).reshape(self.batch, -1)

    @y.setter
    def y(self, val):
        self._set_state(val)

    def _set_state(self, val):
        t = torch.as_tensor(val) if not isinstance(val, torch.Tensor) else val
        t = t.detach().to(device=self.device, dtype=torch.float32)
        nsv = self.spec.n_sv
        if t.shape == (nsv, self.batch, self.n):
            st = t
        elif t.dim() == 1 and t.numel() == nsv * self.n:
            st = t.reshape(nsv, 1, self.n).expand(nsv, self.batch, self.n)
        elif t.dim() == 2 and t.shape == (self.batch, nsv * self.n):
            st = t.reshape(self.batch, nsv, self.n).permute(1, 0, 2)
        else:
            raise RuntimeError(f"state of shape {tuple(t.shape)} does not match n_sv={nsv}, batch={self.batch}, n={self.n}")
        if t.shape != (nsv, self.batch, self.n) and self._planes != sorted(self._planes):
            inv = [self._planes.index(p) for p in range(nsv)]      # reference order -> engine planes
            st = st[inv]
        self._state = st.contiguous().clone()

    @property
    def state(self) -> torch.Tensor:
        """Engine-layout state `[n_sv, B, n]`."""
        return self._state

    @property
    def parameter_names(self) -> list:
        names = ["weights", "in"] + list(self.spec.params.keys()) + list(self.spec.input_vars.keys())
        if self.spec.spiking:
            names.append("spike_var")
        return names

    @property
    def variable_names(self) -> list:
        return list(self._var_map.keys())

    def _param_key(self, name: str) -> str:
        if name == "weights" or name.endswith("/weight") or name.endswith("in_edge_0/weight"):
            return "weights"
        try:
            return self.spec.resolve(name, self.spec.params)
        except KeyError:
            raise KeyError(f"Parameter {name} was not found on the node.")

    def __getitem__(self, item: str):
        try:
            return self._params[self._param_key(item)]
        except KeyError:
            pass
        if item == "out" and self.out_var != abi.RP_VAR_R:
            a, b = self._var_map["out"]
        else:
            try:
                key = self.spec.resolve(item, dict(self.spec.state_vars))
            except KeyError:
                raise KeyError(item)
            a, b = self._var_map[key]
        i = self._planes[a // self.n]
        return self._state[i, 0] if self.batch == 1 else self._state[i]

    def var_index(self, name: str) -> int:
        """Index of a state variable in the engine layout (KeyError if `name` is not a state variable)."""
        key = self.spec.resolve(name, dict(self.spec.state_vars))
        return self._planes[self._var_map[key][0] // self.n]

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def parameters(self, recurse: bool = True) -> Iterator:
        for k in self._train_keys:
            yield self._params[k]

    def set_param(self, param: str, val):
        key = self._param_key(param)
        old = self._params[key]
        if key == "weights":
            new = torch.as_tensor(val).detach().to(device=self.device, dtype=torch.float32).reshape(self.n, self.n).clone()
        else:
            new = _as_param(val, self.n, self.device, self.batch)
        new.requires_grad_(old.requires_grad)
        self._params[key] = new

    # ---- dynamics ----------------------------------------------------------------------------------------
    def forward(self, x: torch.Tensor) -> torch.Tensor:
        """One Euler step with the input current `x` ([n] or [1] or [B, n]); returns the pre-update output slice."""
        from .network import _single_node_step
        return _single_node_step(self, x)

    def detach(self, requires_grad: bool = False, detach_params: bool = False):
        """Cut the state from the autograd graph (rectipy/nodes.py:176-196)."""
        self._state = self._state.detach()
        if detach_params:
            for k in self._train_keys:
                p = self._params[k].detach()
                p.requires_grad_(requires_grad)
                self._params[k] = p

    def reset(self, y=None, idx=None):
        """Restore the state; `None` zeroes it like the reference does (rectipy/nodes.py:198-211)."""
        if y is None:
            self._state = torch.zeros_like(self._state)
            return
        if idx is None:
            self._set_state(y)
            return
        flat = self.y.clone()
        ii = torch.as_tensor(np.asarray(idx), dtype=torch.long, device=self.device)
        yy = torch.as_tensor(y).detach().to(device=self.device, dtype=torch.float32)
        if self.batch == 1:
            flat[ii] = yy
        else:
            flat[:, ii] = yy
        self._set_state(flat)

    def reset_to_template(self):
        self._state = self._y0_template.clone()

    # ---- engine description --------------------------------------------------------------------------------
    def param_slots(self) -> Tuple[Tuple[int, ...], List[torch.Tensor], Tuple[int, ...]]:
        slots, tensors = [], []
        per_neuron = [0] * abi.RP_NUM_PARAMS
        for key, (slot, _) in self.spec.params.items():
            t = self._params[key]
            slots.append(slot)
            tensors.append(t)
            if t.dim() == 2:                      # parameter sweep: [B, 1] or [B, n]
                per_neuron[slot] = 3 if t.shape[1] == self.n and self.n > 1 else 2
            else:
                per_neuron[slot] = 1 if (t.numel() == self.n and self.n > 1) else 0
        return tuple(slots), tensors, tuple(per_neuron)

    @property
    def theta(self) -> float:
        return 0.0

    @property
    def v_reset(self) -> float:
        return 0.0

    @property
    def slope(self) -> float:
        return 0.0


class SpikeResetNet(RateNet):
    """Spiking node: threshold -> spike -> Euler -> reset (rectipy/nodes.py:333-401), surrogate gradient nodes.py:468-481."""

    spiking = True

    def __init__(self, spec: TemplateSpec, n: int, weights, spike_threshold: float = 1e2, spike_reset: float = -1e2,
                 spike_var: str = "spike", reset_var: str = "v", **kwargs):
        spike_slope = kwargs.pop("spike_slope", None)
        spike_center = kwargs.pop("spike_center", 1.0)
        super().__init__(spec, n, weights, **kwargs)
        if not spec.spiking:
            raise ValueError(f"template {spec.name} has no spike input variable")
        if float(spike_center) != 1.0:
            raise NotImplementedError("rectipy_b200 implements the reference default spike_center=1.0 (spike iff v >= theta)")
        if spec.resolve(spike_var, {spec.spike_var: 0}) != spec.spike_var:
            raise KeyError(spike_var)
        rkey = spec.resolve(reset_var, dict(spec.state_vars))
        if self._planes[self._var_map[rkey][0] // self.n] != 0:       # compiled fields: v; run-time compiled fields put it in plane 0
            raise NotImplementedError("rectipy_b200 resets the membrane potential `v`; other reset variables are not supported")
        self._var_map["reset_var"] = self._var_map[rkey]
        self._thresh = float(spike_threshold)
        self._reset = float(spike_reset)
        self._slope = float(spike_slope) if spike_slope is not None else 100.0 / abs(self._thresh - self._reset)

    @property
    def theta(self) -> float:
        return self._thresh

    @property
    def v_reset(self) -> float:
        return self._reset

    @property
    def slope(self) -> float:
        return self._slope


class MultiSpikeResetNet(SpikeResetNet):
    """Several spike / reset variable pairs in one node (rectipy/nodes.py:404-465): every pair is thresholded and reset like
    SpikeResetNet's, one shared threshold / reset value; `forward` returns the POST-update output slice and `y` is the post-update
    state (nodes.py:457-465).  Runs on kernels generated from the template's equations (rectipy_b200/jit.py).
    The reference thresholds a zero vector on the very first step (`_y_reset` is initialised with zeros, nodes.py:430-436); the
    engine thresholds the actual initial state -- identical unless an initial reset variable is already >= theta or theta <= 0."""

    def __init__(self, spec: TemplateSpec, n: int, weights, spike_threshold: float = 1e2, spike_reset: float = -1e2,
                 spike_var=("spike",), reset_var=("v",), **kwargs):
        spike_slope = kwargs.pop("spike_slope", None)
        spike_center = kwargs.pop("spike_center", 1.0)
        RateNet.__init__(self, spec, n, weights, **kwargs)
        prog = spec.jit_program
        if prog is None or not prog.post_out or prog.spiking != len(spike_var):
            raise ValueError("MultiSpikeResetNet needs a template bound with lists of spike and reset variables")
        if float(spike_center) != 1.0:
            raise NotImplementedError("rectipy_b200 implements the reference default spike_center=1.0 (spike iff v >= theta)")
        for j, r in enumerate(reset_var):
            rkey = spec.resolve(r, dict(spec.state_vars))
            assert self._planes[self._var_map[rkey][0] // self.n] == j
            self._var_map[f"spike_reset_{j}"] = self._var_map[rkey]
        self._thresh = float(spike_threshold)
        self._reset = float(spike_reset)
        self._slope = float(spike_slope) if spike_slope is not None else 100.0 / abs(self._thresh - self._reset)


def node_from_template(node, input_var: str, output_var: str, weights=None, source_var: str = None,
                       target_var: str = None, spike_var=None, reset_var=None, train_params: list = None,
                       **kwargs) -> RateNet:
    """Counterpart of `RateNet.from_pyrates` / `SpikeResetNet.from_pyrates` (rectipy/nodes.py:112-164,363-380)."""
    multi = isinstance(spike_var, (list, tuple))
    if multi and not isinstance(reset_var, (list, tuple)):
        raise ValueError("a list of spike variables needs a list of reset variables of the same length")
    spec = resolve_template(node, force_jit=multi)          # MultiSpikeResetNet semantics only exist as generated kernels
    if spec.jit_field is not None and spec.jit_program is None:
        # equations that match no compiled field: generate + compile the kernels now that the variable roles are known
        from . import jit
        spec = jit.bind_spec(spec, source_var if weights is not None else None, target_var if weights is not None else None,
                             input_var, spike_var, reset_var if spike_var is not None else None)
    for k in ("clear", "float_precision", "file_name", "verbose", "auto_diff", "vectorize", "backend", "solver"):
        kwargs.pop(k, None)           # PyRates code-generation switches; nothing to generate here
    kwargs.pop("var_mapping", None)
    kwargs.pop("param_mapping", None)
    if "node_values" in kwargs:
        kwargs["node_vars"] = kwargs.pop("node_values")
    if weights is None:
        n = kwargs.pop("N", None)
        if n is None:
            raise ValueError("Either `weights` or the number of neurons `N` has to be provided.")
        weights = np.zeros((n, n), dtype=np.float32)
    else:
        kwargs.pop("N", None)
        weights = np.asarray(weights) if not isinstance(weights, torch.Tensor) else weights
        if source_var is None or target_var is None:
            raise ValueError("If synaptic weights are passed (`weights`), please provide the names of the source "
                             "and target variable that should be connected via `weights`.")
        if spec.resolve(source_var, {spec.source_var: 0, **dict(spec.state_vars)}) != spec.source_var or \
                spec.resolve(target_var, {spec.target_var: 0}) != spec.target_var:
            raise NotImplementedError(f"rectipy_b200 couples {spec.source_var} -> {spec.target_var} for template "
                                      f"{spec.name}; got {source_var} -> {target_var}")
    if weights.shape[0] != weights.shape[1]:
        raise ValueError("`weights` has to be a square matrix")
    n = weights.shape[0]
    if spike_var is None:
        return RateNet(spec, n, weights, train_params=train_params, input_var=input_var, output_var=output_var, **kwargs)
    if multi:
        return MultiSpikeResetNet(spec, n, weights, train_params=train_params, input_var=input_var, output_var=output_var,
                                  spike_var=list(spike_var), reset_var=list(reset_var), **kwargs)
    return SpikeResetNet(spec, n, weights, train_params=train_params, input_var=input_var, output_var=output_var,
                         spike_var=spike_var, reset_var=reset_var, **kwargs)
