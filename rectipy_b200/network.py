"""`Network` -- the user interface, mirroring rectipy/network.py (add_diffeq_node, add_func_node, add_edge, add_node,
compile, forward, run, fit_bptt, fit_ridge, fit_rls, test, detach, reset, state, parameters, get/set/pop helpers).

What changes behind the unchanged API: `compile()` recognises the chain
    [input func node] -> [Linear W_in] -> diffeq node -> [Linear W_out] -> [output func node]
(the only graphs the reference can execute) and `run` / `forward` / `fit_*` dispatch the whole
horizon to the B200 engine (rectipy_b200/engine.py -> librectipy_b200.so) instead of looping over Python steps
(rectipy/network.py:588-599).  Unsupported graphs raise NotImplementedError; nothing falls back to CPU/eager code.

Extension over the reference: `Network(..., batch=B)` integrates B independent trials at once (inputs `[T, B, m]`);
the reference has no trial axis (rectipy/nodes.py:90).  With `batch=1` all shapes equal the reference's.
"""
from __future__ import annotations

import ctypes as C
import os

from dataclasses import dataclass
from time import perf_counter
from typing import Callable, Dict, Iterator, List, Optional, Tuple, Union

import numpy as np
import torch
from networkx import DiGraph

from . import _cabi as abi
from . import engine
from . import parallel
from .edges import RLS, Linear, LinearFilter, LinearMasked, LinearMemory, LinearMemoryFilter
from .nodes import InstantNode, RateNet, SpikeResetNet, node_from_template
from .observer import Observer
from .utility import add_op_name, retrieve_from_dict


@dataclass
class _Chain:
    diffeq: str
    in_func: Optional[str]
    out_func: Optional[str]
    in_edge: Optional[Linear]
    out_edge: Optional[Linear]


def _precision_code(precision: str, n: int, batch: int) -> int:
    if precision in ("fp32", "float32"):
        return abi.RP_PREC_FP32
    if precision in ("3xtf32", "tf32x3", "3xf16", "f16x3"):
        if not engine.tc_supported(n, batch):
            raise ValueError(f"precision={precision!r} needs n % 128 == 0 and batch % 128 == 0 (n={n}, batch={batch})")
        return abi.RP_PREC_3XTF32 if precision in ("3xtf32", "tf32x3") else abi.RP_PREC_3XF16
    if precision == "auto":
        return abi.RP_PREC_3XF16 if engine.tc_supported(n, batch) else abi.RP_PREC_FP32
    raise ValueError(f"unknown precision {precision!r}; use 'auto', 'fp32', '3xtf32' or '3xf16'")


def _record_steps(T: int, S: int, cutoff: int) -> List[int]:
    return [t for t in range(T) if t >= cutoff and t % S == 0]


def _window_mean(per_step: torch.Tensor, S: int, cutoff: int) -> torch.Tensor:
    """Observer windowing of Network.run (rectipy/network.py:590-597) applied to per-step outputs [T, ...]."""
    T = per_step.shape[0]
    steps = _record_steps(T, S, cutoff)
    if S == 1:
        return per_step[cutoff:]
    csum = torch.cumsum(per_step, dim=0)
    out, prev = [], cutoff - 1
    for r in steps:
        tot = csum[r] - (csum[prev] if prev >= 0 else 0.0)
        out.append(tot / float(r - prev))
        prev = r
    return torch.stack(out) if out else per_step[:0]


def _ceil128(v: int) -> int:
    return (v + 127) // 128 * 128


_PATH_CACHE: Dict[Tuple[int, int, int, int], int] = {}


def _padded_shape(node: RateNet, T: int) -> Optional[Tuple[int, int]]:
    """(n_pad, batch_pad) when the horizon should run on the tensor-core path with the trial / neuron axes padded to multiples of 128,
    else None.  Applies to precision="auto" on shapes the tcgen05 kernels do not take as they are (batch or n not a multiple of 128)
    and that the persistent few-trial kernels do not hold either (rp_plan_path): there the alternative is one fp32 FFMA contraction
    launch per step, 5-9x slower than the padded pass at 32..96 trials, break-even at 8 (profiles/r2_trial_padding.md, n=1024..4096).  Padded
    trials and neurons are inert by construction -- zero weight rows / columns, zero input and readout weights, zero loss gradient --
    so results and gradients of the real entries are those of the unpadded problem."""
    n, B = node.n, node.batch
    if node.precision != "auto" or node.spec.jit_program is not None or engine.tc_supported(n, B) or T < 8:
        return None
    n_pad, b_pad = _ceil128(n), _ceil128(B)
    if n_pad < 512 or B <= 8 or os.environ.get("RECTIPY_B200_NO_PADDING"):      # <= 8 trials: the per-step fp32 path is faster (measured)
        return None
    if n_pad != n and node.spec.model in (abi.RP_IKU, abi.RP_IK_BIEXP):       # population means would see the padded neurons
        return None
    dev = node.device.index if node.device.index is not None else torch.cuda.current_device()
    ck = (node.spec.model, n, B, dev)
    if ck not in _PATH_CACHE:
        d = abi.rp_desc()
        d.model, d.n, d.batch, d.precision = node.spec.model, n, B, abi.RP_PREC_FP32
        with torch.cuda.device(dev):
            _PATH_CACHE[ck] = int(abi.load().rp_plan_path(C.byref(d)))
    return (n_pad, b_pad) if _PATH_CACHE[ck] == 0 else None


def _pad_axis(t: torch.Tensor, dim: int, size: int, replicate: bool) -> torch.Tensor:
    """Grow `dim` of `t` to `size`: zeros, or copies of the last entry (state and parameters: a padded neuron / trial then behaves
    like a real one -- bounded, no special cases in the kernels -- while nothing it does can reach a real entry)."""
    cur = t.shape[dim]
    if cur == size:
        return t
    if replicate:
        idx = torch.arange(size, device=t.device).clamp_(max=cur - 1)
        return t.index_select(dim, idx)
    shape = list(t.shape)
    shape[dim] = size - cur
    return torch.cat((t, t.new_zeros(shape)), dim)


def _engine_call(node: RateNet, x: Optional[torch.Tensor], in_mode: int, W_in: Optional[torch.Tensor],
                 out_mode: int, W_out: Optional[torch.Tensor], T: int, S: int, cutoff: int, truncate: int,
                 rec_vars: Tuple[int, ...], rec_reduce: Tuple[int, ...], want_out: bool, out_var: Optional[int] = None):
    """Build/fetch the plan for this node + projection shape and run the whole horizon on the device."""
    if node.device.type != "cuda":
        raise RuntimeError("rectipy_b200 executes on a CUDA device only (there is no CPU fallback); create the "
                           f"Network with device='cuda:0' (got {node.device}).")
    slots, ptensors, per_neuron = node.param_slots()
    n_in = W_in.shape[1] if in_mode == abi.RP_IN_PROJ else 0
    n_out = W_out.shape[0] if out_mode == abi.RP_OUT_READOUT else 0
    prog = node.spec.jit_program
    if prog is not None and node.precision not in ("auto", "fp32", "float32"):
        raise NotImplementedError("rectipy_b200: run-time compiled templates run on the per-step fp32 path (precision='fp32' or 'auto')")
    jit_kw = {} if prog is None else dict(jit_key=prog.key, jit_nsv=prog.nsv, jit_spiking=int(prog.spiking), jit_post_out=int(prog.post_out),
                                            jit_src_plane=prog.src_plane)
    n, B = node.n, node.batch
    W, state = node["weights"], node.state
    rec_reduce = tuple(int(r) for r in rec_reduce)
    eng_reduce = rec_reduce
    pad = _padded_shape(node, T)
    if pad is not None:
        n, B = pad
        W = _pad_axis(_pad_axis(W, 0, n, False), 1, n, False)
        state = _pad_axis(_pad_axis(state, 2, n, True), 1, B, True)
        if W_in is not None:
            W_in = _pad_axis(W_in, 0, n, False)
        if W_out is not None:
            W_out = _pad_axis(W_out, 1, n, False)
        if x is not None:
            x = _pad_axis(x, 1, B, False)
            if in_mode == abi.RP_IN_DENSE:
                x = _pad_axis(x, 2, n, False)
            x = x.contiguous()
        padded = []
        for t_, mode in zip(ptensors, (per_neuron[s_] for s_ in slots)):       # 0 shared, 1 [n], 2 [B,1], 3 [B,n]
            if mode & 1:
                t_ = _pad_axis(t_, t_.dim() - 1, n, True)
            if mode >= 2:
                t_ = _pad_axis(t_, 0, B, True)
            padded.append(t_)
        ptensors = padded
        if n != node.n:
            eng_reduce = tuple(0 for _ in rec_reduce)          # neuron means are taken over the real neurons below
    key = engine.PlanKey(
        model=node.spec.model, n=n, batch=B, in_mode=in_mode, n_in=n_in, in_target=node.in_target,
        out_mode=out_mode, n_out=n_out, out_var=node.out_var if out_var is None else out_var,
        precision=abi.RP_PREC_FP32 if prog is not None else _precision_code(node.precision, n, B), dt=node.dt,
        theta=node.theta, v_reset=node.v_reset, slope=node.slope, per_neuron=per_neuron,
        device=node.device.index if node.device.index is not None else torch.cuda.current_device(), **jit_kw)
    plan = engine.get_plan(key)
    cfg = engine.RunConfig(T=T, sampling_steps=S, cutoff=cutoff, truncate_steps=truncate, rec_vars=tuple(rec_vars),
                           rec_reduce=eng_reduce, want_out=want_out, param_slots=slots)
    res = engine.EngineRun.apply(plan, cfg, x, W, W_in, W_out, state, *ptensors)
    out_rec, yT, recs = res[0], res[1], list(res[2:])
    if pad is not None:
        nb, nn = node.batch, node.n
        yT = yT[:, :nb, :nn].contiguous()
        if want_out:
            out_rec = out_rec[:, :nb] if out_mode == abi.RP_OUT_READOUT else out_rec[:, :nb, :nn]
        for i, red in enumerate(rec_reduce):
            if eng_reduce[i]:
                recs[i] = recs[i][:, :nb]
            else:
                recs[i] = recs[i][:, :nb, :nn]
                if red:
                    recs[i] = recs[i].mean(dim=-1)
    node._state = yT
    return (out_rec if want_out else None), recs


def _single_node_step(node: RateNet, x) -> torch.Tensor:
    """`node.forward(x)`: one Euler step, returns the PRE-update output slice (rectipy/nodes.py:166-170,382-392)."""
    x = torch.as_tensor(x, dtype=torch.float32, device=node.device) if not isinstance(x, torch.Tensor) else x.to(node.device, torch.float32)
    if x.dim() == 0:
        x = x.reshape(1)
    if x.dim() == 1:
        if x.shape[0] not in (1, node.n):
            raise RuntimeError(f"input of size {x.shape[0]} does not match the node's input size {node.n}")
        x = x.expand(node.n) if x.shape[0] == 1 else x
        x = x.reshape(1, 1, node.n).expand(1, node.batch, node.n)
    elif x.dim() == 2:
        if x.shape != (node.batch, node.n):
            raise RuntimeError(f"input of shape {tuple(x.shape)} does not match [batch={node.batch}, n={node.n}]")
        x = x.reshape(1, node.batch, node.n)
    else:
        raise RuntimeError("node input must be [n] or [batch, n]")
    out, _ = _engine_call(node, x.contiguous(), abi.RP_IN_DENSE, None, abi.RP_OUT_DENSE, None, 1, 1, 0, 0, (), (), True)
    out = out[0]
    return out[0] if node.batch == 1 else out


class Network:
    """Main user interface for initializing, training, testing, and running networks (rectipy/network.py:16-1194)."""

    def __init__(self, dt: float, device: str = "cuda:0", dtype: torch.dtype = torch.float32, batch: int = 1,
                 precision: str = "auto"):
        self.graph = DiGraph()
        self.device = device
        self.dtype = torch.float32
        if dtype not in (torch.float32, None):
            import warnings
            warnings.warn(f"rectipy_b200 computes in float32; Network dtype {dtype} is ignored.", stacklevel=2)
        self.dt = dt
        self.batch = int(batch)
        self.precision = precision
        self._record = {}
        self._var_map = {}
        self._in_node = None
        self._out_node = None
        self._bwd_graph = {}
        self._train_edge = ()
        self._chain: Optional[_Chain] = None

    # ---- container protocol --------------------------------------------------------------------------------
    def __getitem__(self, item):
        if isinstance(item, tuple):
            return self.graph[item[0]][item[1]]
        return self.graph.nodes[item]

    def __iter__(self):
        for n in self.graph.nodes:
            yield self[n]

    def __len__(self):
        return len(self.graph.nodes)

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    @property
    def n_out(self) -> int:
        try:
            return self[self._out_node]["n_out"]
        except (AttributeError, KeyError):
            return 0

    @property
    def n_in(self) -> int:
        try:
            return self[self._in_node]["n_in"]
        except (AttributeError, KeyError):
            return 0

    @property
    def nodes(self):
        return self.graph.nodes

    @property
    def state(self) -> dict:
        """State vectors of every differential-equation node (rectipy/network.py:88-98); snapshots, not views."""
        states = {}
        for n in self.nodes:
            node = self.get_node(n)
            if hasattr(node, "y"):
                states[n] = node.state.detach().clone()
        return states

    def get_node(self, node: str):
        return self[node]["node"]

    def get_edge(self, source: str, target: str) -> Linear:
        return self[source, target]["edge"]

    def get_var(self, node: str, var: str):
        try:
            return self.get_node(node)[self._relabel_var(var)]
        except KeyError:
            return self[node][var]

    def set_var(self, node: str, var: str, val):
        try:
            n = self.get_node(node)
            try:
                n.set_param(var, val)
            except KeyError:
                v = n[var]
                v[:] = torch.as_tensor(val, dtype=v.dtype, device=v.device)
        except KeyError:
            raise KeyError(f"Variable {var} was not found on node {node}.")

    # ---- graph construction ----------------------------------------------------------------------------------
    def add_node(self, label: str, node, node_type: str, op: str = None, **node_attrs) -> None:
        """Add a node instance to the graph (rectipy/network.py:178-211)."""
        if op:
            for p in node.parameter_names:
                add_op_name(op, p, self._var_map)
            for v in node.variable_names:
                add_op_name(op, v, self._var_map)
        self.graph.add_node(label, node=node, node_type=node_type, n_out=node.n_out, n_in=node.n_in, eval=True,
                            out=torch.zeros(node.n_out, device=self.device), **node_attrs)
        self._chain = None

    def add_diffeq_node(self, label: str, node, input_var: str, output_var: str, weights: np.ndarray = None,
                        source_var: str = None, target_var: str = None, spike_var=None, reset_var=None,
                        reset: bool = True, op: str = None, train_params: list = None, **kwargs) -> RateNet:
        """Add a differential-equation (rate or spiking) node (rectipy/network.py:213-306)."""
        var_dict = {"svar": source_var, "tvar": target_var, "in_ext": input_var, "out": output_var,
                    "spike": spike_var, "reset": reset_var}
        kwargs.pop("record_vars", None)
        self._var_map = {}
        if op is not None:
            for key, var in var_dict.copy().items():
                if type(var) is list:
                    var_dict[key] = [add_op_name(op, v, self._var_map) for v in var]
                else:
                    var_dict[key] = add_op_name(op, var, self._var_map)
            if train_params:
                train_params = [add_op_name(op, p, self._var_map) for p in train_params]
            if "node_vars" in kwargs:
                nv = {}
                for key, val in kwargs["node_vars"].items():
                    nv[key if "/" in key else f"all/{op}/{key}"] = val
                kwargs["node_vars"] = nv
        kwargs_tmp = {"weights": weights, "source_var": var_dict["svar"], "target_var": var_dict["tvar"],
                      "train_params": train_params, "device": self.device, "dt": self.dt}
        kwargs.setdefault("batch", self.batch)
        kwargs.setdefault("precision", self.precision)
        if spike_var is not None:
            if reset_var is None:
                raise ValueError("To define a reservoir with a spiking neural network layer, please provide the "
                                 "name of the variable that should be reset after a spike occurred (`reset_var`).")
            if not reset:
                raise NotImplementedError("rectipy_b200: `reset=False` selects the reference's SpikeNet, which is not "
                                          "executable upstream (rectipy/nodes.py:324 reads an undefined attribute).")
            kwargs_tmp["spike_var"] = var_dict["spike"]
            kwargs_tmp["reset_var"] = var_dict["reset"]
        kwargs.update(kwargs_tmp)
        node_obj = node_from_template(node, var_dict["in_ext"], var_dict["out"], **kwargs)
        self.add_node(label, node=node_obj, node_type="diff_eq", op=op)
        return node_obj

    def add_func_node(self, label: str, n: int, activation_function: str, **kwargs) -> InstantNode:
        """Add an activation-function node (rectipy/network.py:308-338)."""
        kwargs.pop("node_type", None)
        node = InstantNode(n, activation_function, **kwargs)
        self.add_node(label, node=node, node_type="func_instant")
        return node

    def add_edge(self, source: str, target: str, weights=None, train: Optional[str] = None, edge_attrs: dict = None,
                 **kwargs) -> Linear:
        """Add a linear projection between two nodes (rectipy/network.py:340-400)."""
        if not edge_attrs:
            edge_attrs = {}
        if "delays" in kwargs:                      # same selection rule as rectipy/network.py:375-383
            LinEdge = LinearMemoryFilter if "filter_weights" in kwargs else LinearMemory
        elif "filter_weights" in kwargs:
            LinEdge = LinearFilter
        else:
            LinEdge = LinearMasked if "mask" in kwargs else Linear
        kwargs.update({"n_in": self[source]["n_out"], "n_out": self[target]["n_in"], "weights": weights,
                       "dtype": self.dtype})
        trainable = True
        if train is None:
            trainable = False
            edge = LinEdge(**kwargs, detach=True)
        elif train == "gd":
            edge = LinEdge(**kwargs, detach=False)
        elif train == "rls":
            edge = RLS(**kwargs)
            self._train_edge = (source, target)
        else:
            raise ValueError("Invalid option for keyword argument `train`. Please see the docstring of "
                             "`Network.add_edge` for valid options.")
        self.graph.add_edge(source, target, edge=edge.to(self.device), trainable=trainable, n_in=edge.n_in,
                            n_out=edge.n_out, **edge_attrs)
        self._chain = None
        return edge

    def pop_node(self, node: str):
        node_data = self.get_node(node)
        self.graph.remove_node(node)
        self._chain = None
        return node_data

    def pop_edge(self, source: str, target: str) -> Linear:
        edge = self.get_edge(source, target)
        self.graph.remove_edge(source, target)
        self._chain = None
        return edge

    def clear(self):
        for node in list(self.nodes):
            self.pop_node(node)

    # ---- compile ---------------------------------------------------------------------------------------------
    def compile(self):
        """Find the unique input/output node (rectipy/network.py:439-460) and match the engine's chain pattern."""
        in_nodes = [n for n in self.graph.nodes if self.graph.in_degree(n) == 0]
        if len(in_nodes) != 1:
            raise ValueError(f"Unable to identify the input node of the Network. "
                             f"Nodes that have no input edges: {in_nodes}."
                             f"Make sure that exactly one such node without input edges exists in the network.")
        self._in_node = in_nodes.pop()
        out_nodes = [n for n in self.graph.nodes if self.graph.out_degree(n) == 0]
        if len(out_nodes) != 1:
            raise ValueError(f"Unable to identify the output node of the Network. "
                             f"Nodes that have no outgoing edges: {out_nodes}."
                             f"Make sure that exactly one such node without outgoing edges exists in the network.")
        self._out_node = out_nodes.pop()
        self._bwd_graph = self._compile_bwd_graph(self._out_node, dict())
        self._chain = None

    def _compile_bwd_graph(self, n: str, graph: dict) -> dict:
        sources = list(self.graph.predecessors(n))
        if len(sources) > 0:
            graph[n] = sources
        for s in sources:
            graph = self._compile_bwd_graph(s, graph)
        return graph

    def _get_chain(self) -> _Chain:
        """Match  [func] -> edge -> diffeq -> edge -> [func]; anything else has no engine plan."""
        if self._chain is not None:
            return self._chain
        if self._in_node is None or self._out_node is None:
            self.compile()
        diffeq = [n for n in self.graph.nodes if self[n]["node_type"] == "diff_eq"]
        if len(diffeq) != 1:
            raise NotImplementedError(f"rectipy_b200: the fused single-node plan needs exactly one differential-equation node "
                                      f"(found {len(diffeq)}: {diffeq}); feed-forward chains run through Network._get_path().")
        d = diffeq[0]
        g = self.graph
        preds, succs = list(g.predecessors(d)), list(g.successors(d))
        if len(preds) > 1 or len(succs) > 1:
            raise NotImplementedError("rectipy_b200: fan-in / fan-out at the differential-equation node is not supported "
                                      "(the reference cannot execute it either, rectipy/network.py:968).")
        in_func = in_edge = out_func = out_edge = None
        if preds:
            in_func = preds[0]
            if g.in_degree(in_func) != 0 or self[in_func]["node_type"] == "diff_eq":
                raise NotImplementedError("rectipy_b200: only one function node may precede the differential-equation node")
            in_edge = self.get_edge(in_func, d)
        if succs:
            out_func = succs[0]
            if g.out_degree(out_func) != 0 or self[out_func]["node_type"] == "diff_eq":
                raise NotImplementedError("rectipy_b200: only one function node may follow the differential-equation node")
            out_edge = self.get_edge(d, out_func)
        if len(g.nodes) != 1 + (in_func is not None) + (out_func is not None):
            raise NotImplementedError("rectipy_b200: the graph contains nodes outside the chain input -> diffeq -> output")
        self._chain = _Chain(d, in_func, out_func, in_edge, out_edge)
        return self._chain

    def _engine_device(self) -> torch.device:
        return next(self.get_node(n) for n in self.graph.nodes if self[n]["node_type"] == "diff_eq").device

    # ---- feed-forward chains of several differential-equation nodes ------------------------------------------
    def _is_multi(self) -> bool:
        """True when the graph runs node by node on per-step series: several diffeq nodes, or a stateful (delay / filter) edge."""
        if sum(1 for n in self.graph.nodes if self[n]["node_type"] == "diff_eq") > 1:
            return True
        return any(getattr(self.graph[s][t]["edge"], "stateful", False) for s, t in self.graph.edges)

    def _get_path(self) -> List[str]:
        """Nodes from the input node to the output node when the graph is one simple path (every graph the reference's
        recursive walk can execute, rectipy/network.py:962-973, is of this form: fan-in fails there).
        A node's output never depends on a later node, so node i can be integrated for the whole horizon before node
        i+1 starts: each differential-equation node is one engine call with dense per-step input and output, and
        autograd chains the calls (dense input gradients)."""
        if self._in_node is None or self._out_node is None:
            self.compile()
        g = self.graph
        if any(g.in_degree(n) > 1 or g.out_degree(n) > 1 for n in g.nodes):
            raise NotImplementedError("rectipy_b200: fan-in / fan-out is not supported (the reference cannot execute it "
                                      "either, rectipy/network.py:968).")
        path, n = [self._in_node], self._in_node
        while g.out_degree(n) == 1:
            n = next(iter(g.successors(n)))
            path.append(n)
        if len(path) != len(g.nodes) or path[-1] != self._out_node:
            raise NotImplementedError("rectipy_b200: the graph is not a single chain from the input to the output node")
        return path

    def _run_engine_multi(self, x: torch.Tensor, S: int, cutoff: int, truncate: int, rec_specs: list, want_out: bool):
        """Chain of several diffeq nodes: x [T,B,width] -> (out [n_rec,B,k] | None, [recorded vars], record steps)."""
        path = self._get_path()
        T = x.shape[0]
        steps = _record_steps(T, S, cutoff)
        series = x
        per_node: Dict[str, list] = {}
        for name, vi, red in rec_specs:
            per_node.setdefault(name, []).append((vi, red))
        got: Dict[Tuple[str, int, int], torch.Tensor] = {}
        for i, name in enumerate(path):
            if i > 0:
                edge = self.get_edge(path[i - 1], name)
                series = edge.apply_series(series) if edge.stateful else series @ edge.effective_weights().T
            node = self.get_node(name)
            if self[name]["node_type"] == "diff_eq":
                want = per_node.get(name, [])
                # per-step PRE-update outputs (nodes.py:170,392) and per-step state records; windows are applied at the end
                out, recs = _engine_call(node, series.contiguous(), abi.RP_IN_DENSE, None, abi.RP_OUT_DENSE, None, T, 1, 0,
                                         truncate, tuple(v for v, _ in want), tuple(r for _, r in want), True)
                for (vi, red), r in zip(want, recs):
                    got[(name, vi, red)] = r
                series = out
            else:
                series = node.apply_batched(series)
        out = _window_mean(series, S, cutoff) if want_out else None
        idx = torch.as_tensor(steps, dtype=torch.long, device=x.device)
        recs = [got[key].index_select(0, idx) for key in rec_specs]
        return out, recs, steps

    # ---- engine dispatch -----------------------------------------------------------------------------------
    def _prepare_inputs(self, inputs, T_axis: bool = True) -> torch.Tensor:
        """-> float32 device tensor [T, B, width]."""
        if self._is_multi():
            first = self._get_path()[0]
            fnode = self.get_node(first)
            dnode = next(self.get_node(n) for n in self._get_path() if self[n]["node_type"] == "diff_eq")
            node, width, has_in_func = dnode, (fnode.n if self[first]["node_type"] == "diff_eq" else self[first]["n_in"]), \
                self[first]["node_type"] != "diff_eq"
        else:
            chain = self._get_chain()
            node = self.get_node(chain.diffeq)
            width = self[chain.in_func]["n_in"] if chain.in_func is not None else node.n
            has_in_func = chain.in_func is not None
        x = inputs
        if not isinstance(x, torch.Tensor):
            x = torch.as_tensor(np.asarray(x), dtype=torch.float32)
        x = x.to(device=node.device, dtype=torch.float32, non_blocking=True)
        if x.dim() == 1:
            x = x.reshape(-1, 1) if T_axis else x.reshape(1, -1)
        if x.dim() == 2:
            x = x.unsqueeze(1).expand(x.shape[0], node.batch, x.shape[1])
        elif x.dim() != 3:
            raise RuntimeError("inputs must be [T, m] or [T, batch, m]")
        if x.shape[1] != node.batch:
            raise RuntimeError(f"inputs carry {x.shape[1]} trials but the network was built with batch={node.batch}")
        if x.shape[2] != width:
            if x.shape[2] == 1 and not has_in_func:
                x = x.expand(x.shape[0], x.shape[1], width)
            else:
                raise RuntimeError(f"Input dimensionality {x.shape[2]} does not match the network input size {width}.")
        return x

    def _run_engine(self, x: torch.Tensor, S: int, cutoff: int, truncate: int, rec_specs: list, want_out: bool):
        """x [T,B,width] -> (out [n_rec,B,k] | None, [recorded vars], record steps)."""
        if self._is_multi():
            return self._run_engine_multi(x, S, cutoff, truncate, rec_specs, want_out)
        chain = self._get_chain()
        node: RateNet = self.get_node(chain.diffeq)
        T = x.shape[0]
        # input side (network.py:962-977: input node activation, then the input edge)
        if chain.in_func is not None:
            x = self.get_node(chain.in_func).apply_batched(x)
        W_in = None
        if chain.in_edge is not None:
            W_in_eff = chain.in_edge.effective_weights()
            if W_in_eff.shape[1] <= abi.RP_MAX_IN:
                in_mode, W_in = abi.RP_IN_PROJ, W_in_eff
            else:
                in_mode, x = abi.RP_IN_DENSE, x @ W_in_eff.T
        else:
            in_mode = abi.RP_IN_DENSE
        # output side
        out_func = self.get_node(chain.out_func) if chain.out_func is not None else None
        nonlinear_out = out_func is not None and out_func.name != "identity"
        W_out = W_out_eff = None
        out_mode = abi.RP_OUT_DENSE
        if chain.out_edge is not None:
            W_out_eff = chain.out_edge.effective_weights()
            if W_out_eff.shape[0] <= abi.RP_MAX_OUT:
                out_mode, W_out = abi.RP_OUT_READOUT, W_out_eff
        eS, ecut = (1, 0) if (nonlinear_out and want_out) else (S, cutoff)
        rec_idx = tuple(r[0] for r in rec_specs)
        rec_red = tuple(r[1] for r in rec_specs)
        out, recs = _engine_call(node, x.contiguous(), in_mode, W_in, out_mode, W_out, T, eS, ecut, truncate,
                                 rec_idx, rec_red, want_out)
        steps = _record_steps(T, S, cutoff)
        if want_out:
            if chain.out_edge is not None and out_mode == abi.RP_OUT_DENSE:
                out = out @ W_out_eff.T                      # wide readout: window mean commutes with the projection
            if nonlinear_out:
                out = _window_mean(out_func.apply_batched(out), S, cutoff)
                if recs:
                    idx = torch.as_tensor(steps, dtype=torch.long, device=out.device)
                    recs = [r.index_select(0, idx) for r in recs]
        return out, recs, steps

    def _rec_specs(self, obs: Observer) -> list:
        if self._is_multi():          # (node, variable index, reduce) for every recorded variable of any diffeq node of the chain
            specs = []
            for (n, v), red in zip(obs.recorded_state_variables, obs.reduce_flags):
                if n not in self.graph.nodes or self[n]["node_type"] != "diff_eq":
                    raise KeyError(f"Variable {v} can only be recorded from a differential-equation node (got node {n}).")
                try:
                    specs.append((n, self.get_node(n).var_index(self._relabel_var(v)), int(red)))
                except KeyError:
                    raise KeyError(f"Variable {v} was not found on node {n}.")
            return specs
        chain = self._get_chain()
        specs = []
        for (n, v), red in zip(obs.recorded_state_variables, obs.reduce_flags):
            if n != chain.diffeq:
                raise KeyError(f"Variable {v} can only be recorded from the differential-equation node {chain.diffeq}.")
            node = self.get_node(n)
            try:
                specs.append((node.var_index(self._relabel_var(v)), int(red)))
            except KeyError:
                raise KeyError(f"Variable {v} was not found on node {n}.")
        return specs

    def _squeeze(self, t: Optional[torch.Tensor]) -> Optional[torch.Tensor]:
        """Drop the trial axis for batch == 1 so that shapes equal the reference's."""
        if t is None or self.batch != 1:
            return t
        return t.squeeze(1)

    # ---- forward / run ---------------------------------------------------------------------------------------
    def forward(self, x) -> torch.Tensor:
        """One integration step of the whole network (rectipy/network.py:462-478); returns the network output."""
        if self._is_multi():
            xt = self._prepare_inputs(x, T_axis=False)
            if xt.shape[0] != 1:
                raise RuntimeError("Network.forward expects the input of a single step")
            out, _, _ = self._run_engine_multi(xt, 1, 0, 0, [], True)
            return self._squeeze(out[0:1])[0] if self.batch == 1 else out[0]
        chain = self._get_chain()
        node = self.get_node(chain.diffeq)
        xt = self._prepare_inputs(x, T_axis=False)
        if xt.shape[0] != 1:
            raise RuntimeError("Network.forward expects the input of a single step")
        pre = node[node._out_key] if node.out_var != abi.RP_VAR_R else None
        pre = pre.detach().clone() if pre is not None else None
        out, _, _ = self._run_engine(xt, 1, 0, 0, [], True)
        out = self._squeeze(out[0:1])[0] if self.batch == 1 else out[0]
        # keep the per-node "out" attributes of the reference up to date (network.py:971-972; used by fit_rls)
        if chain.in_func is not None:
            self[chain.in_func]["out"] = self.get_node(chain.in_func).apply_batched(xt[0, 0] if self.batch == 1 else xt[0])
        if pre is not None:
            self[chain.diffeq]["out"] = pre
        if chain.out_func is not None:
            self[chain.out_func]["out"] = out
        elif pre is None:
            self[chain.diffeq]["out"] = out
        return out

    def parameters(self, recurse: bool = True) -> Iterator:
        for p in self._get_parameters(self.graph, recurse=recurse):
            yield p

    def _get_parameters(self, g: DiGraph, recurse: bool = True) -> Iterator:
        for node in g:
            for p in self.get_node(node).parameters(recurse=recurse):
                yield p
        for s, t in g.edges:
            for p in g[s][t]["edge"].parameters():
                yield p

    def detach(self, requires_grad: bool = True, detach_params: bool = False) -> None:
        for node in self.nodes:
            n = self.get_node(node)
            if hasattr(n, "y"):
                n.detach(requires_grad=requires_grad, detach_params=detach_params)

    def reset(self, state: dict = None):
        for node in self.nodes:
            n = self.get_node(node)
            if hasattr(n, "y"):
                if state and node in state:
                    n.reset(y=state[node])
                else:
                    n.reset()

    def run(self, inputs, sampling_steps: int = 1, cutoff: int = 0, verbose: bool = True, enable_grad: bool = True,
            **kwargs) -> Observer:
        """Integrate the input-driven network for `inputs.shape[0]` steps (rectipy/network.py:542-601).

        inputs: `[T, m]` (or `[T, batch, m]`); returns an `Observer` holding the window means of the output at every
        `sampling_steps`-th step >= `cutoff` and the requested state variables at those steps.
        """
        steps = inputs.shape[0]
        truncate_steps = kwargs.pop("truncate_steps", steps)
        self.compile()
        if "obs" in kwargs:
            obs = kwargs.pop("obs")
        else:
            obs = Observer(dt=self.dt, record_loss=kwargs.pop("record_loss", False), **kwargs)
        x = self._prepare_inputs(inputs)
        rec_specs = self._rec_specs(obs)
        grad = torch.enable_grad if enable_grad else torch.no_grad
        with grad():
            out, recs, rec_steps = self._run_engine(x, sampling_steps, cutoff, truncate_steps if truncate_steps < steps else 0,
                                                    rec_specs, obs._record_out)
        if verbose:
            print(f"Progress: {steps}/{steps} integration steps finished ({len(rec_steps)} samples recorded on device).")
        obs.record_block(rec_steps, self._squeeze(out), 0.0, [self._squeeze(r) for r in recs])
        return obs

    # ---- training --------------------------------------------------------------------------------------------
    def fit_bptt(self, inputs, targets, optimizer: str = "sgd", optimizer_kwargs: dict = None, loss: str = "mse",
                 loss_kwargs: dict = None, lr: float = 1e-3, sampling_steps: int = 1, update_steps: int = 100,
                 verbose: bool = True, **kwargs) -> Observer:
        """Backpropagation through time (rectipy/network.py:603-707).  The forward horizon runs in the engine; the
        reverse pass is the engine's fused adjoint, not autograd over an unrolled loop."""
        self.compile()
        loss_fn = self._get_loss_function(loss, loss_kwargs=loss_kwargs)
        opt = self._get_optimizer(optimizer, lr, self.parameters(), optimizer_kwargs=optimizer_kwargs)
        step_kwargs = retrieve_from_dict(["closure"], kwargs)
        error_kwargs = retrieve_from_dict(["retain_graph"], kwargs)
        obs_kwargs = retrieve_from_dict(["record_output", "record_loss", "record_vars"], kwargs)
        obs = Observer(dt=self.dt, **obs_kwargs)
        t0 = perf_counter()
        if type(inputs) is list:
            if len(inputs) != len(targets):
                raise ValueError("Wrong dimensions of input and target output. Please make sure that `inputs` and "
                                 "`targets` agree in the first dimension (epochs).")
            obs = self._bptt_epochs(inputs, targets, loss=loss_fn, optimizer=opt, obs=obs, error_kwargs=error_kwargs,
                                    step_kwargs=step_kwargs, sampling_steps=sampling_steps, verbose=verbose)
        else:
            inp = self._prepare_inputs(inputs)
            tgt = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets,
                                  dtype=torch.float32).to(inp.device)
            if inp.shape[0] != tgt.shape[0]:
                raise ValueError("Wrong dimensions of input and target output. Please make sure that `inputs` and "
                                 "`targets` agree in the first dimension.")
            obs = self._bptt(inp, tgt, loss_fn, opt, obs, error_kwargs, step_kwargs, sampling_steps=sampling_steps,
                             optim_steps=update_steps, verbose=verbose)
        if verbose:
            print(f"Finished optimization after {perf_counter() - t0} s.")
        return obs

    def _bptt_epochs(self, inp: list, target: list, loss: Callable, optimizer, obs: Observer, error_kwargs: dict,
                     step_kwargs: dict, sampling_steps: int = 1, verbose: bool = False, **kwargs) -> Observer:
        y0 = self.state
        epochs = len(inp)
        epoch_losses = []
        dev = self._engine_device()
        for epoch in range(epochs):
            obs = self.run(inp[epoch], verbose=False, sampling_steps=sampling_steps, enable_grad=True, **kwargs)
            tgt = torch.as_tensor(np.asarray(target[epoch]) if not isinstance(target[epoch], torch.Tensor) else target[epoch],
                                  dtype=torch.float32).to(dev)
            epoch_loss = self._bptt_step(torch.stack(obs["out"]), tgt, optimizer=optimizer, loss=loss,
                                         error_kwargs=error_kwargs, step_kwargs=step_kwargs)
            epoch_losses.append(epoch_loss)
            self.reset(y0)
            if verbose:
                print(f"Progress: {epoch + 1}/{epochs} training epochs finished.")
                print(f"Epoch loss: {epoch_loss}.")
                print("")
        obs.save("epoch_loss", epoch_losses)
        obs.save("epochs", np.arange(epochs))
        return obs

    def _bptt(self, inp: torch.Tensor, target: torch.Tensor, loss: Callable, optimizer, obs: Observer,
              error_kwargs: dict, step_kwargs: dict, sampling_steps: int = 100, optim_steps: int = 1000,
              verbose: bool = False) -> Observer:
        """Truncated BPTT (rectipy/network.py:1016-1048): an optimizer step + detach every `optim_steps` steps; the
        engine integrates each chunk in one call."""
        rec_specs = self._rec_specs(obs)
        steps = inp.shape[0]
        error = 0.0
        for start in range(0, steps, optim_steps):
            stop = min(start + optim_steps, steps)
            complete = (stop - start) == optim_steps
            with torch.enable_grad():
                # per-step predictions; recorded vars are read at the observer's own sampling grid below
                pred, recs, _ = self._run_engine(inp[start:stop], 1, 0, 0, rec_specs, True)
            pred_s = self._squeeze(pred)
            prev_error = error
            if complete:
                error = self._bptt_step(pred_s, target[start:stop], optimizer=optimizer, loss=loss,
                                        error_kwargs=error_kwargs, step_kwargs=step_kwargs)
                self.detach()
            rec_local = [t - start for t in range(start, stop) if t % sampling_steps == 0]
            if rec_local:
                idx = torch.as_tensor(rec_local, dtype=torch.long, device=pred.device)
                # the reference records a step before the optimizer step of its chunk has run -- all steps of a chunk but its
                # last carry the previous chunk's loss (rectipy/network.py:1035-1046)
                losses = [error if (complete and start + t == stop - 1) else prev_error for t in rec_local]
                obs.record_block([start + t for t in rec_local], pred_s.detach().index_select(0, idx), losses,
                                 [self._squeeze(r).index_select(0, idx) for r in recs])
            if verbose:
                print(f"Progress: {stop}/{steps} training steps finished. Current loss: {error}.")
        return obs

    def _bptt_step(self, predictions: torch.Tensor, targets: torch.Tensor, optimizer, loss: Callable, error_kwargs: dict,
                   step_kwargs: dict) -> float:
        """One optimizer step (rectipy/network.py:1123-1130).  When `torch.distributed` is initialised (one process per GPU,
        trials sharded over the ranks, parameters replicated) the gradients are summed over the ranks -- weighted by the
        ranks' trial counts, so the update is that of the global batch -- between `backward()` and `optimizer.step()`;
        every rank then applies the same update and the replicas stay identical.  The returned loss is the global one."""
        error = loss(predictions, targets)
        optimizer.zero_grad()
        error.backward(**error_kwargs)
        if parallel.dist.is_initialized() and parallel.dist.get_world_size() > 1:
            sync = self.allreduce_gradients(async_op=True, params=[p for g in optimizer.param_groups for p in g["params"]])
            # the scalar loss rides along while the gradient collectives are in flight
            n_loc, n_glob = self.batch, self._global_trials()
            err = parallel.allreduce_scalar(error.detach().double() * (n_loc / n_glob))
            if sync is not None:
                sync.wait()
            optimizer.step(**step_kwargs)
            return err.item()
        optimizer.step(**step_kwargs)
        return error.item()

    # ---- trial-sharded data parallelism (one process per GPU) -----------------------------------------
    def _global_trials(self) -> int:
        if getattr(self, "_dp_trials", None) is None or self._dp_trials[0] != self.batch:
            self._dp_trials = (self.batch, parallel.global_trial_count(self.batch, device=self._engine_device()))
        return self._dp_trials[1]

    def allreduce_gradients(self, async_op: bool = False, params=None):
        """Sum the `.grad` of every trainable parameter of the network over the ranks of `torch.distributed`, weighted by
        the ranks' trial counts (`batch` of each replica).  `fit_bptt` calls this itself before every optimizer step; a
        user-level loop (`obs = net.run(...); loss.backward(); net.allreduce_gradients(); opt.step()`) calls it by hand.
        No-op without an initialised process group.  async_op=True returns a handle whose `.wait()` must precede the
        optimizer step."""
        if not (parallel.dist.is_initialized() and parallel.dist.get_world_size() > 1):
            return None
        if params is None:
            params = list(self.parameters())
        return parallel.allreduce_gradients(params, self.batch, self._global_trials(), async_op=async_op)

    def fit_ridge(self, inputs, targets, sampling_steps: int = 100, alpha: float = 1e-4, verbose: bool = True,
                  add_readout_node: bool = True, **kwargs) -> Observer:
        """Ridge-regression readout on the recorded network states (rectipy/network.py:709-784); the state matrix
        stays on the device from the Observer recording to the normal-equation solve."""
        self.compile()
        dev = self._engine_device()
        target_tensor = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets,
                                        dtype=torch.float32).to(dev)
        if inputs.shape[0] != target_tensor.shape[0]:
            raise ValueError("Wrong dimensions of input and target output. Please make sure that `inputs` and "
                             "`targets` agree in the first dimension.")
        t0 = perf_counter()
        obs = self.run(inputs=inputs, sampling_steps=sampling_steps, verbose=False, enable_grad=False, **kwargs)
        if verbose:
            print(f"Finished network state collection after {perf_counter() - t0} s.")
        t0 = perf_counter()
        X = torch.stack(obs["out"])
        if X.dim() == 3:
            X = X.reshape(-1, X.shape[-1])
            target_tensor = target_tensor.reshape(-1, target_tensor.shape[-1])
        Xt = X.T
        A = Xt @ X + alpha * torch.eye(X.shape[1], device=dev, dtype=X.dtype)
        w_out = torch.linalg.solve(A, Xt @ target_tensor)
        y = X @ w_out
        if verbose:
            print(f"Finished fitting of read-out weights after {perf_counter() - t0} s.")
        if add_readout_node:
            self.add_func_node("readout", n=w_out.shape[1], activation_function="identity")
            self.add_edge(self._out_node, target="readout", weights=w_out.T.contiguous())
        obs.save("y", y)
        obs.save("w_out", w_out)
        return obs

    def fit_rls(self, inputs, targets, update_steps: int = 1, sampling_steps: int = 100, verbose: bool = True,
                **kwargs) -> Observer:
        """Recursive-least-squares training of the `train='rls'` readout edge (rectipy/network.py:786-856,1050-1121)."""
        self.compile()
        obs_kwargs = retrieve_from_dict(["record_output", "record_loss", "record_vars"], kwargs)
        obs = Observer(dt=self.dt, **obs_kwargs)
        t0 = perf_counter()
        if type(inputs) is list:
            if len(inputs) != len(targets):
                raise ValueError("Wrong dimensions of input and target output. Please make sure that `inputs` and "
                                 "`targets` agree in the first dimension (epochs).")
            y0 = self.state
            losses = []
            for inp_e, tgt_e in zip(inputs, targets):
                loss_e = self._rls(inp_e, tgt_e, None, sampling_steps, update_steps)
                losses.append(float(loss_e[-1]) if len(loss_e) else 0.0)
                self.reset(y0)
            obs.save("epoch_loss", losses)
            obs.save("epochs", np.arange(len(inputs)))
        else:
            if inputs.shape[0] != targets.shape[0]:
                raise ValueError("Wrong dimensions of input and target output. Please make sure that `inputs` and "
                                 "`targets` agree in the first dimension.")
            self._rls(inputs, targets, obs, sampling_steps, update_steps)
        if verbose:
            print(f"Finished optimization after {perf_counter() - t0} s.")
        return obs

    def _rls(self, inputs, targets, obs: Optional[Observer], sampling_steps: int, optim_steps: int) -> torch.Tensor:
        if not self._train_edge:
            raise ValueError("fit_rls requires an edge that was added with train='rls'.")
        chain = self._get_chain()
        if self.batch != 1:
            raise NotImplementedError("rectipy_b200: fit_rls is sequential over one trial (batch must be 1)")
        if chain.out_edge is None or (self._train_edge[0], self._train_edge[1]) != (chain.diffeq, chain.out_func):
            raise NotImplementedError("rectipy_b200: the RLS edge must be the readout edge diffeq -> output node")
        out_func = self.get_node(chain.out_func)
        if out_func.name != "identity":
            raise NotImplementedError("rectipy_b200: fit_rls supports an identity output node")
        edge: RLS = chain.out_edge
        node = self.get_node(chain.diffeq)
        x = self._prepare_inputs(inputs)
        tgt = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets,
                              dtype=torch.float32).to(node.device)
        rec_specs = self._rec_specs(obs) if obs is not None else []
        # 1) integrate the reservoir, recording the readout's source activity at every step (no feedback => exact)
        saved_edge = chain.out_edge
        chain.out_edge, chain_out_func = None, chain.out_func
        try:
            with torch.no_grad():
                chain.out_func = None
                X, recs, _ = self._run_engine(x, 1, 0, 0, rec_specs, True)
        finally:
            chain.out_edge, chain.out_func = saved_edge, chain_out_func
        X = X[:, 0, :].contiguous()
        # 2) sequential RLS over the recorded states on the device
        edge.weights = edge.weights.contiguous()
        edge.P = edge.P.contiguous()
        loss, pred = engine.rls_run(X, tgt.reshape(X.shape[0], -1), edge.weights, edge.P, edge.beta, optim_steps)
        if optim_steps > 1 and loss.numel():
            # between updates the reference keeps reporting the loss of the last update (rectipy/network.py:1108-1118)
            held = (torch.arange(loss.shape[0], device=loss.device) // optim_steps) * optim_steps
            loss = loss.index_select(0, held)
        edge.loss = loss[-1] if loss.numel() else 0.0
        if obs is not None:
            steps = [t for t in range(X.shape[0]) if t % sampling_steps == 0]
            idx = torch.as_tensor(steps, dtype=torch.long, device=X.device)
            obs.record_block(steps, pred.index_select(0, idx), loss.index_select(0, idx),
                             [self._squeeze(r).index_select(0, idx) for r in recs])
        return loss

    def fit_eprop(self, *args, **kwargs):
        raise NotImplementedError("Method is currently not implemented")   # as in the reference (network.py:896)

    def test(self, inputs, targets, loss: str = "mse", loss_kwargs: dict = None, sampling_steps: int = 100,
             verbose: bool = True, **kwargs) -> tuple:
        """Loss of the frozen model on test data (rectipy/network.py:898-944)."""
        loss_fn = self._get_loss_function(loss, loss_kwargs=loss_kwargs)
        obs = self.run(inputs=inputs, sampling_steps=sampling_steps, verbose=verbose, enable_grad=False, **kwargs)
        output = torch.stack(obs["out"])
        target_tensor = torch.as_tensor(np.asarray(targets) if not isinstance(targets, torch.Tensor) else targets,
                                        dtype=torch.float32).to(output.device)
        return obs, loss_fn(output, target_tensor).item()

    def _relabel_var(self, var: str) -> str:
        try:
            return self._var_map[var]
        except KeyError:
            return var

    @staticmethod
    def _get_optimizer(optimizer: str, lr: float, model_params: Iterator, optimizer_kwargs: dict = None):
        if optimizer_kwargs is None:
            optimizer_kwargs = {}
        table = {"sgd": torch.optim.SGD, "adam": torch.optim.Adam, "adamw": torch.optim.AdamW,
                 "adagrad": torch.optim.Adagrad, "adadelta": torch.optim.Adadelta, "adamax": torch.optim.Adamax,
                 "rmsprop": torch.optim.RMSprop, "rprop": torch.optim.Rprop}
        if optimizer not in table:
            raise ValueError("Invalid optimizer choice. Please see the documentation of the `Network.fit_bptt()` "
                             "method for valid options.")
        return table[optimizer](model_params, lr=lr, **optimizer_kwargs)

    @staticmethod
    def _get_loss_function(loss: str, loss_kwargs: dict = None) -> Callable:
        if loss_kwargs is None:
            loss_kwargs = {}
        table = {"mse": torch.nn.MSELoss, "l1": torch.nn.L1Loss, "nll": torch.nn.NLLLoss,
                 "ce": torch.nn.CrossEntropyLoss, "kld": torch.nn.KLDivLoss, "hinge": torch.nn.HingeEmbeddingLoss}
        if loss not in table:
            raise ValueError("Invalid loss function choice. Please see the documentation of the `Network.fit_bptt()` "
                             "method for valid options.")
        return table[loss](**loss_kwargs)
