"""`FeedbackNetwork`: networks with feedback edges (rectipy/network.py:1196-1357, documentation/rnn_tryout.py).

Reference semantics, restated: `compile()` takes the edges flagged `feedback=True` out of the graph, so that the remaining
feed-forward graph defines the input node, the output node and the evaluation order (network.py:1204-1228).  During one
network step every node is evaluated once along the feed-forward path; before a node is evaluated, the feedback edges that
end at it are applied to `get_node(source)["out"]` (network.py:1354-1357) -- for a differential-equation node that is the
output slice of the source node's attribute `y` (nodes.py:92-99).  What `y` holds depends on the node class: a `RateNet`
keeps its current state there (nodes.py:169: already updated in this step when the source precedes the target on the
feed-forward path, not yet updated otherwise), a `SpikeResetNet` leaves `y` at the state BEFORE its last step (nodes.py:387;
the live state is in `_y_start/_y_spike/_y_stop`), so feedback from a spiking population that sits later on the path arrives
with one step of delay (the state itself until the node has stepped once after construction / `reset`).  The sum enters the
target node's input variable.

Execution here: a feedback edge makes a node depend on a later node's state of the same step, so the node-by-node
whole-horizon execution of `Network._run_engine_multi` cannot express it.  The populations are advanced in lockstep
instead: one engine call (`rp_forward` with T = 1, the same CUDA kernels) per differential-equation node and step, the
edge projections in between as device matmuls, autograd chaining the calls through the node states.  This is the host-
stepped path of the package -- it is there for coverage of the reference's API, not for speed (a launch sequence per step
and node instead of one persistent / tensor-core plan per horizon).
"""
from __future__ import annotations

from typing import Dict, Iterator, List, Optional, Tuple

import torch
from networkx import DiGraph

from . import _cabi as abi
from .edges import Linear
from .network import Network, _engine_call, _record_steps, _window_mean


class FeedbackNetwork(Network):
    """`Network` with feedback edges: `add_edge(..., feedback=True)` (rectipy/network.py:1196-1357)."""

    def __init__(self, dt: float, device: str = "cuda:0", **kwargs):
        super().__init__(dt, device, **kwargs)
        self._fb_graph: Optional[DiGraph] = None
        # spiking nodes: name -> (state before the node's last step, state after it); see `_visible_state`
        self._last_step: Dict[str, Tuple[torch.Tensor, torch.Tensor]] = {}

    # ---- graph bookkeeping (network.py:1204-1328) ------------------------------------------------------------
    def compile(self):
        if self._fb_graph is not None:
            for s, t in self._fb_graph.edges:                  # put the feedback edges back before re-sorting
                self.graph.add_edge(s, t, **self._fb_graph[s][t])
            self._fb_graph = None
        ffwd, fb = [], []
        for e in self.graph.edges:
            (fb if self.graph[e[0]][e[1]].get("feedback") else ffwd).append(e)
        g_fwd = DiGraph(self.graph.edge_subgraph(ffwd))
        self._fb_graph = DiGraph(self.graph.edge_subgraph(fb))
        self.graph = g_fwd
        super().compile()

    def add_edge(self, source: str, target: str, weights=None, train: Optional[str] = None, feedback: bool = False,
                 edge_attrs: dict = None, **kwargs) -> Linear:
        """As `Network.add_edge`; `feedback=True` keeps the edge out of the feed-forward path that connects the network input
        to its output (network.py:1230-1266)."""
        edge_attrs = dict(edge_attrs) if edge_attrs else {}
        edge_attrs["feedback"] = bool(feedback)
        kwargs.pop("dtype", None)
        return super().add_edge(source, target, weights=weights, train=train, edge_attrs=edge_attrs, **kwargs)

    def get_edge(self, source: str, target: str) -> Linear:
        try:
            return super().get_edge(source, target)
        except KeyError:
            if self._fb_graph is None:
                raise
            return self._fb_graph[source][target]["edge"]

    def get_node(self, node: str):
        try:
            return super().get_node(node)
        except KeyError:
            if self._fb_graph is None:
                raise
            return self._fb_graph.nodes[node]["node"]

    def parameters(self, recurse: bool = True) -> Iterator:
        """Trainable parameters of both graphs (network.py:1308-1325); a node that appears in both is listed once."""
        seen = set()
        for g in (self.graph, self._fb_graph):
            if g is None:
                continue
            for p in self._get_parameters(g, recurse=recurse):
                if id(p) not in seen:
                    seen.add(id(p))
                    yield p

    def _get_parameters(self, g: DiGraph, recurse: bool = True) -> Iterator:
        for node in g:
            for p in g.nodes[node]["node"].parameters(recurse=recurse):
                yield p
        for s, t in g.edges:
            for p in g[s][t]["edge"].parameters():
                yield p

    # ---- execution -------------------------------------------------------------------------------------------
    def _is_multi(self) -> bool:
        return True

    def _feedback_sources(self, name: str) -> List[str]:
        if self._fb_graph is None or name not in self._fb_graph:
            return []
        return list(self._fb_graph.predecessors(name))

    def _visible_state(self, name: str) -> torch.Tensor:
        """Engine-layout state [n_sv, B, n] that the reference's `node.y` would show right now (module docstring)."""
        node = self.get_node(name)
        if getattr(node, "spiking", False):
            last = self._last_step.get(name)
            if last is not None and last[1] is node.state:        # stepped by this network and not reset / re-assigned since
                return last[0]
        return node.state

    def detach(self, requires_grad: bool = True, detach_params: bool = False) -> None:
        before = {name: (pre, self.get_node(name).state is post) for name, (pre, post) in self._last_step.items()}
        super().detach(requires_grad=requires_grad, detach_params=detach_params)
        for name, (pre, live) in before.items():
            if live:
                self._last_step[name] = (pre.detach(), self.get_node(name).state)
            else:
                self._last_step.pop(name, None)

    def _feedback_input(self, source: str, target: str) -> torch.Tensor:
        """`edge.forward(get_node(source)["out"])` (network.py:1354-1357) for all trials: [B, n_target]."""
        node = self.get_node(source)
        if not hasattr(node, "state"):
            raise NotImplementedError("rectipy_b200: a feedback edge must start at a differential-equation node (the reference "
                                      "reads the source node's state vector, which a function node does not have).")
        if node.out_var == abi.RP_VAR_R:
            raise NotImplementedError("rectipy_b200: the source of a feedback edge must output a state variable")
        edge = self._fb_graph[source][target]["edge"]
        if getattr(edge, "stateful", False):
            raise NotImplementedError("rectipy_b200: delay / filter edges cannot be feedback edges")
        y_out = self._visible_state(source)[node.var_index(node._out_key)]              # [B, n]
        return y_out @ edge.effective_weights().T

    def _step(self, xt: torch.Tensor, want: Dict[str, list]) -> Tuple[torch.Tensor, Dict[Tuple[str, int, int], torch.Tensor]]:
        """One network step (network.py:1330-1352).  xt: [B, width].  `want`: node -> [(state plane, reduce)] to record.
        Returns the output node's output [B, k] and the requested records ([B, n] or [B])."""
        path = self._get_path()
        val = xt
        got: Dict[Tuple[str, int, int], torch.Tensor] = {}
        for i, name in enumerate(path):
            if i > 0:
                edge = Network.get_edge(self, path[i - 1], name)
                if getattr(edge, "stateful", False):
                    raise NotImplementedError("rectipy_b200: delay / filter edges are not supported inside a FeedbackNetwork")
                val = val @ edge.effective_weights().T
            for src in self._feedback_sources(name):
                val = val + self._feedback_input(src, name)
            node = self.get_node(name)
            if self[name]["node_type"] == "diff_eq":
                recs_wanted = want.get(name, [])
                if val.shape[-1] != node.n:
                    val = val.expand(val.shape[0], node.n)
                pre_state = node.state
                out, recs = _engine_call(node, val.reshape(1, node.batch, node.n).contiguous(), abi.RP_IN_DENSE, None,
                                         abi.RP_OUT_DENSE, None, 1, 1, 0, 0, tuple(v for v, _ in recs_wanted),
                                         tuple(r for _, r in recs_wanted), True)
                if getattr(node, "spiking", False):
                    self._last_step[name] = (pre_state, node.state)
                for (vi, red), r in zip(recs_wanted, recs):
                    got[(name, vi, red)] = r[0]
                val = out[0]
            else:
                val = node.apply_batched(val)
        return val, got

    def _run_engine_multi(self, x: torch.Tensor, S: int, cutoff: int, truncate: int, rec_specs: list, want_out: bool):
        """x [T,B,width] -> (out [n_rec,B,k] | None, [recorded vars], record steps); same contract as the base class."""
        T = x.shape[0]
        steps = _record_steps(T, S, cutoff)
        is_rec = set(steps)
        want: Dict[str, list] = {}
        for name, vi, red in rec_specs:
            want.setdefault(name, []).append((vi, red))
        outs: List[torch.Tensor] = []
        rec_series: Dict[Tuple[str, int, int], List[torch.Tensor]] = {key: [] for key in rec_specs}
        for t in range(T):
            out, got = self._step(x[t], want if t in is_rec else {})
            if want_out:
                outs.append(out)
            for key, r in got.items():
                rec_series[key].append(r)
            if truncate and t % truncate == truncate - 1:               # network.py:598-599
                self.detach()
        out = _window_mean(torch.stack(outs), S, cutoff) if (want_out and outs) else None
        recs = []
        for key in rec_specs:
            name, vi, red = key
            node = self.get_node(name)
            if rec_series[key]:
                recs.append(torch.stack(rec_series[key]))
            else:
                shape = (0, node.batch) if red else (0, node.batch, node.n)
                recs.append(torch.zeros(shape, dtype=torch.float32, device=x.device))
        return out, recs, steps
