"""Edges between nodes -- host-side mirror of rectipy/edges.py (Linear, LinearMasked, RLS).

Inside a compiled `Network` the projections are not executed here: `Network.compile` hands the edge weights to the
engine, which fuses `W_in @ x` and `W_out @ y` into the step / observer / adjoint kernels (rp_kernels.cuh).  The
`forward` methods below exist for the reference's stand-alone edge use (rectipy_tests/test_edges.py) and operate on
whatever device the weights live on.  The stateful edges (LinearMemory, LinearFilter, LinearMemoryFilter: a delay buffer /
a linear filter in front of the projection) are maps of the edge's input SERIES; `Network` applies them to the per-step
series between two whole-horizon engine calls (`apply_series`), with the reference's per-step semantics.
"""
from __future__ import annotations

from typing import Iterator, Optional, Union

import numpy as np
import torch


def _to_tensor(w, dtype) -> torch.Tensor:
    if isinstance(w, np.ndarray):
        return torch.tensor(w, dtype=dtype)
    return w.detach().clone().to(dtype)


class Linear:
    """`weights @ x` with weights stored `[n_out, n_in]` (rectipy/edges.py:8-65)."""

    _tensors = ["weights"]
    stateful = False      # True: the edge carries state across steps (delay / filter edges) and maps a whole input series

    def __init__(self, n_in: int, n_out: int, weights: Union[np.ndarray, torch.Tensor] = None,
                 dtype: torch.dtype = torch.float64, detach: bool = True, **kwargs):
        if weights is None:
            weights = torch.randn(n_out, n_in, dtype=dtype)
        else:
            weights = _to_tensor(weights, dtype)
        if weights.dim() != 2:
            raise ValueError("Edge weights have to be a 2D array.")
        if weights.shape[0] == n_in and weights.shape[1] == n_out:
            # a view, like the reference (edges.py:22-23).  NB: for a square matrix this branch is always taken, so a user's
            # N x N `weights` acts as `weights.T @ x` -- a reference quirk that ported scripts rely on
            # (documentation/rnn_tryout.py:23,26); pinned by tests/golden/edges_square.npz
            weights = weights.T
        elif weights.shape[0] != n_out or weights.shape[1] != n_in:
            raise ValueError("Shape of the provided weights does not match the input and output dimensions of the "
                             "source and target nodes.")
        self.n_in = n_in
        self.n_out = n_out
        self.weights = weights
        self.train_params = []
        if not detach:
            train_params = kwargs.pop("train_params", ["weights"])
            for key in self._tensors:
                if key in train_params:
                    val = getattr(self, key)
                    val.requires_grad_(True)
                    self.train_params.append(val)

    def __call__(self, *args, **kwargs):
        return self.forward(*args, **kwargs)

    def effective_weights(self) -> torch.Tensor:
        return self.weights

    def forward(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        return self.effective_weights() @ x

    def parameters(self, recurse: bool = True) -> Iterator:
        for p in self.train_params:
            yield p

    def to(self, device: str, **kwargs):
        for attr in self._tensors:
            val = getattr(self, attr)
            req = val.requires_grad
            new = val.detach().to(device)
            if req:
                new.requires_grad_(True)
                self.train_params = [new if p is val else p for p in self.train_params]
            setattr(self, attr, new)
        return self

    def detach(self):
        pass   # the reference's Linear.detach is a no-op (edges.py:62-65)


class _SeriesEdge(Linear):
    """Stateful edge: `forward(x)` advances the state by one step like the reference's edge; `apply_series` does so for a
    whole `[T, B, n_in]` series (state kept per trial, created on first use) and returns `[T, B, n_out]`."""

    stateful = True

    def _state_step(self, x: torch.Tensor) -> torch.Tensor:      # x [B, n_in] -> pre-projection vector [B, n_in]
        raise NotImplementedError

    def _expand_state(self, B: int) -> None:
        raise NotImplementedError

    def forward(self, x: torch.Tensor, **kwargs) -> torch.Tensor:
        single = x.dim() == 1
        xb = x.reshape(1, -1) if single else x
        self._expand_state(xb.shape[0])
        z = self._state_step(xb)
        out = z @ self.weights.T
        return out[0] if single else out

    def apply_series(self, x: torch.Tensor) -> torch.Tensor:
        self._expand_state(x.shape[1])
        w = self.weights.to(x.dtype)
        return torch.stack([self._state_step(x[t]) @ w.T for t in range(x.shape[0])])


class LinearMemory(_SeriesEdge):
    """Delay buffer in front of the projection (rectipy/edges.py:68-94).  Per step the reference rolls its
    `[n_in, max_delay+1]` buffer one column to the left, writes the input vector into the columns `delays` of EVERY row
    (`buffer[:, delays] = x` broadcasts over rows), and projects column 0.  That literal behaviour is reproduced: column
    `delays[j]` of all rows receives `x[j]` (the last channel wins when two channels share a delay)."""

    _tensors = ["weights", "buffer", "delays"]

    def __init__(self, n_in: int, n_out: int, delays: Union[np.ndarray, torch.Tensor],
                 weights: Union[np.ndarray, torch.Tensor] = None, intrinsic_weights=None,
                 dtype: torch.dtype = torch.float64, detach: bool = True, **kwargs):
        if isinstance(delays, np.ndarray):
            delays = torch.tensor(delays, dtype=torch.long)
        if len(delays) != n_in:
            raise ValueError("The number of delays must match the number of node inputs.")
        self.delays = delays.to(torch.long)
        self.buffer = torch.zeros((n_in, int(self.delays.max()) + 1), dtype=dtype)
        train_params = kwargs.pop("train_params", ["weights"])
        super().__init__(n_in, n_out, weights=weights, dtype=dtype, detach=detach, train_params=train_params, **kwargs)

    def _writers(self):
        """(written columns, channel whose value lands there): duplicates resolved like a sequential assignment (last wins)."""
        d = self.delays.tolist()
        cols = sorted(set(d))
        winners = [max(j for j, dj in enumerate(d) if dj == c) for c in cols]
        dev = self.buffer.device
        return torch.tensor(cols, dtype=torch.long, device=dev), torch.tensor(winners, dtype=torch.long, device=dev)

    def _expand_state(self, B: int) -> None:
        if self.buffer.dim() == 2:
            self.buffer = self.buffer.unsqueeze(0).repeat(B, 1, 1)
        elif self.buffer.shape[0] != B:
            raise RuntimeError(f"edge state holds {self.buffer.shape[0]} trials, input has {B}")

    def _premix(self, buf: torch.Tensor) -> torch.Tensor:
        return buf

    def _state_step(self, x: torch.Tensor) -> torch.Tensor:
        cols, winners = self._writers()
        buf = self._premix(torch.roll(self.buffer.to(x.dtype), -1, dims=2)).clone()
        buf[:, :, cols] = x[:, None, winners]
        self.buffer = buf
        return buf[:, :, 0]

    def to(self, device: str, **kwargs):
        super().to(device, **kwargs)
        return self


class LinearFilter(_SeriesEdge):
    """Linear filter in front of the projection (rectipy/edges.py:97-120):  y <- filter @ y + x ;  out = weights @ y."""

    _tensors = ["weights", "filter", "y"]

    def __init__(self, n_in: int, n_out: int, filter_weights: Union[np.ndarray, torch.Tensor],
                 weights: Union[np.ndarray, torch.Tensor] = None, dtype: torch.dtype = torch.float64,
                 detach: bool = True, **kwargs):
        if isinstance(filter_weights, np.ndarray):
            filter_weights = torch.tensor(filter_weights, dtype=dtype)
        if filter_weights.shape[0] != n_in or filter_weights.shape[1] != n_in:
            raise ValueError("Intrinsic weights have to be a square matrix with the number of rows and columns matching"
                             "the number of inputs to the edge.")
        self.filter = filter_weights.to(dtype)
        self.y = torch.zeros(n_in, dtype=dtype)
        train_params = kwargs.pop("train_params", ["weights", "filter"])
        super().__init__(n_in, n_out, weights=weights, dtype=dtype, detach=detach, train_params=train_params, **kwargs)

    def _expand_state(self, B: int) -> None:
        if self.y.dim() == 1:
            self.y = self.y.unsqueeze(0).repeat(B, 1)
        elif self.y.shape[0] != B:
            raise RuntimeError(f"edge state holds {self.y.shape[0]} trials, input has {B}")

    def _state_step(self, x: torch.Tensor) -> torch.Tensor:
        self.y = self.y.to(x.dtype) @ self.filter.to(x.dtype).T + x
        return self.y


class LinearMemoryFilter(LinearMemory):
    """Delay buffer whose rows are mixed by a filter matrix at every step (rectipy/edges.py:123-147):
    buffer <- filter @ roll(buffer) ; buffer[:, delays] = x ; out = weights @ buffer[:, 0]."""

    _tensors = ["weights", "buffer", "delays", "filter"]

    def __init__(self, n_in: int, n_out: int, delays: Union[np.ndarray, torch.Tensor],
                 filter_weights: Union[np.ndarray, torch.Tensor], weights: Union[np.ndarray, torch.Tensor] = None,
                 dtype: torch.dtype = torch.float64, detach: bool = True, **kwargs):
        if isinstance(filter_weights, np.ndarray):
            filter_weights = torch.tensor(filter_weights, dtype=dtype)
        if filter_weights.shape[0] != n_in or filter_weights.shape[1] != n_in:
            raise ValueError("Intrinsic weights have to be a square matrix with the number of rows and columns matching"
                             "the number of inputs to the edge.")
        self.filter = filter_weights.to(dtype)
        train_params = kwargs.pop("train_params", ["weights", "filter"])
        super().__init__(n_in, n_out, delays=delays, weights=weights, dtype=dtype, detach=detach,
                         train_params=train_params, **kwargs)

    def _premix(self, buf: torch.Tensor) -> torch.Tensor:
        return torch.matmul(self.filter.to(buf.dtype), buf)      # [n_in, n_in] @ [B, n_in, D+1]


class LinearMasked(Linear):
    """`(weights * mask) @ x` (rectipy/edges.py:150-174)."""

    _tensors = ["weights", "mask"]

    def __init__(self, n_in: int, n_out: int, mask: Union[np.ndarray, torch.Tensor],
                 weights: Union[np.ndarray, torch.Tensor] = None, dtype: torch.dtype = torch.float64,
                 detach: bool = True, **kwargs):
        mask = _to_tensor(mask, dtype)
        if mask.shape[0] == n_in and mask.shape[1] == n_out:      # square masks are transposed too (edges.py:160-161)
            mask = mask.T.contiguous()
        elif mask.shape[0] != n_out or mask.shape[1] != n_in:
            raise ValueError("Shape of the provided mask does not match the input and output dimensions of the "
                             "source and target nodes.")
        self.mask = mask
        train_params = kwargs.pop("train_params", ["weights"])
        super().__init__(n_in, n_out, weights=weights, dtype=dtype, detach=detach, train_params=train_params, **kwargs)

    def effective_weights(self) -> torch.Tensor:
        return self.weights * self.mask


class RLS(Linear):
    """Recursive-least-squares readout edge (rectipy/edges.py:177-234).  `update` runs on the weights' device; the
    whole-horizon training loop of `Network.fit_rls` uses the fused device path `rp_rls_run` instead."""

    _tensors = ["weights", "P"]

    def __init__(self, n_in: int, n_out: int, weights: Union[np.ndarray, torch.Tensor] = None,
                 dtype: torch.dtype = torch.float64, beta: float = 1.0, alpha: float = 1.0, **kwargs):
        if beta > 1 or beta < 0:
            raise ValueError("Parameter beta should be a positive scalar between 0 and 1.")
        if alpha < 0:
            raise ValueError("Parameter alpha should be a positive scalar.")
        if weights is None:
            weights = torch.zeros((n_out, n_in), dtype=dtype)
        self.beta = beta ** (-1)
        self.P = alpha * torch.eye(n_in, dtype=dtype)
        self.loss = 0.0
        super().__init__(n_in, n_out, weights=weights, dtype=dtype, detach=True)
        self.train_params = []

    def update(self, x: torch.Tensor, y: torch.Tensor, y_hat: torch.Tensor) -> None:
        # operation order follows edges.py:229-234 exactly so that results are bit-identical on the same device
        z = self.beta * self.P @ x
        kappa = (1.0 + x @ z) ** (-1)
        err = y - y_hat
        self.weights += torch.outer((y - kappa * x @ (self.weights + torch.outer(y, z)).T), z)
        self.P -= kappa * torch.outer(z, z)
        self.loss = torch.inner(err, err)
