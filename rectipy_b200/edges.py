"""Edges between nodes -- host-side mirror of rectipy/edges.py (Linear, LinearMasked, RLS).

Inside a compiled `Network` the projections are not executed here: `Network.compile` hands the edge weights to the
engine, which fuses `W_in @ x` and `W_out @ y` into the step / observer / adjoint kernels (rp_kernels.cuh).  The
`forward` methods below exist for the reference's stand-alone edge use (rectipy_tests/test_edges.py) and operate on
whatever device the weights live on.  Delay/filter edges (LinearMemory, LinearFilter, LinearMemoryFilter) are out of
scope of the engine and raise at construction time via `Network.add_edge`.
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

    def __init__(self, n_in: int, n_out: int, weights: Union[np.ndarray, torch.Tensor] = None,
                 dtype: torch.dtype = torch.float64, detach: bool = True, **kwargs):
        if weights is None:
            weights = torch.randn(n_out, n_in, dtype=dtype)
        else:
            weights = _to_tensor(weights, dtype)
        if weights.dim() != 2:
            raise ValueError("Edge weights have to be a 2D array.")
        if weights.shape[0] == n_in and weights.shape[1] == n_out and n_in != n_out:
            weights = weights.T          # a view, like the reference (edges.py:22-23): keeps results bit-identical to it
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


class LinearMasked(Linear):
    """`(weights * mask) @ x` (rectipy/edges.py:150-174)."""

    _tensors = ["weights", "mask"]

    def __init__(self, n_in: int, n_out: int, mask: Union[np.ndarray, torch.Tensor],
                 weights: Union[np.ndarray, torch.Tensor] = None, dtype: torch.dtype = torch.float64,
                 detach: bool = True, **kwargs):
        mask = _to_tensor(mask, dtype)
        if mask.shape[0] == n_in and mask.shape[1] == n_out and n_in != n_out:
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
