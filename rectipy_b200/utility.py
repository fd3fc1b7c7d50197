"""Host-side helpers (one-off set-up work, not accelerated): connectivity generators, normalisation, scores.

Same call signatures and semantics as rectipy/utility.py so that example scripts keep working.
"""
from __future__ import annotations

from typing import Union

import numpy as np
import torch


def retrieve_from_dict(keys: list, data: dict) -> dict:
    """Pop `keys` out of `data` into a new dict."""
    return {k: data.pop(k) for k in keys if k in data}


def add_op_name(op: str, var: Union[str, None], new_var_names: dict) -> Union[str, None]:
    """Prefix a variable identifier with its operator name (rectipy/utility.py:32-56)."""
    if var is None or var == "weights":
        return var
    if "/" in var:
        short = var.split("/")[-1]
        new_var_names[short] = var
        return var
    new_var_names[var] = f"{op}/{var}"
    return new_var_names[var]


def to_device(x, device: str):
    return x.to(device) if hasattr(x, "to") else x


def _wrap(idxs: np.ndarray, N: int) -> np.ndarray:
    return np.mod(idxs, N)


def _ring_or_line(N, p, spatial_distribution, homogeneous_weights, wrap: bool) -> np.ndarray:
    C = np.zeros((N, N))
    n_conns = int(N * p)
    for n in range(N):
        dist = np.asarray(spatial_distribution.rvs(size=n_conns))
        signs = np.where(np.random.rand(n_conns) < 0.5, -1, 1)
        conns = n + dist * signs
        if wrap:
            conns = _wrap(conns, N)
        else:
            conns = conns[(conns > 0) & (conns < N)]
        uniq, counts = np.unique(conns, return_counts=True)
        if len(uniq) == 0:
            continue
        if homogeneous_weights:
            C[n, uniq] = 1.0 / len(uniq)
        else:
            C[n, uniq] = counts / (n_conns if wrap else len(conns))
    return C


def circular_connectivity(N: int, p: float, spatial_distribution, homogeneous_weights: bool = True) -> np.ndarray:
    """Coupling matrix of nodes on a ring; distances drawn from `spatial_distribution` (scipy rv_discrete)."""
    return _ring_or_line(N, p, spatial_distribution, homogeneous_weights, wrap=True)


def line_connectivity(N: int, p: float, spatial_distribution, homogeneous_weights: bool = True) -> np.ndarray:
    """Coupling matrix of nodes on a line (no wrap-around)."""
    return _ring_or_line(N, p, spatial_distribution, homogeneous_weights, wrap=False)


def random_connectivity(n: int, m: int, p: float, normalize: bool = True) -> np.ndarray:
    """Each row gets int(m*p) randomly placed entries of 1/int(m*p) (or 1)  (rectipy/utility.py:153-178)."""
    C = np.zeros((n, m))
    n_conns = int(m * p)
    for row in range(n):
        cols = np.random.permutation(m)[:n_conns]
        C[row, cols] = 1.0 / n_conns if normalize else 1.0
    return C


def input_connections(n: int, m: int, p: float, variance: float = 1.0, zero_mean: bool = True) -> np.ndarray:
    """Random sparse input weights with given row variance, optionally zero row mean."""
    C_tmp = random_connectivity(n, m, p, normalize=False)
    C = np.zeros_like(C_tmp)
    for row in range(n):
        idx = np.flatnonzero(C_tmp[row] > 0)
        if len(idx) == 0:
            continue
        w = np.random.randn(len(idx)) * np.sqrt(variance)
        if zero_mean and len(idx) > 1:
            w -= np.mean(w)
        C[row, idx] = w
    return C


def normalize(x: np.ndarray, mode: str = "minmax", row_wise: bool = False) -> np.ndarray:
    """`minmax` -> [0,1], `zscore` -> zero mean / unit variance, `sum` -> unit sum; optionally per row."""
    x = np.array(x, dtype=float, copy=True)
    if row_wise:
        for i in range(x.shape[0]):
            x[i] = normalize(x[i], mode=mode, row_wise=False)
        return x
    if mode == "minmax":
        x -= np.min(x)
        mx = np.max(x)
        if mx > 0:
            x /= mx
    elif mode == "zscore":
        x -= np.mean(x)
        sd = np.std(x)
        if sd > 0:
            x /= sd
    elif mode == "sum":
        sm = np.sum(x)
        if sm != 0:
            x /= sm
    else:
        raise ValueError(f"Invalid normalization mode: {mode}")
    return x


def wta_score(x: np.ndarray, y: np.ndarray) -> float:
    """Winner-takes-all score: fraction of samples whose argmax agrees."""
    return float(np.mean(np.argmax(x, axis=1) == np.argmax(y, axis=1)))


def readout(X: np.ndarray, y: np.ndarray, k: int = 1, verbose: bool = True, **kwargs):
    """Ridge readout with k-fold cross validation on host arrays (convenience, scikit-learn backed)."""
    from sklearn.linear_model import Ridge
    from sklearn.model_selection import StratifiedKFold
    clf = Ridge(**kwargs)
    if k > 1:
        scores, coefs = [], []
        for train, test in StratifiedKFold(n_splits=k).split(X, np.argmax(y, axis=1) if y.ndim > 1 else y):
            clf.fit(X[train], y[train])
            scores.append(clf.score(X[test], y[test]))
            coefs.append(clf.coef_)
        if verbose:
            print(f"Average score: {np.mean(scores)}")
        return np.mean(scores), np.mean(coefs, axis=0)
    clf.fit(X, y)
    return clf.score(X, y), clf.coef_
