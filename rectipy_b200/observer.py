"""Observer -- recordings of a run (mirror of rectipy/observer.py).

The engine records on the device at the sample rate (rp_kernels.cuh: k_observe) into `[n_rec, ...]` tensors; this
class exposes them through the reference's interface: `obs["out"]` is a list of per-sample tensors (so
`torch.stack(obs["out"])` keeps working and stays differentiable), `to_numpy`, `to_dataframe`, `save`.
"""
from __future__ import annotations

from typing import Any, Iterable, Tuple, Union

import numpy as np
import torch


class Observer:

    def __init__(self, dt: float, record_output: bool = True, record_loss: bool = True, record_vars: list = None):
        if not record_vars:
            record_vars = []
        self._dt = dt
        self._state_vars = [tuple(v[:2]) for v in record_vars]
        self._reduce_vars = [bool(v[2]) if len(v) > 2 else False for v in record_vars]
        self._recordings = {v: [] for v in self._state_vars}
        self._record_loss = record_loss
        self._record_out = record_output
        if record_loss:
            self._recordings["loss"] = []
        if record_output:
            self._recordings["out"] = []
        self._recordings["steps"] = []
        self._additional_storage = {}

    def __getitem__(self, item: Union[str, Tuple[str, str]]):
        if isinstance(item, list):
            item = tuple(item)
        try:
            return self._recordings[item]
        except KeyError:
            return self._additional_storage[item]

    @property
    def recorded_state_variables(self) -> list:
        return self._state_vars

    @property
    def reduce_flags(self) -> list:
        return self._reduce_vars

    @property
    def recorded_variables(self) -> list:
        return list(self._recordings.keys())

    @property
    def recordings(self):
        from pandas import DataFrame
        columns = list(self._state_vars)
        if self._record_out:
            columns.append("out")
        if self._record_loss:
            columns.append("loss")
        data = np.asarray([self.to_numpy(v) for v in columns]).T
        return DataFrame(index=np.asarray(self._recordings["steps"]) * self._dt, data=data, columns=columns)

    def to_dataframe(self, item: Union[str, Tuple[str, str]]):
        from pandas import DataFrame
        try:
            data = self.to_numpy(item)
            if data.ndim > 2:
                data = data.reshape(data.shape[0], -1)
            return DataFrame(index=np.asarray(self._recordings["steps"]) * self._dt, data=data)
        except (KeyError, AttributeError):
            return self[item]

    def record(self, step: int, output: torch.Tensor, loss: Union[float, torch.Tensor],
               record_vars: Iterable[torch.Tensor]) -> None:
        """Single recording step (rectipy/observer.py:79-105)."""
        recs = self._recordings
        recs["steps"].append(step)
        for key, val, reduce in zip(self._state_vars, record_vars, self._reduce_vars):
            recs[key].append(torch.mean(val) if reduce else val)
        if self._record_out:
            recs["out"].append(output)
        if self._record_loss:
            recs["loss"].append(loss)

    def record_block(self, steps, out: torch.Tensor = None, loss=None, rec_vars: Iterable[torch.Tensor] = ()) -> None:
        """Append a whole device-side recording block: `out` [n_rec, ...], each var [n_rec, ...]."""
        steps = [int(s) for s in steps]
        self._recordings["steps"].extend(steps)
        if self._record_out and out is not None:
            self._recordings["out"].extend(out.unbind(0))
        for key, val in zip(self._state_vars, rec_vars):
            self._recordings[key].extend(val.unbind(0))
        if self._record_loss:
            if loss is None:
                self._recordings["loss"].extend([0.0] * len(steps))
            elif isinstance(loss, torch.Tensor) and loss.dim() > 0:
                self._recordings["loss"].extend(loss.unbind(0))
            elif isinstance(loss, (list, tuple)):
                self._recordings["loss"].extend(loss)
            else:
                self._recordings["loss"].extend([loss] * len(steps))

    def save(self, key: str, val: Any):
        self._additional_storage[key] = val

    def to_numpy(self, item: Union[str, Tuple[str, str]]) -> np.ndarray:
        try:
            val = self._recordings[item]
        except KeyError:
            val = self._additional_storage[item]
        if isinstance(val, torch.Tensor):
            return val.detach().cpu().numpy()
        if len(val) > 0 and isinstance(val[0], torch.Tensor):
            return torch.stack([v.detach() for v in val]).cpu().numpy()     # one device->host copy, not one per sample
        return np.asarray(val)

    def plot(self, y, x=None, ax=None, **kwargs):
        import matplotlib.pyplot as plt   # optional dependency, plotting only
        if ax is None:
            _, ax = plt.subplots(**{k: kwargs.pop(k) for k in ["figsize"] if k in kwargs})
        if x is None:
            ax.plot(self.to_dataframe(y), **kwargs)
        else:
            ax.plot(self.to_numpy(x), self.to_numpy(y), **kwargs)
        ax.set_xlabel("time" if x is None else str(x))
        ax.set_ylabel(str(y))
        return ax

    def matshow(self, v, ax=None, **kwargs):
        import matplotlib.pyplot as plt
        if ax is None:
            _, ax = plt.subplots(**{k: kwargs.pop(k) for k in ["figsize"] if k in kwargs})
        sig = np.asarray(self.to_dataframe(v))
        shrink = kwargs.pop("shrink", 0.6)
        im = ax.imshow(sig.T, **kwargs)
        plt.colorbar(im, ax=ax, shrink=shrink)
        ax.set_xlabel("time")
        ax.set_ylabel(str(v))
        return ax
