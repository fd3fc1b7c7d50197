"""Trial-level data parallelism: one process per GPU, trials sharded, parameters replicated.

Trials never interact in the forward pass and time steps are strictly sequential, so the only sensible shard axis is
the trial axis.  A parameter sweep / forward simulation needs no communication at all; `fit_bptt`
needs one fp32 sum all-reduce of the parameter gradients per optimizer step (NCCL over NVLink on the GPU box, gloo in
the CPU tests).  The reference has no distributed code; this module is new functionality around the unchanged API.
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: str = None) -> Tuple[int, int, int]:
    """Initialise torch.distributed from torchrun's env (RANK/LOCAL_RANK/WORLD_SIZE/MASTER_*). Returns (rank, local_rank, world)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        if backend == "nccl":
            torch.cuda.set_device(local_rank)
            dist.init_process_group(backend=backend, device_id=torch.device("cuda", local_rank))
        else:
            dist.init_process_group(backend=backend)
    return rank, local_rank, world


def shard_trials(n_trials: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous trial range [start, stop) owned by `rank`; the first `n_trials % world` ranks get one extra trial."""
    if world <= 0 or not (0 <= rank < world):
        raise ValueError(f"bad rank/world {rank}/{world}")
    base, rem = divmod(n_trials, world)
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


class GradientSync:
    """Handles of an in-flight gradient all-reduce; `wait()` blocks the current stream until the sums have landed."""

    def __init__(self):
        self.works = []
        self.unpack = []          # (flat bucket, [grad tensors]) to scatter back after the collective

    def wait(self) -> None:
        for w in self.works:
            w.wait()
        for flat, grads in self.unpack:
            off = 0
            for g in grads:
                g.copy_(flat[off:off + g.numel()].view_as(g))
                off += g.numel()
        self.works, self.unpack = [], []


def allreduce_gradients(params: Iterable[torch.Tensor], n_trials_local: int = None, n_trials_global: int = None,
                        bucket_bytes: int = 1 << 20, async_op: bool = False):
    """Sum parameter gradients over ranks in place (fp32 sum all-reduce, NCCL over NVLink on the GPU box).

    * Every parameter takes part, in the order given, on every rank: a parameter whose `.grad` is None on this rank (its
      shard produced no dependence on it) contributes zeros, so the collectives' sizes never differ between ranks.
    * A gradient of at least `bucket_bytes` (the recurrent weights: 64 MiB at N=4096, 256 MiB at N=8192) is reduced in
      place, without staging copies; smaller ones are packed into one flat bucket per dtype.
    * If the local losses are *means* over local trials, pass the trial counts: each rank's gradient is weighted by
      n_local / n_global before the sum, so the result equals the gradient of the global mean.
    * `async_op=True` returns a `GradientSync` right after the collectives are enqueued (on the communicator's own stream);
      call `.wait()` before the optimizer step.  Otherwise the call waits itself and returns None.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return None
    params = [p for p in params if p.requires_grad]
    if not params:
        return None
    scale = None
    if n_trials_local is not None and n_trials_global is not None:
        scale = float(n_trials_local) / float(n_trials_global)
    sync = GradientSync()
    small = {}                    # dtype -> [grad]
    for p in params:
        if p.grad is None:
            p.grad = torch.zeros_like(p)
        g = p.grad
        if scale is not None:
            g.mul_(scale)
        if g.is_contiguous() and g.numel() * g.element_size() >= bucket_bytes:
            sync.works.append(dist.all_reduce(g, op=dist.ReduceOp.SUM, async_op=True))
        else:
            small.setdefault(g.dtype, []).append(g)
    for dtype, grads in small.items():
        flat = torch.cat([g.reshape(-1) for g in grads])
        sync.works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, async_op=True))
        sync.unpack.append((flat, grads))
    if async_op:
        return sync
    sync.wait()
    return None


def global_trial_count(n_trials_local: int, device=None) -> int:
    """Sum of the ranks' local trial counts (one scalar all-reduce; call once per fit, not per step)."""
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return int(n_trials_local)
    t = torch.tensor([float(n_trials_local)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return int(round(float(t.item())))


def allreduce_scalar(value: torch.Tensor, op=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(value, op=op or dist.ReduceOp.SUM)
    return value
