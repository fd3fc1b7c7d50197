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


def allreduce_gradients(params: Iterable[torch.Tensor], n_trials_local: int = None, n_trials_global: int = None,
                        bucket_bytes: int = 256 << 20) -> None:
    """Sum parameter gradients over ranks in place (flat fp32 buckets, one collective per bucket).

    If the local losses are *means* over local trials, pass the trial counts: each rank's gradient is weighted by
    n_local / n_global before the sum, so the result equals the gradient of the global mean.
    """
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return
    grads = [p.grad for p in params if p.grad is not None]
    if not grads:
        return
    scale = None
    if n_trials_local is not None and n_trials_global is not None:
        scale = float(n_trials_local) / float(n_trials_global)
    bucket: List[torch.Tensor] = []
    size = 0

    def flush():
        nonlocal bucket, size
        if not bucket:
            return
        flat = torch.cat([g.reshape(-1) for g in bucket])
        if scale is not None:
            flat.mul_(scale)
        dist.all_reduce(flat, op=dist.ReduceOp.SUM)
        off = 0
        for g in bucket:
            g.copy_(flat[off:off + g.numel()].view_as(g))
            off += g.numel()
        bucket, size = [], 0

    for g in grads:
        nbytes = g.numel() * g.element_size()
        if size and size + nbytes > bucket_bytes:
            flush()
        bucket.append(g)
        size += nbytes
    flush()


def allreduce_scalar(value: torch.Tensor, op=None) -> torch.Tensor:
    if dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(value, op=op or dist.ReduceOp.SUM)
    return value
