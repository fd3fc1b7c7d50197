"""World-size-2 gloo tests (CPU) of the trial-parallel plumbing: sharding and the gradient all-reduce."""
import os
import socket
import sys

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank))
    from rectipy_b200 import parallel
    r, lr, w = parallel.init_from_env("gloo")
    assert (r, w) == (rank, world)
    n_trials = 7
    lo, hi = parallel.shard_trials(n_trials, rank, world)
    # toy "loss": per-trial quadratic in two parameter tensors; local loss is the MEAN over local trials
    torch.manual_seed(0)
    data = torch.randn(n_trials, 5, dtype=torch.float64)
    Wp = torch.ones(5, 5, dtype=torch.float64, requires_grad=True)
    bp = torch.zeros(5, dtype=torch.float64, requires_grad=True)
    local = ((data[lo:hi] @ Wp + bp) ** 2).sum(1).mean()
    local.backward()
    parallel.allreduce_gradients([Wp, bp], hi - lo, n_trials, bucket_bytes=64)      # tiny bucket -> several collectives
    W2 = torch.ones(5, 5, dtype=torch.float64, requires_grad=True)
    b2 = torch.zeros(5, dtype=torch.float64, requires_grad=True)
    ((data @ W2 + b2) ** 2).sum(1).mean().backward()
    ok = torch.allclose(Wp.grad, W2.grad, atol=1e-12) and torch.allclose(bp.grad, b2.grad, atol=1e-12)
    tot = parallel.allreduce_scalar(torch.tensor([float(hi - lo)]))
    ok = ok and float(tot) == n_trials
    ok = ok and parallel.global_trial_count(hi - lo) == n_trials
    # a parameter without a gradient on ONE rank still takes part (zeros), so the collectives' sizes match
    qa = torch.ones(40, dtype=torch.float64, requires_grad=True)
    qb = torch.ones(3, dtype=torch.float32, requires_grad=True)
    if rank == 0:
        (qa.sum() * 2.0).backward()
    (qb.sum() * float(rank + 1)).backward()
    h = parallel.allreduce_gradients([qa, qb], bucket_bytes=64, async_op=True)
    h.wait()
    ok = ok and torch.allclose(qa.grad, torch.full((40,), 2.0, dtype=torch.float64)) and torch.allclose(qb.grad, torch.full((3,), 3.0))

    # Network._bptt_step (rectipy/network.py:1123-1130) with trials sharded over the ranks == one process on the global batch
    import rectipy_b200 as rp

    class Stub(rp.Network):
        def __init__(self, batch, params):
            self.batch, self._params = batch, params

        def _engine_device(self):
            return torch.device("cpu")

        def parameters(self, recurse=True):
            return iter(self._params)

    tgt = torch.randn(n_trials, 5, dtype=torch.float64)
    Wl = torch.ones(5, 5, dtype=torch.float64, requires_grad=True)
    net = Stub(hi - lo, [Wl])
    opt = torch.optim.SGD([Wl], lr=0.1)
    err = net._bptt_step(data[lo:hi] @ Wl, tgt[lo:hi], optimizer=opt, loss=torch.nn.MSELoss(), error_kwargs={}, step_kwargs={})
    Wg = torch.ones(5, 5, dtype=torch.float64, requires_grad=True)
    og = torch.optim.SGD([Wg], lr=0.1)
    lg = torch.nn.MSELoss()(data @ Wg, tgt)
    lg.backward(); og.step()
    ok = ok and torch.allclose(Wl, Wg, atol=1e-12) and abs(err - float(lg)) < 1e-12
    ret[rank] = bool(ok)
    dist.barrier()
    dist.destroy_process_group()


def test_shard_trials_partition():
    from rectipy_b200.parallel import shard_trials
    for n in (1, 7, 8, 1024, 8192, 5):
        for world in (1, 2, 3, 4, 8):
            ranges = [shard_trials(n, r, world) for r in range(world)]
            assert ranges[0][0] == 0 and ranges[-1][1] == n
            assert all(ranges[i][1] == ranges[i + 1][0] for i in range(world - 1))
            sizes = [b - a for a, b in ranges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_trials(4, 2, 2)


def test_gradient_allreduce_world2_gloo():
    world, port = 2, _free_port()
    mgr = mp.Manager()
    ret = mgr.dict()
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert dict(ret) == {0: True, 1: True}
