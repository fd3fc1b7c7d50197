#!/usr/bin/env python
"""Multi-GPU correctness of the trial-sharded BPTT path, on the engine itself.

    python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_dp_equivalence.py

World size W: every rank builds the same QIF network, owns B/W of the B trials, runs `Network.run` + loss.backward() and
`Network.allreduce_gradients()` (the hook `fit_bptt` calls before each optimizer step).  Rank 0 additionally runs all B trials
alone.  Asserts  dW, dW_out (W ranks x B/W trials)  ==  dW, dW_out (1 rank x B trials)  to <= 1e-5 relative, and that one
`fit_bptt` epoch leaves bit-identical parameters on every rank.  Prints one JSON line (committed under profiles/)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402  (problem recipe only)
from rectipy_b200 import parallel  # noqa: E402


def main():
    os.environ.setdefault("NCCL_DEBUG", "WARN")       # keep stdout to the JSON line
    rank, local_rank, world = parallel.init_from_env("nccl")
    torch.cuda.set_device(local_rank)
    dev = f"cuda:{local_rank}"
    n, B, T = 1024, 256 * world, 60
    bench.N_NEURONS = n
    W, w_in, w_out, etas, x, tgt = bench.make_problem(7, n, B, T)
    y0 = bench.spread_state(11, n, B)
    lo, hi = parallel.shard_trials(B, rank, world)

    def grads(b0, b1):
        net, node = bench.build_network(W, w_in, w_out, etas, b1 - b0, dev)
        node.reset(y0[b0:b1])
        obs = net.run(torch.tensor(x[:, b0:b1], device=dev), sampling_steps=1, verbose=False, enable_grad=True)
        loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]), torch.tensor(tgt[:, b0:b1], device=dev))
        loss.backward()
        return net, node, loss

    net, node, loss = grads(lo, hi)
    net.allreduce_gradients()
    gW, gWo = node["weights"].grad.clone(), net.get_edge("qif", "out").weights.grad.clone()
    res = {"world": world, "n": n, "trials": B, "T": T}
    if rank == 0:
        # the single-rank run must not see the process group
        net1, node1, loss1 = grads(0, B)
        rW, rWo = node1["weights"].grad, net1.get_edge("qif", "out").weights.grad
        eW = float((gW - rW).abs().max() / rW.abs().max())
        eWo = float((gWo - rWo).abs().max() / rWo.abs().max())
        res.update(dW_rel_err=eW, dW_out_rel_err=eWo, dW_max=float(rW.abs().max()))
    # one fit_bptt epoch: replicas must stay bit-identical
    net2, node2 = bench.build_network(W, w_in, w_out, etas, hi - lo, dev)
    node2.reset(y0[lo:hi])
    net2.fit_bptt([x[:, lo:hi]], [tgt[:, lo:hi]], optimizer="sgd", lr=1e-2, sampling_steps=1, verbose=False)
    flat = torch.cat([node2["weights"].detach().reshape(-1), net2.get_edge("qif", "out").weights.detach().reshape(-1)])
    ref = flat.clone()
    dist.broadcast(ref, src=0)
    same = torch.tensor([float(torch.equal(flat, ref))], device=dev)
    dist.all_reduce(same, op=dist.ReduceOp.MIN)
    moved = float((node2["weights"].detach() - torch.tensor(W, device=dev)).abs().max())
    if rank == 0:
        res.update(replicas_bit_identical_after_fit_bptt=bool(same.item() == 1.0), weights_moved_by=moved)
        ok = res["dW_rel_err"] <= 1e-5 and res["dW_out_rel_err"] <= 1e-5 and res["replicas_bit_identical_after_fit_bptt"] and moved > 0
        res["ok"] = bool(ok)
        print(json.dumps(res))
    dist.barrier(device_ids=[local_rank])
    dist.destroy_process_group()
    if rank == 0 and not res["ok"]:
        sys.exit(1)


if __name__ == "__main__":
    main()
