"""Experiment (GPU): does running the batch as G independent trial groups on G streams hide the fused epilogue's HBM burst?
Aggregate forward (and BPTT) neuron-steps/s of G networks with 1024/G trials each vs one network with 1024 trials."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import bench

T = int(os.environ.get("EXP_T", "100"))
n = bench.N_NEURONS
dev = "cuda:0"


def build(B, seed):
    W, w_in, w_out, etas, x_np, tgt_np = bench.make_problem(1234 + seed, n, B, T)
    net, node = bench.build_network(W, w_in, w_out, etas, B, dev)
    node.reset(bench.spread_state(4321 + seed, n, B))
    return dict(net=net, node=node, y0=net.state, x=torch.tensor(x_np, device=dev), tgt=torch.tensor(tgt_np, device=dev),
                params=[node["weights"], net.get_edge("qif", "out").weights])


def run(groups, grad, streams):
    for g, st in zip(groups, streams):
        with torch.cuda.stream(st):
            g["net"].reset(g["y0"])
            for p in g["params"]:
                p.grad = None
            obs = g["net"].run(g["x"], sampling_steps=1, verbose=False, enable_grad=grad)
            if grad:
                torch.nn.functional.mse_loss(torch.stack(obs["out"]), g["tgt"]).backward()


for G in (1, 2, 4):
    B = 1024 // G
    groups = [build(B, i) for i in range(G)]
    streams = [torch.cuda.Stream() for _ in range(G)]
    for grad in (False, True):
        for _ in range(2):
            run(groups, grad, streams)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        reps = 3
        for _ in range(reps):
            run(groups, grad, streams)
        torch.cuda.synchronize()
        dt = (time.perf_counter() - t0) / reps
        print(f"G={G} B/group={B} grad={grad}: {dt * 1e3:.2f} ms per pass of T={T}  -> {n * 1024 * T / dt:.3e} neuron-steps/s", flush=True)
    del groups
    from rectipy_b200 import engine
    engine.clear_plans()
    torch.cuda.empty_cache()
