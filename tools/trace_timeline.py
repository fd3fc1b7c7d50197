"""CTA timeline of one BPTT pass of the headline workload (GPU): which kernels share the SMs, and when.

    python tools/trace_timeline.py [--t-inner 24] [--out gpurun_out/timeline.txt]

Uses the rp_trace_* debug hooks of the C ABI: every traced CTA records (tag, smid, start, end) from %globaltimer.
Prints, per kernel launch (a run of records with the same tag, ordered by start time), its start / end relative to the
first record, the number of CTAs and of distinct SMs, and how many of its CTAs ran on an SM while a CTA of another traced
kernel was resident there (co-residency)."""
from __future__ import annotations

import argparse
import ctypes as C
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
TAGS = {1: "gemm_store(Z)", 2: "gemm_wgrad_slice", 3: "gemm_fwd", 4: "gemm_adj", 5: "adj_step", 6: "adj_convert"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--t-inner", type=int, default=24)
    ap.add_argument("--out", default=None)
    args = ap.parse_args()
    import torch
    import bench
    from rectipy_b200 import _cabi as abi
    n, B, T = bench.N_NEURONS, bench.BATCH, args.t_inner
    lib = abi.load()
    dev = "cuda:0"
    W, w_in, w_out, etas, x_np, tgt_np = bench.make_problem(1234, n, B, T)
    net, node = bench.build_network(W, w_in, w_out, etas, B, dev)
    node.reset(bench.spread_state(4321, n, B))
    y0 = net.state
    x = torch.tensor(x_np, device=dev)
    tgt = torch.tensor(tgt_np, device=dev)
    params = [node["weights"], net.get_edge("qif", "out").weights]

    def step():
        net.reset(y0)
        for p in params:
            p.grad = None
        obs = net.run(x, sampling_steps=1, verbose=False, enable_grad=True)
        loss = torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt)
        loss.backward()
        torch.cuda.synchronize()

    for _ in range(2):
        step()
    cap = 400000
    abi.check(lib.rp_trace_enable(cap), "rp_trace_enable")
    step()
    dt = np.dtype([("tag", "<u4"), ("smid", "<u4"), ("t0", "<u8"), ("t1", "<u8")])
    buf = np.zeros(cap, dtype=dt)
    n = lib.rp_trace_read(buf.ctypes.data_as(C.c_void_p), cap)
    lib.rp_trace_enable(0)
    rec = buf[:n]
    rec = rec[rec["t1"] > 0]
    rec = rec[np.argsort(rec["t0"], kind="stable")]
    t_base = rec["t0"].min()
    lines = [f"{n} CTA records, span {(rec['t1'].max() - t_base) / 1e3:.1f} us"]
    # launches: maximal runs of one tag whose CTAs start before the previous run of the same tag ended + 3 us gap rule
    launches = []
    for tag in np.unique(rec["tag"]):
        r = rec[rec["tag"] == tag]
        start = 0
        for i in range(1, len(r) + 1):
            if i == len(r) or r["t0"][i] > r["t1"][start:i].max() + 2000:
                launches.append((int(tag), r[start:i]))
                start = i
    launches.sort(key=lambda lr: lr[1]["t0"].min())
    # co-residency: for every CTA, was a CTA with another tag resident on the same SM during its lifetime?
    by_sm = {}
    for sm in np.unique(rec["smid"]):
        by_sm[int(sm)] = rec[rec["smid"] == sm]
    lines.append(f"{'kernel':18s} {'start us':>9s} {'end us':>9s} {'dur us':>8s} {'CTAs':>6s} {'SMs':>4s} {'co-resident CTAs':>17s}")
    for tag, r in launches:
        co = 0
        for c in r:
            o = by_sm[int(c["smid"])]
            if np.any((o["tag"] != tag) & (o["t0"] < c["t1"]) & (o["t1"] > c["t0"])):
                co += 1
        lines.append(f"{TAGS.get(tag, str(tag)):18s} {(r['t0'].min() - t_base) / 1e3:9.1f} {(r['t1'].max() - t_base) / 1e3:9.1f} "
                     f"{(r['t1'].max() - r['t0'].min()) / 1e3:8.1f} {len(r):6d} {len(np.unique(r['smid'])):4d} {co:17d}")
    # per-CTA detail of a few launches in the middle of the forward and reverse sweeps: start / end quantiles relative to the launch start
    for want in (3, 1, 5, 6):
        sel = [r for tag, r in launches if tag == want]
        if not sel:
            continue
        r = sel[len(sel) // 2]
        b = r["t0"].min()
        q = lambda a: " ".join(f"{v / 1e3:7.1f}" for v in np.quantile(a - b, [0, 0.25, 0.5, 0.75, 1.0]))
        lines.append(f"{TAGS[want]:18s} CTA start q0/25/50/75/100: {q(r['t0'])} | end: {q(r['t1'])} us")
    text = "\n".join(lines)
    print(text)
    if args.out:
        with open(args.out, "w") as fh:
            fh.write(text + "\n")


if __name__ == "__main__":
    main()
