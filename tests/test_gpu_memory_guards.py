"""GPU: the engine never writes outside the buffers the caller hands it.

compute-sanitizer is refused on this pool, so this is the stand-in for `memcheck` on the caller-owned side of
the C ABI: every tensor `engine.EngineRun` allocates for a call (records, final state, checkpoints, every gradient) is carved out
of a larger allocation whose margins hold a canary pattern; after forward + backward on each execution path the margins must be
untouched and the payload must be fully written (no canary left inside).  Paths: persistent multi-step tcgen05 forward + fused
reverse kernel (binary16), tf32 tensor-core path, persistent few-trial kernels (1-D and 2-D grid), per-step FFMA path with ragged
sizes, mean-field template."""
import numpy as np
import pytest
import torch

from golden_util import TEMPLATE_PATH, orc

pytestmark = pytest.mark.gpu

CANARY = 0x7FC0BEEF            # a quiet-NaN bit pattern no kernel produces
MARGIN = 4096                  # elements on either side


class GuardedAlloc:
    def __init__(self):
        self.parents = []
        self._empty = torch.empty

    def empty(self, *size, **kw):
        shape = size[0] if len(size) == 1 and isinstance(size[0], (tuple, list, torch.Size)) else size
        dev = kw.get("device")
        if dev is None or torch.device(dev).type != "cuda" or kw.get("dtype", torch.float32) != torch.float32:
            return self._empty(*size, **kw)
        numel = int(np.prod(shape)) if len(shape) else 1
        parent = self._empty(numel + 2 * MARGIN, device=dev, dtype=torch.int32)
        parent.fill_(CANARY)
        view = parent[MARGIN:MARGIN + numel].view(torch.float32).view(*shape)
        self.parents.append((parent, numel, tuple(shape)))
        return view

    def empty_like(self, t, **kw):
        return self.empty(tuple(t.shape), device=kw.get("device", t.device), dtype=kw.get("dtype", t.dtype))

    def check(self):
        assert self.parents, "no guarded allocation was made"
        for parent, numel, shape in self.parents:
            lo, hi = parent[:MARGIN], parent[MARGIN + numel:]
            assert bool((lo == CANARY).all()) and bool((hi == CANARY).all()), f"write outside a caller buffer of shape {shape}"
            if len(shape) == 4:
                continue          # checkpoints: the drive plane of the last slot (ik templates) is legitimately never written
            left = int((parent[MARGIN:MARGIN + numel] == CANARY).sum())
            assert left == 0, f"{left} of {numel} elements of a caller buffer of shape {shape} were never written"


CASES = [("qif", 256, 256, "auto", 60), ("qif_sfa", 128, 128, "3xtf32", 40), ("qif", 1000, 1, "fp32", 80), ("lif", 130, 12, "fp32", 60),
         ("li_tanh", 203, 70, "fp32", 30), ("iku", 128, 128, "auto", 30),
         ("qif", 600, 40, "auto", 40)]         # ragged shape padded onto the tensor-core path (640 x 128): the guarded buffers are the padded ones


@pytest.mark.parametrize("model,n,B,prec,T", CASES)
def test_engine_stays_inside_caller_buffers(model, n, B, prec, T, monkeypatch):
    import rectipy_b200 as rp
    from rectipy_b200 import engine
    engine.clear_plans()
    rng = np.random.default_rng(n + B)
    rate = model.startswith("li_")
    dt = 1e-2 if rate else (1e-1 if model == "iku" else 1e-3)
    m, k = 2, 3
    W = rng.standard_normal((n, n)) * 2.0 / np.sqrt(n)
    if model == "iku":
        W = np.abs(W) / np.sqrt(n)
    w_in, w_out = rng.standard_normal((n, m)) * (10.0 if model == "iku" else 1.0), rng.standard_normal((k, n)) / np.sqrt(n)
    path, op, svar, tvar = TEMPLATE_PATH[model]
    params = {"qif": dict(eta=orc.lorentzian_etas(n)), "qif_sfa": dict(eta=orc.lorentzian_etas(n, eta=0.0), alpha=0.4),
              "lif": dict(eta=10.0, tau=rng.uniform(10, 20, n), tau_s=5.0, k=2.0), "li_tanh": dict(tau=rng.uniform(1, 2, n), k=1.2),
              "iku": dict(eta=rng.uniform(60.0, 160.0, n), g=1.5)}[model]
    skw = dict(spike_threshold=10.0, spike_reset=-10.0) if model == "lif" else (dict(spike_threshold=40.0, spike_reset=-60.0) if model == "iku" else {})
    net = rp.Network(dt, device="cuda:0", batch=B, precision=prec)
    kw = dict(weights=W, source_var=svar, target_var=tvar, input_var=f"{op}/I_ext", node_vars={f"{op}/{p}": v for p, v in params.items()},
              train_params=["weights", f"{op}/eta"])
    if rate:
        kw.update(output_var=f"{op}/v")
    else:
        kw.update(spike_var=f"{op}/spike", reset_var=f"{op}/v", output_var=f"{op}/s", **skw)
    node = net.add_diffeq_node("rnn", path, **kw)
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=w_in, train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=w_out, train="gd")
    if not rate and model != "iku":
        node.reset(np.concatenate([rng.uniform(-50, 9 if model == "lif" else 99, (B, n)), np.zeros((B, (3 if model == "qif_sfa" else 2) * n - n))],
                                  axis=1).astype(np.float32) if B > 1 else
                   np.concatenate([rng.uniform(-50, 99, n), np.zeros((3 if model == "qif_sfa" else 2) * n - n)]).astype(np.float32))
    x = (10.0 * np.sin(2 * np.pi * 3.0 * np.arange(T)[:, None, None] * dt + rng.uniform(0, 6.28, (1, B, m))) + 8.0).astype(np.float32)
    guard = GuardedAlloc()
    monkeypatch.setattr(engine.torch, "empty", guard.empty)
    monkeypatch.setattr(engine.torch, "empty_like", guard.empty_like)
    obs = net.run(x if B > 1 else x[:, 0], sampling_steps=3, cutoff=2, verbose=False, enable_grad=True,
                  record_vars=[("rnn", f"{op}/v", False), ("rnn", f"{op}/v", True)] if prec == "fp32" else [])
    out = torch.stack(obs["out"])
    out.square().sum().backward()
    torch.cuda.synchronize()
    monkeypatch.undo()
    guard.check()
    assert torch.isfinite(out).all() and torch.isfinite(node["weights"].grad).all()
    engine.clear_plans()


@pytest.mark.parametrize("template,B,spiking", [("fhn", 3, False), ("adex", 5, True)])
def test_generated_kernels_stay_inside_caller_buffers(template, B, spiking, tmp_path, monkeypatch):
    """The run-time compiled step / adjoint kernels (rectipy_b200/jit.py) under the same canary guard."""
    import rectipy_b200 as rp
    from rectipy_b200 import engine
    from test_gpu_jit import YAML
    (tmp_path / "mymodels").mkdir()
    (tmp_path / "mymodels" / "custom.yaml").write_text(YAML)
    monkeypatch.chdir(tmp_path)
    engine.clear_plans()
    n, m, k, T, dt = 37, 2, 3, 50, 0.01
    rng = np.random.default_rng(B)
    W = rng.standard_normal((n, n)) / np.sqrt(n)
    net = rp.Network(dt, device="cuda:0", batch=B)
    if spiking:
        node = net.add_diffeq_node("rnn", "mymodels.custom.adex", weights=W, source_var="s", target_var="s_in", input_var="I_ext", output_var="s",
                                   spike_var="spike", reset_var="v", train_params=["weights", "adex_op/b"], spike_threshold=2.0, spike_reset=-1.5)
    else:
        node = net.add_diffeq_node("rnn", "mymodels.custom.fhn", weights=W, source_var="r", target_var="r_in", input_var="I_ext", output_var="v",
                                   train_params=["weights", "fhn_op/a"])
    net.add_func_node("inp", m, "identity"); net.add_edge("inp", "rnn", weights=rng.standard_normal((n, m)), train="gd")
    net.add_func_node("out", k, "identity"); net.add_edge("rnn", "out", weights=rng.standard_normal((k, n)) / np.sqrt(n), train="gd")
    x = (np.sin(np.arange(T)[:, None, None] * 0.3 + rng.uniform(0, 6.28, (1, B, m))) + 1.2).astype(np.float32)
    guard = GuardedAlloc()
    monkeypatch.setattr(engine.torch, "empty", guard.empty)
    monkeypatch.setattr(engine.torch, "empty_like", guard.empty_like)
    obs = net.run(x, sampling_steps=3, cutoff=2, verbose=False, enable_grad=True, record_vars=[("rnn", "w", True)])
    out = torch.stack(obs["out"])
    out.square().sum().backward()
    torch.cuda.synchronize()
    monkeypatch.undo()
    guard.check()
    assert torch.isfinite(out).all() and torch.isfinite(node["weights"].grad).all()
    engine.clear_plans()
