"""Scratch experiment (GPU): 3xTF32 accumulation error vs K-chunking, and raw contraction timings."""
import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from rectipy_b200 import engine, _cabi as abi

def rel(a, b): return float((a.double() - b).abs().max() / b.abs().max())

torch.manual_seed(0)
P, Q, K = 512, 1024, 2048
A = torch.randn(P, K, device="cuda"); B = torch.randn(Q, K, device="cuda")
ref = B.double() @ A.double().T
print("one shot  :", rel(engine.gemm_tn(A, B, precision=abi.RP_PREC_3XTF32), ref))
for chunk in (1024, 512, 256, 128, 64):
    C = torch.zeros(Q, P, device="cuda")
    for k0 in range(0, K, chunk):
        engine.gemm_tn(A[:, k0:k0 + chunk], B[:, k0:k0 + chunk], precision=abi.RP_PREC_3XTF32, out=C, accumulate=True)
    print(f"chunk {chunk:5d}:", rel(C, ref))
print("fp32 simt :", rel(engine.gemm_tn(A, B), ref), " torch:", rel(B @ A.T, ref))
# positive operands (all products same sign) expose a truncating accumulator most clearly
Ap, Bp = A.abs(), B.abs(); refp = Bp.double() @ Ap.double().T
Cp = engine.gemm_tn(Ap, Bp, precision=abi.RP_PREC_3XTF32)
print("positive operands: 3xtf32 signed mean rel err", float(((Cp.double() - refp) / refp).mean()), " fp32 simt", float(((engine.gemm_tn(Ap, Bp).double() - refp) / refp).mean()))

def timeit(fn, n=10):
    for _ in range(3): fn()
    torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize(); return e0.elapsed_time(e1) / n

# engine-level timings at the headline shape
import numpy as np, rectipy_b200 as rp
for prec in ("3xtf32", "fp32"):
    n, Bt, T = 4096, 1024, 20
    net = rp.Network(1e-3, device="cuda:0", batch=Bt, precision=prec)
    rng = np.random.default_rng(0)
    node = net.add_diffeq_node("qif", "neuron_model_templates.spiking_neurons.qif.qif", weights=rng.standard_normal((n, n)).astype(np.float32) * 2 / np.sqrt(n),
                               source_var="s", target_var="s_in", input_var="I_ext", output_var="s", spike_var="spike", reset_var="v", op="qif_op", train_params=["weights"])
    net.add_func_node("inp", 2, "identity"); net.add_edge("inp", "qif", weights=rng.standard_normal((n, 2)))
    net.add_func_node("out", 3, "identity"); net.add_edge("qif", "out", weights=rng.standard_normal((3, n)) / 64, train="gd")
    x = torch.randn(T, Bt, 2, device="cuda") + 8.0
    tgt = torch.randn(T, Bt, 3, device="cuda")
    def fwd():
        net.reset(); net.run(x, verbose=False, enable_grad=False)
    def bptt():
        net.reset(); obs = net.run(x, verbose=False, enable_grad=True)
        torch.nn.functional.mse_loss(torch.stack(obs["out"]), tgt).backward()
    tf = timeit(fwd, 3); tb = timeit(bptt, 3)
    print(f"{prec}: fwd {tf / T * 1e3:.1f} us/step -> {n * Bt * T / tf * 1e3:.3e} neuron-steps/s ; bptt {tb / T * 1e3:.1f} us/step -> {n * Bt * T / tb * 1e3:.3e} neuron-steps/s")
    flop = 2.0 * n * n * Bt
    print(f"   fwd GEMM-equivalent {flop * T / tf * 1e3 / 1e12:.1f} TFLOP/s (x3 MMA passes for 3xtf32)")
