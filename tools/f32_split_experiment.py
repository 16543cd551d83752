"""Experiment: fp32-mode GEMMs as K-concatenated bf16 pieces on the tcgen05 kernel vs the SIMT fp32 kernel."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
from tcdiff_b200 import ops
orig = ops.gemm
MODE = [0]
def up8(v): return (v + 7) // 8 * 8
def split(x, Kp):
    x0 = x.bfloat16(); r = x - x0.float(); x1 = r.bfloat16(); r2 = r - x1.float(); x2 = r2.bfloat16()
    pad = lambda t: torch.nn.functional.pad(t, (0, Kp - t.shape[1]))
    return pad(x0), pad(x1), pad(x2)
def gemm(a, w, bias, act, out, M=None, N=None, K=None, lda=None, ldw=None, ldc=None):
    if a.dtype != torch.float32 or MODE[0] == 0:
        return orig(a, w, bias, act, out, M=M, N=N, K=K, lda=lda, ldw=ldw, ldc=ldc)
    M = a.shape[0] if M is None else M; K = a.shape[1] if K is None else K; N = w.shape[0] if N is None else N
    lda = a.stride(0) if lda is None else lda; ldw = w.stride(0) if ldw is None else ldw
    A = torch.as_strided(a, (M, K), (lda, 1)); W = torch.as_strided(w, (N, K), (ldw, 1))
    Kp = up8(K)
    a0, a1, a2 = split(A, Kp); w0, w1, w2 = split(W, Kp)
    if MODE[0] == 3:
        A2 = torch.cat([a0, a0, a1], 1); W2 = torch.cat([w0, w1, w0], 1)
    else:
        A2 = torch.cat([a0, a0, a1, a1, a0, a2], 1); W2 = torch.cat([w0, w1, w0, w1, w2, w0], 1)
    return orig(A2.contiguous(), W2.contiguous(), bias, act, out, M=M, N=N, K=A2.shape[1], ldc=ldc)
ops.gemm = gemm
dev = torch.device("cuda:0")
# 1) single GEMM error vs float64
g = torch.Generator().manual_seed(0)
for (M, N, K) in [(4096, 512, 512), (4096, 1024, 512), (4096, 512, 1024), (2048, 512, 2190)]:
    a = torch.randn(M, K, generator=g).to(dev); w = (torch.randn(N, K, generator=g) / K ** 0.5).to(dev)
    ref = (a.double() @ w.double().T)
    for mode in (0, 3, 6):
        MODE[0] = mode
        out = torch.empty(M, N, device=dev)
        ops.gemm(a, w, None, 0, out)
        e = (out.double() - ref)
        print(f"gemm {M}x{N}x{K} mode {mode}: rel-L2 {float(e.norm() / ref.norm()):.3e} max {float(e.abs().max()):.3e} mean-bias {float(e.mean()):.3e}")
# 2) model forward fp32 vs golden
from test_gpu_model import build, load_golden, rel, synth
for name in ("tiny", "c1"):
    gld = load_golden(f"{name}_forward.pt")
    cfg, sd, m, _ = build(name, "fp32", dev)
    B, L = gld["B"], 150 * cfg["dancers"]
    x = torch.randn(B, L, 151, generator=torch.Generator().manual_seed(gld["x_seed"]))
    cond = synth.make_music(B, cfg["cond_feature_dim"]); t = torch.tensor(gld["times"]); st = gld["row_stride"]
    for mode in (0, 3, 6):
        MODE[0] = mode
        torch.cuda.synchronize(); t0 = time.perf_counter()
        gd = m.guided_forward(x.to(dev), cond.to(dev), t.to(dev), 2.0).cpu()
        dt_ = time.perf_counter() - t0
        print(f"forward {name} mode {mode}: rel vs reference golden {rel(gd[:, ::st], gld['guided']):.3e}  ({dt_ * 1e3:.1f} ms)")
