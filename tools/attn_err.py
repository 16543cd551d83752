"""Forward attention error of the loaded library against an fp32 torch reference on the same bf16 inputs (A/B aid)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tcdiff_b200 import ops, _lib
if "--lib" in sys.argv:
    _lib.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
dev = torch.device("cuda:0")
print("library", _lib.LIB_PATH, "attn_2q", _lib.lib().tcd_tuning(b"attn_2q"))
H, hd = 8, 64
for (n, Lq, Lk, amp) in [(3, 450, 450, 1.5), (3, 450, 152, 1.5), (2, 750, 750, 1.5), (2, 750, 750, 0.2), (2, 450, 450, 4.0)]:
    g = torch.Generator(device=dev).manual_seed(n * 1000 + Lq + Lk)
    q = (torch.randn(n, Lq, H * hd, device=dev, generator=g) * amp).to(torch.bfloat16)
    k = (torch.randn(n, Lk, H * hd, device=dev, generator=g) * amp).to(torch.bfloat16)
    v = torch.randn(n, Lk, H * hd, device=dev, generator=g).to(torch.bfloat16)
    o, lse = ops.attention_train_forward(q, k, v, H, 0.125)
    s = torch.einsum("nqhd,nkhd->nhqk", q.float().view(n, Lq, H, hd), k.float().view(n, Lk, H, hd)) * 0.125
    ref = torch.einsum("nhqk,nkhd->nqhd", s.softmax(-1), v.float().view(n, Lk, H, hd)).reshape(n, Lq, H * hd)
    lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
    e = (o.float() - ref)
    print(f"n{n} Lq{Lq} Lk{Lk} amp{amp}: O rel-L2 {float(e.norm() / ref.norm()):.3e} max {float(e.abs().max()):.3e} mean-signed {float(e.mean()):.2e}  "
          f"lse max|d| {float((lse - lse_ref).abs().max()):.3e} lse mean-signed {float((lse - lse_ref).mean()):.2e}")
