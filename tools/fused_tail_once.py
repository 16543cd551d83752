"""Two launches of tcd_gemm_film_residual_norm (cross-attention tail shape of the c2 sampler: 96 000 rows, K = 512,
x in place, bf16 operand out) for an `ncu --set full -k regex:gemm_frn --launch-skip 1 -c 1` capture."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from tcdiff_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
R, K, D, L = 96000, 512, 512, 750
g = torch.Generator().manual_seed(0)
a = (torch.randn(R, K, generator=g) * 0.5).to(dev, torch.bfloat16)
w = (torch.randn(D, K, generator=g) / K ** 0.5).to(dev, torch.bfloat16)
x = torch.randn(R, D, generator=g).to(dev)
gi, bi, gn, bn = (torch.randn(D, generator=g).to(dev) for _ in range(4))
film = (0.3 * torch.randn(R // L, 2 * D, generator=g)).to(dev)
plain = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
for _ in range(2):
    ops.gemm_film_residual_norm(a, w, None, x, x, (gi, bi), 1e-6, film, film.stride(0), 0, (gn, bn), 1e-5, plain, None, None,
                                None, R, L)
torch.cuda.synchronize()
print("ok", float(plain.float().abs().mean()))
