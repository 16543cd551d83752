import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
sys.path.insert(0, "tests")
from test_gpu_fused_tail import _case, _reference, D, L
from tcdiff_b200 import ops
dev = torch.device("cuda:0")
R = 96000
c = _case(dev, R, 512, False, True)
rr = torch.arange(R, device=dev, dtype=torch.float32)[:, None]
cc = torch.arange(D, device=dev, dtype=torch.float32)[None, :]
c["x"] = ((rr % 1024) + cc / 1024.0).contiguous()
v_ref, n_ref, r_ref = _reference(c, R, 0, dev)
for t in range(10):
    x = c["x"].clone()
    plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16)
    ops.gemm_film_residual_norm(c["a"], c["w"], None, x, x, c["ln_in"], 1e-6, c["film"], c["film"].stride(0), 0, c["ln_next"], 1e-5,
                                plain, None, None, None, R, L)
    torch.cuda.synchronize()
    bad = (x - v_ref).abs() > 2e-2
    if bad.any():
        bi = bad.nonzero()
        print("trial", t, "bad", int(bad.sum()))
        for a, b in bi[:8].tolist():
            d = float(v_ref[a, b] - c["x"][a, b])
            xu = float(x[a, b]) - d
            print(f"   row {a} (tile {a//128}, wave {a//128//74}, r {a%128}) col {b}: x_true {float(c['x'][a,b]):.4f} x_used {xu:.4f} -> src row%1024 {int(xu)} col {round((xu-int(xu))*1024)}  (delta rows {int(xu) - a % 1024})")
    else:
        print("trial", t, "ok")
