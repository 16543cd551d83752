"""tcd_gemm_film_residual_norm (csrc/gemm_frn.cu): the `fc` / `linear2` GEMM fused with the FiLM + residual + LayerNorm
tail, checked against a torch fp32 restatement of model/model.py:103-106,171-173,327,334,339 and against the
unfused tcd_gemm + tcd_film_residual_norm pair.  The entry ships in the product library whether or not the engine's tails use
it (csrc/tuning.cuh TCD_TUNE_FUSE_TAILS), so its parity is always tested."""
import pytest
import torch

pytestmark = pytest.mark.gpu

D, L = 512, 750


def _case(dev, R, K, bias, inner, seed=3):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    n = (R + L - 1) // L
    c = dict(a=(rn(R, K) * 0.5).to(dev, torch.bfloat16), w=(rn(D, K) / K ** 0.5).to(dev, torch.bfloat16),
             bias=(rn(D) * 0.1).to(dev) if bias else None, x=rn(R, D).to(dev),
             ln_in=((1 + 0.1 * rn(D)).to(dev), (0.1 * rn(D)).to(dev)) if inner else None,
             film=(0.3 * rn(n, 6 * D)).to(dev), ln_next=((1 + 0.1 * rn(D)).to(dev), (0.1 * rn(D)).to(dev)))
    ang = torch.arange(L, dtype=torch.float32)[:, None] * (10000.0 ** (-torch.arange(0, D, 2).float() / D))[None, :]
    c["cos"], c["sin"] = ang.cos().to(dev).contiguous(), ang.sin().to(dev).contiguous()
    return c


def _reference(c, R, foff, dev):
    y = c["a"].float() @ c["w"].float().t()
    if c["bias"] is not None:
        y = y + c["bias"]
    if c["ln_in"] is not None:
        y = torch.nn.functional.layer_norm(y, (D,), *c["ln_in"], 1e-6)
    samp = torch.arange(R, device=dev) // L
    v = c["x"] + (1 + c["film"][samp, foff:foff + D]) * y + c["film"][samp, foff + D:foff + 2 * D]
    nrm = torch.nn.functional.layer_norm(v, (D,), *c["ln_next"], 1e-5)
    pos = torch.arange(R, device=dev) % L
    cs, sn = c["cos"][pos], c["sin"][pos]
    rot = torch.empty_like(nrm)
    rot[:, 0::2] = nrm[:, 0::2] * cs - nrm[:, 1::2] * sn
    rot[:, 1::2] = nrm[:, 1::2] * cs + nrm[:, 0::2] * sn
    return v, nrm, rot


@pytest.mark.parametrize("R,K,bias,inner,foff,want_x,want_plain,want_rot", [
    (96000, 512, False, True, 0, True, False, True),        # self-attention tail at the c2 size
    (3000, 512, False, True, 2 * D, True, True, False),     # cross-attention tail
    (1130, 1024, True, False, 4 * D, False, True, False),   # feed-forward tail, ragged rows, dead residual
    (750, 512, False, True, 0, True, True, True),
])
def test_gemm_film_residual_norm(dev, R, K, bias, inner, foff, want_x, want_plain, want_rot):
    from tcdiff_b200 import ops
    c = _case(dev, R, K, bias, inner)
    v_ref, n_ref, r_ref = _reference(c, R, foff, dev)
    x = c["x"].clone()
    plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_plain else None
    rot = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_rot else None
    ops.gemm_film_residual_norm(c["a"], c["w"], c["bias"], x, x if want_x else None, c["ln_in"], 1e-6, c["film"],
                                c["film"].stride(0), foff, c["ln_next"], 1e-5, plain, rot, c["cos"] if want_rot else None,
                                c["sin"] if want_rot else None, R, L)
    torch.cuda.synchronize()
    # tolerances: x is fp32 arithmetic on an fp32 accumulator of bf16 products (1e-3 covers summation order);
    # the bf16 operands carry half an ulp of bf16 (2^-9 relative) on values of magnitude <= ~6
    if want_x:
        assert float((x - v_ref).abs().max()) < 2e-3
    else:
        assert torch.equal(x, c["x"])
    if want_plain:
        assert float((plain.float() - n_ref).abs().max()) < 4e-2
    if want_rot:
        assert float((rot.float() - r_ref).abs().max()) < 4e-2
