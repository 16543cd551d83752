"""tcd_gemm_film_residual_norm (csrc/gemm_frn.cu): the `fc` / `linear2` GEMM fused with the FiLM + residual + LayerNorm
tail, checked against a torch fp32 restatement of model/model.py:103-106,171-173,327,334,339.  The entry ships in the
product library whether or not the engine's tails use it (csrc/tuning.cuh TCD_TUNE_FUSE_TAILS), so its parity is always
tested — including, at the full c2 row count (eleven tiles per CTA pair), repeated launches with a residual whose every
element encodes its own (row, column): the kernel's residual ring, tensor-memory accumulators and statistics exchange are
reused across tiles, and each of the races found while it was written (r02) showed up as a few 16-byte pieces of x taken from
another box, roughly once per ten launches."""
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
    c["cos_t"], c["sin_t"] = c["cos"].t().contiguous(), c["sin"].t().contiguous()      # the fused entry reads them angle-major
    return c


def _reference(c, R, foff, dev):
    y = c["a"].float() @ c["w"].float().t()
    if c["bias"] is not None:
        y = y + c["bias"]
    if c["ln_in"] is not None:
        y = torch.nn.functional.layer_norm(y, (D,), *c["ln_in"], 1e-6)
    samp = torch.arange(R, device=dev) // L
    if c["film"] is not None:
        y = (1 + c["film"][samp, foff:foff + D]) * y + c["film"][samp, foff + D:foff + 2 * D]
    v = y if c["x"] is None else c["x"] + y
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
                                c["film"].stride(0), foff, c["ln_next"], 1e-5, plain, rot, c["cos_t"] if want_rot else None,
                                c["sin_t"] if want_rot else None, R, L)
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


@pytest.mark.parametrize("R,K,bias,inner,foff,want_x", [
    (96000, 512, False, True, 2 * D, True),      # cross-attention tail, c2 size
    (1130, 1024, True, False, 4 * D, False),     # feed-forward tail, ragged rows, dead residual
])
def test_fused_tail_without_next_affine(dev, R, K, bias, inner, foff, want_x):
    """ln_next = None: the tail emits the normalised rows without gamma / beta (the engine folds norm3 / norm4's affine into
    linear1 / linear3's weights, csrc/tuning.cuh TCD_TUNE_FOLD_LN); asking for the rotated operand that way is an error."""
    from tcdiff_b200 import ops, _lib
    c = _case(dev, R, K, bias, inner)
    c["ln_next"] = (torch.ones(D, device=dev), torch.zeros(D, device=dev))
    v_ref, n_ref, _ = _reference(c, R, foff, dev)
    x = c["x"].clone()
    plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16)
    ops.gemm_film_residual_norm(c["a"], c["w"], c["bias"], x, x if want_x else None, c["ln_in"], 1e-6, c["film"],
                                c["film"].stride(0), foff, None, 1e-5, plain, None, None, None, R, L)
    torch.cuda.synchronize()
    if want_x:
        assert float((x - v_ref).abs().max()) < 2e-3
    assert float((plain.float() - n_ref).abs().max()) < 4e-2
    with pytest.raises(_lib.TcdError):
        ops.gemm_film_residual_norm(c["a"], c["w"], c["bias"], x, None, c["ln_in"], 1e-6, c["film"], c["film"].stride(0), foff,
                                    None, 1e-5, None, plain, c["cos_t"], c["sin_t"], R, L)


@pytest.mark.parametrize("name,K,bias,inner,want_x,want_plain,want_rot", [
    ("self-attention", 512, False, True, True, False, True),
    ("cross-attention", 512, False, True, True, True, False),
    ("feed-forward", 1024, True, False, False, True, False),
])
def test_fused_tail_many_tiles_repeated(dev, name, K, bias, inner, want_x, want_plain, want_rot):
    """96 000 rows (11 tiles per CTA pair), 8 launches, x[r, c] = (r mod 1024) + c / 1024: a piece of x read from the
    wrong box, a stale accumulator or a torn statistics exchange shows up as an error of order 1."""
    from tcdiff_b200 import ops
    R = 96000
    c = _case(dev, R, K, bias, inner)
    rr = torch.arange(R, device=dev, dtype=torch.float32)[:, None]
    cc = torch.arange(D, device=dev, dtype=torch.float32)[None, :]
    c["x"] = ((rr % 1024) + cc / 1024.0).contiguous()
    v_ref, n_ref, r_ref = _reference(c, R, 0, dev)
    for trial in range(8):
        x = c["x"].clone()
        plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_plain else None
        rot = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_rot else None
        ops.gemm_film_residual_norm(c["a"], c["w"], c["bias"], x, x if want_x else None, c["ln_in"], 1e-6, c["film"],
                                    c["film"].stride(0), 0, c["ln_next"], 1e-5, plain, rot, c["cos_t"] if want_rot else None,
                                    c["sin_t"] if want_rot else None, R, L)
        torch.cuda.synchronize()
        if want_x:                                  # |x| <= 1024: fp32 ulp 1.2e-4
            assert float((x - v_ref).abs().max()) < 5e-3, (name, trial)
        if want_plain:
            assert float((plain.float() - n_ref).abs().max()) < 4e-2, (name, trial)
        if want_rot:
            assert float((rot.float() - r_ref).abs().max()) < 4e-2, (name, trial)


@pytest.mark.parametrize("R,xin,film,bias,inner", [
    (1130, True, False, False, True),       # residual without modulation
    (3000, False, False, True, False),      # linear3-style: v = y + bias, LayerNorm + rotary of it, v written out
    (1000, False, True, True, True),
])
def test_fused_tail_optional_parts(dev, R, xin, film, bias, inner):
    from tcdiff_b200 import ops
    c = _case(dev, R, 512, bias, inner)
    if not xin:
        c["x"] = None
    if not film:
        c["film"] = None
    v_ref, n_ref, r_ref = _reference(c, R, 0, dev)
    xo = torch.zeros(R, D, device=dev)
    plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16)
    rot = torch.zeros(R, D, device=dev, dtype=torch.bfloat16)
    ops.gemm_film_residual_norm(c["a"], c["w"], c["bias"], c["x"], xo, c["ln_in"], 1e-6, c["film"],
                                c["film"].stride(0) if film else 0, 0, c["ln_next"], 1e-5, plain, rot, c["cos_t"], c["sin_t"], R, L)
    torch.cuda.synchronize()
    assert float((xo - v_ref).abs().max()) < 2e-3
    assert float((plain.float() - n_ref).abs().max()) < 4e-2
    assert float((rot.float() - r_ref).abs().max()) < 4e-2
