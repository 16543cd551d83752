"""-m gpu: every C-ABI kernel against the CPU oracle / a plain torch fp32 evaluation of the same op."""
import math

import pytest
import torch
import torch.nn.functional as F

from conftest import load_golden
from oracle import synth, tcdiff_oracle as O, p3d

pytestmark = pytest.mark.gpu


def _ops():
    from tcdiff_b200 import ops
    return ops


def rel(a, b):
    return float((a - b).abs().max() / b.abs().max().clamp_min(1e-12))


# ------------------------------------------------------------------------------------------ step kernels
@pytest.mark.parametrize("n_tokens,last,traj", [(450, False, True), (37, False, False), (301, True, True), (0, False, True)])
def test_cfg_ddim_step_bit_exact(dev, n_tokens, last, traj):
    """Integer-free but fp32: the kernel uses round-to-nearest intrinsics in the reference's op order, so
    with identical inputs the result equals the CPU torch evaluation bit for bit."""
    ops = _ops()
    g = torch.Generator().manual_seed(3)
    x, con, unc, nz = (torch.randn(n_tokens, 151, generator=g) * s for s in (5.0, 1.2, 1.2, 1.0))
    tr = torch.randn(n_tokens, 3, generator=g)
    sched = O.make_schedule("cosine", 1000)
    t, tn = 979, 959
    sa, c, sigma = O.ddim_coeffs(sched, t, tn)
    sr, srm1 = sched["sqrt_recip_alphas_cumprod"][t], sched["sqrt_recipm1_alphas_cumprod"][t]
    w = 2.0
    o = unc + (con - unc) * w
    x0 = o.clamp(-1.0, 1.0)
    eps = (sr * x - x0) / srm1
    ref = x0.clone() if last else x0 * sa + c * eps + sigma * nz
    if traj:
        ref[:, 4], ref[:, 5] = tr[:, 0], tr[:, 1]
    xd, cd, ud, nd, td = (v.to(dev) for v in (x, con, unc, nz, tr))
    out = torch.empty_like(xd)
    x0o = torch.empty_like(xd)
    xpad = torch.zeros(n_tokens, 160, dtype=torch.bfloat16, device=dev)
    ops.cfg_ddim_step(xd, cd, ud, None if last else nd, td if traj else None, out, x0o, xpad, 160, n_tokens, w,
                      float(sr), float(srm1), float(sa), float(c), float(sigma), True, last)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(x0o.cpu(), x0)
    assert torch.equal(xpad[:, :151].float().cpu(), ref.bfloat16().float())
    assert float(xpad[:, 151:].abs().sum()) == 0.0
    # in place
    ops.cfg_ddim_step(xd, cd, ud, None if last else nd, td if traj else None, xd, None, None, 0, n_tokens, w,
                      float(sr), float(srm1), float(sa), float(c), float(sigma), True, last)
    assert torch.equal(xd.cpu(), ref)


def test_cfg_ddpm_step_and_constraint(dev):
    ops = _ops()
    g = torch.Generator().manual_seed(4)
    n = 333
    x, con, unc, nz, val = (torch.randn(n, 151, generator=g) for _ in range(5))
    mask = (torch.rand(n, 151, generator=g) > 0.5).float()
    sched = O.make_schedule("cosine", 1000)
    for i, use_mask in ((500, False), (0, False), (7, True)):
        c1, c2 = sched["posterior_mean_coef1"][i], sched["posterior_mean_coef2"][i]
        std = (0.5 * sched["posterior_log_variance_clipped"][i]).exp()
        w = 1.0 if i < 100 else 2.0
        x0 = (unc + (con - unc) * w).clamp(-1, 1)
        nzm = 0.0 if i == 0 else 1.0
        ref = c1 * x0 + c2 * x + nzm * std * nz
        if use_mask:
            ref = val * mask + (1.0 - mask) * ref
        out = torch.empty(n, 151, device=dev)
        ops.cfg_ddpm_step(x.to(dev), con.to(dev), unc.to(dev), nz.to(dev), out, None, 0, n, w, float(c1), float(c2),
                          float(std), i != 0, mask.to(dev) if use_mask else None, val.to(dev) if use_mask else None)
        assert torch.equal(out.cpu(), ref), i
        # predict_epsilon=True: the blended output is the noise, x_recon = sr * x - srm1 * out (model/diffusion.py:181-185)
        sr, srm1 = sched["sqrt_recip_alphas_cumprod"][i], sched["sqrt_recipm1_alphas_cumprod"][i]
        x0e = (sr * x - srm1 * (unc + (con - unc) * w)).clamp(-1, 1)
        refe = c1 * x0e + c2 * x + nzm * std * nz
        if use_mask:
            refe = val * mask + (1.0 - mask) * refe
        ops.cfg_ddpm_step(x.to(dev), con.to(dev), unc.to(dev), nz.to(dev), out, None, 0, n, w, float(c1), float(c2),
                          float(std), i != 0, mask.to(dev) if use_mask else None, val.to(dev) if use_mask else None,
                          eps_coef=(float(sr), float(srm1)))
        assert torch.equal(out.cpu(), refe), i


def test_q_sample_matches_p_losses_front(dev):
    ops = _ops()
    B, dn, S = 3, 2, 150
    xs = synth.make_motion(B, dn, S, seed=7)
    noise = torch.randn(B, S, dn, 151, generator=torch.Generator().manual_seed(8))
    t = torch.tensor([0, 500, 999])
    sched = O.make_schedule("cosine", 1000)
    xp = xs.permute(0, 2, 1, 3)
    ref = O.q_sample(sched, xp, t, noise)
    ref[:, :, :, [4, 5]] = xp[:, :, :, [4, 5]]
    out = torch.empty(B, S, dn, 151, device=dev)
    tgt = torch.empty(B, S, dn, 151, device=dev)
    ops.q_sample(xs.to(dev), noise.to(dev), t.to(dev), sched["sqrt_alphas_cumprod"].to(dev),
                 sched["sqrt_one_minus_alphas_cumprod"].to(dev), out, tgt, None, 0, B, dn, S, True, True)
    assert torch.equal(out.cpu(), ref)
    assert torch.equal(tgt.cpu(), xp.contiguous())


def test_inpaint_traj_exact(dev):
    ops = _ops()
    n = 300
    x = torch.randn(n, 151, device=dev)
    ref = x.clone()
    tr = torch.randn(n, 3, device=dev)
    ref[:, 4], ref[:, 5] = tr[:, 0], tr[:, 1]
    ops.inpaint_traj(x, tr, None, 0, n)
    assert torch.equal(x, ref)
    xpad = torch.zeros(n, 160, dtype=torch.bfloat16, device=dev)
    x2 = torch.randn(n, 151, device=dev)
    r2 = x2.clone()
    r2[:, 4], r2[:, 5] = tr[:, 0], tr[:, 1]
    ops.inpaint_traj(x2, tr, xpad, 160, n)
    assert torch.equal(x2, r2) and torch.equal(xpad[:, :151].float(), r2.bfloat16().float())


# ------------------------------------------------------------------------------------------ kinematics / loss
def test_ax_from_6v_and_fk_golden(dev):
    from tcdiff_b200 import SMPLSkeleton, ax_from_6v
    g = load_golden("fk.pt")
    aa = ax_from_6v(g["d6"].to(dev)).cpu()
    # tolerance: 1e-4 relative (north_star); the w==0 branch rows compare on rotation, not on the axis sign
    R_ref = p3d.axis_angle_to_matrix(g["axis_angle"])
    R_got = p3d.axis_angle_to_matrix(aa)
    assert float((R_ref - R_got).abs().max()) < 2e-5
    generic = torch.ones(g["d6"].shape[:-1], dtype=torch.bool)
    generic[0, 0, :3] = False
    assert float((aa - g["axis_angle"])[generic].abs().max()) < 1e-4
    pos = SMPLSkeleton(dev).forward(g["axis_angle"].to(dev), g["root"].to(dev)).cpu()
    assert rel(pos, g["positions"]) < 1e-5
    zp = SMPLSkeleton(dev).forward(torch.zeros(1, 1, 24, 3, device=dev), torch.zeros(1, 1, 3, device=dev)).cpu()
    assert float((zp - g["zero_pose"]).abs().max()) < 1e-7
    off = torch.tensor(O.SMPL_OFFSETS)
    assert float((zp[0, 0, 10] - (off[1] + off[4] + off[7] + off[10])).abs().max()) < 1e-6   # SURVEY §4 KAT


def test_motion_fk_matches_oracle_chain(dev):
    """Direct matrix-chain FK == 6D -> axis-angle -> quaternion FK of the reference, 1e-4 relative."""
    from tcdiff_b200 import SMPLSkeleton
    m = synth.make_prediction(4, 3, seed=11).reshape(-1, 151)
    ref = O.smpl_forward(O.ax_from_6v(m[:, 7:].reshape(1, -1, 24, 6)), m[:, 4:7].reshape(1, -1, 3))[0]
    got = SMPLSkeleton(dev).motion_forward(m.to(dev)).cpu()
    assert rel(got, ref) < 1e-4
    assert float((got - ref).abs().max()) < 2e-5


@pytest.mark.parametrize("B,dn", [(3, 3), (2, 5), (1, 1)])
def test_loss_forward_vs_oracle(dev, B, dn):
    ops = _ops()
    S = 150
    target = synth.make_motion(B, dn, S, seed=50).permute(0, 2, 1, 3).contiguous()
    pred = synth.make_prediction(B, dn, S, seed=51).reshape(B, S, dn, 151)
    p2w = torch.rand(B, generator=torch.Generator().manual_seed(1)) + 0.5
    tot, parts = O.loss_terms(pred, target, p2w)
    ref = torch.stack([tot] + list(parts))
    got = ops.loss_forward(pred.to(dev), target.to(dev), p2w.to(dev), B, S, dn).cpu()
    assert float(ref[4]) > 0
    assert float(((got - ref).abs() / ref.abs()).max()) < 1e-4, (got, ref)


def test_loss_forward_golden(dev):
    ops = _ops()
    g = load_golden("loss_terms.pt")
    B, dn = g["B"], g["dn"]
    target = synth.make_motion(B, dn, seed=g["target_seed"]).permute(0, 2, 1, 3).contiguous()
    pred = synth.make_prediction(B, dn, seed=g["pred_seed"]).reshape(B, 150, dn, 151)
    got = ops.loss_forward(pred.to(dev), target.to(dev), None, B, 150, dn).cpu()
    assert float(((got - g["losses"]).abs() / g["losses"].abs()).max()) < 1e-4, (got, g["losses"])


def test_loss_l1_forward_golden_and_backward(dev):
    """loss_type "l1" (F.l1_loss, the reference constructor's default, model/diffusion.py:172): the four terms against the
    reference's own numbers for a given prediction, and d total / d model_out against autograd through the oracle."""
    ops = _ops()
    g = load_golden("tiny_defaults.pt")
    B, dn, S = g["B"], 3, 150
    pred = synth.make_prediction(B, dn, seed=51).reshape(B, S, dn, 151)
    nz = torch.randn(B, S, dn, 151, generator=torch.Generator().manual_seed(g["terms_noise_seed"])).clamp(-1, 1)
    got = ops.loss_forward(pred.to(dev), nz.to(dev), None, B, S, dn, "l1").cpu()
    # the target of this variant is clamped Gaussian noise read as 6-D rotations: near-degenerate Gram-Schmidt inputs, where
    # the kernel's direct matrix chain and the reference's matrix -> quaternion -> axis-angle -> quaternion route round
    # differently (measured 1.7e-4 on the FK term, 0 on the others)
    assert float(((got - g["terms_losses"]).abs() / g["terms_losses"].abs()).max()) < 5e-4, (got, g["terms_losses"])
    target = synth.make_motion(B, dn, S, seed=70).permute(0, 2, 1, 3).contiguous()
    p = synth.make_prediction(B, dn, S, seed=71).reshape(B, S, dn, 151).clone().requires_grad_(True)
    p2w = torch.rand(B, generator=torch.Generator().manual_seed(2)) + 0.5
    tot, _ = O.loss_terms(p, target, p2w, "l1")
    (tot * 1.7).backward()
    gb = ops.loss_backward(p.detach().to(dev), target.to(dev), p2w.to(dev), 1.7, B, S, dn, "l1").cpu()
    assert torch.isfinite(gb).all()
    # sign() flips where a difference is within rounding of zero: compare in the mean, and require the bulk to agree
    err = (gb - p.grad).abs()
    assert float(err.mean() / p.grad.abs().mean()) < 2e-3
    assert float((err > 1e-3 * p.grad.abs().max()).float().mean()) < 2e-3


# ------------------------------------------------------------------------------------------ GEMM
ACTS = {0: lambda v: v, 1: F.relu, 2: F.gelu, 3: F.mish, 4: F.silu}


@pytest.mark.parametrize("M,N,K,act,bias", [(900, 512, 151, 0, True), (300, 1024, 1024, 1, True), (65, 151, 512, 0, True),
                                            (257, 70, 35, 2, False), (1, 512, 512, 3, True), (128, 64, 16, 4, True)])
def test_gemm_f32(dev, M, N, K, act, bias):
    ops = _ops()
    g = torch.Generator().manual_seed(M + N + K)
    a, w = torch.randn(M, K, generator=g), torch.randn(N, K, generator=g) / math.sqrt(K)
    b = torch.randn(N, generator=g) if bias else None
    ref = ACTS[act](F.linear(a, w, b))
    out = torch.empty(M, N, device=dev)
    ops.gemm(a.to(dev), w.to(dev), b.to(dev) if bias else None, act, out)
    assert rel(out.cpu(), ref) < 2e-5


@pytest.mark.parametrize("M,N,K,act,bias,outbf", [
    (128, 256, 64, 0, False, False),       # exactly one tile, one k-block
    (128, 256, 512, 0, True, False),       # one tile, 8 k-blocks (pipeline wraps twice)
    (300, 512, 160, 0, True, True),        # M tail, K tail (160 = 2.5 k-blocks), bf16 out
    (1000, 1024, 512, 2, True, True),      # multi-tile persistent loop, GELU epilogue
    (257, 151, 512, 0, True, False),       # N tail + unaligned ldc (scalar stores)
    (2, 2048, 512, 3, True, True),         # tiny M (time MLP), Mish
    (5000, 2560, 1024, 1, True, True),     # fusion-projection shape, > 148 tiles
    (9, 24576, 512, 0, True, False),       # FiLM table shape
])
def test_gemm_bf16_tcgen05(dev, M, N, K, act, bias, outbf):
    """tcgen05/TMA GEMM vs fp32 matmul of the same bf16-rounded operands (fp32 accumulate => tolerance
    is accumulation-order noise, plus one bf16 rounding when the output is bf16)."""
    ops = _ops()
    g = torch.Generator().manual_seed(M * 7 + N * 3 + K)
    a = torch.randn(M, K, generator=g).bfloat16()
    w = (torch.randn(N, K, generator=g) / math.sqrt(K)).bfloat16()
    b = torch.randn(N, generator=g) if bias else None
    ref = ACTS[act](F.linear(a.float(), w.float(), b))
    out = torch.full((M, N), float("nan"), device=dev, dtype=torch.bfloat16 if outbf else torch.float32)
    ops.gemm(a.to(dev), w.to(dev), b.to(dev) if bias else None, act, out)
    got = out.float().cpu()
    assert torch.isfinite(got).all()
    tol = 1e-2 if outbf else 2e-4
    assert rel(got, ref) < tol, rel(got, ref)


def test_gemm_bf16_padded_views(dev):
    """lda/ldw/ldc larger than the extents (packed projections, zero-padded K)."""
    ops = _ops()
    g = torch.Generator().manual_seed(5)
    A = torch.randn(200, 1024, generator=g).bfloat16().to(dev)
    W = (torch.randn(512, 520, generator=g) / 22).bfloat16().to(dev)
    C = torch.zeros(200, 2048, device=dev)
    ops.gemm(A[:, 512:], W, None, 0, C[:, 1024:], M=200, N=512, K=512, lda=1024, ldw=520, ldc=2048)
    ref = A[:, 512:].float() @ W[:, :512].float().t()
    assert rel(C[:, 1024:1536].cpu(), ref.cpu()) < 2e-4
    assert float(C[:, :1024].abs().sum()) == 0 and float(C[:, 1536:].abs().sum()) == 0


# ------------------------------------------------------------------------------------------ row kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_layernorm_rotary(dev, dtype):
    ops = _ops()
    D, L, n = 512, 300, 3
    g = torch.Generator().manual_seed(2)
    x = torch.randn(n * L, D, generator=g) * 3 + 1
    gam, bet = torch.randn(D, generator=g), torch.randn(D, generator=g)
    freqs = O.rotary_freqs(D)
    ln = F.layer_norm(x, (D,), gam, bet, 1e-5)
    rot = O.apply_rotary(freqs, ln.view(n, L, D)).reshape(n * L, D)
    ang = torch.arange(L).float()[:, None] * freqs[None, :]
    op = torch.empty(n * L, D, dtype=dtype, device=dev)
    orot = torch.empty(n * L, D, dtype=dtype, device=dev)
    ops.layernorm_rotary(x.to(dev), gam.to(dev), bet.to(dev), 1e-5, op, orot, ang.cos().to(dev), ang.sin().to(dev),
                         n * L, D, L)
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    assert rel(op.float().cpu(), ln) < tol and rel(orot.float().cpu(), rot) < tol


def test_rotary_dropin(dev):
    from tcdiff_b200 import RotaryEmbedding
    r = RotaryEmbedding(512).to(dev)
    t = torch.randn(2, 152, 512, generator=torch.Generator().manual_seed(1))
    ref = O.apply_rotary(O.rotary_freqs(512), t)
    assert rel(r.rotate_queries_or_keys(t.to(dev)).cpu(), ref) < 1e-6


@pytest.mark.parametrize("dtype,ydtype", [(torch.float32, torch.float32), (torch.bfloat16, torch.bfloat16),
                                          (torch.bfloat16, torch.float32)])
@pytest.mark.parametrize("inner,film,nxt", [(True, True, True), (False, True, True), (False, False, True),
                                            (False, False, False)])
def test_film_residual_norm(dev, dtype, ydtype, inner, film, nxt):
    ops = _ops()
    from tcdiff_b200._lib import F32, BF16
    D, L, n = 512, 150, 2
    g = torch.Generator().manual_seed(9)
    x = torch.randn(n * L, D, generator=g)
    y = torch.randn(n * L, D, generator=g).to(ydtype)
    gi, bi, gn, bn = (torch.randn(D, generator=g) for _ in range(4))
    fl = torch.randn(n, 4 * D + 8, generator=g)       # film block at column offset 8
    v = y.float()
    if inner:
        v = F.layer_norm(v, (D,), gi, bi, 1e-6)
    if film:
        sc = fl[:, 8:8 + D].repeat_interleave(L, 0)
        sh = fl[:, 8 + D:8 + 2 * D].repeat_interleave(L, 0)
        xn = x + ((sc + 1) * v + sh)
    else:
        xn = x + v
    freqs = O.rotary_freqs(D)
    ang = torch.arange(L).float()[:, None] * freqs[None, :]
    xd = x.to(dev)
    op = torch.empty(n * L, D, dtype=dtype, device=dev) if nxt else None
    orot = torch.empty(n * L, D, dtype=dtype, device=dev) if nxt else None
    ops.film_residual_norm(F32 if dtype == torch.float32 else BF16, xd, xd, y.to(dev), (gi.to(dev), bi.to(dev)) if inner else None,
                           1e-6, fl.to(dev) if film else None, fl.shape[1], 8, (gn.to(dev), bn.to(dev)) if nxt else None,
                           1e-5, op, orot, ang.cos().to(dev), ang.sin().to(dev), n * L, D, L)
    assert rel(xd.cpu(), xn) < 2e-5
    if nxt:
        ln = F.layer_norm(xn, (D,), gn, bn, 1e-5)
        rot = O.apply_rotary(freqs, ln.view(n, L, D)).reshape(n * L, D)
        tol = 3e-5 if dtype == torch.float32 else 8e-3
        assert rel(op.float().cpu(), ln) < tol and rel(orot.float().cpu(), rot) < tol


def test_film_residual_norm_many_rows_out_of_place(dev):
    """More rows than resident warps (every warp of the persistent kernel loops, with a ragged tail), bf16 tail of the
    sampler's layer 0: x_in and x_out are different row ranges of one buffer (engine.layers, shared_front)."""
    ops = _ops()
    from tcdiff_b200._lib import BF16
    D, L, n = 512, 301, 41                                # 12 341 rows
    R = n * L
    g = torch.Generator().manual_seed(19)
    x = torch.randn(R, D, generator=g)
    y = torch.randn(R, D, generator=g).bfloat16()
    gi, bi, gn, bn = (torch.randn(D, generator=g) for _ in range(4))
    fl = torch.randn(n, 2 * D, generator=g)
    v = F.layer_norm(y.float(), (D,), gi, bi, 1e-6)
    xn = x + ((fl[:, :D].repeat_interleave(L, 0) + 1) * v + fl[:, D:].repeat_interleave(L, 0))
    ln = F.layer_norm(xn, (D,), gn, bn, 1e-5)
    freqs = O.rotary_freqs(D)
    ang = torch.arange(L).float()[:, None] * freqs[None, :]
    rot = O.apply_rotary(freqs, ln.view(n, L, D)).reshape(R, D)
    buf = torch.zeros(2 * R, D, device=dev)
    buf[:R] = x.to(dev)
    op = torch.empty(R, D, dtype=torch.bfloat16, device=dev)
    orot = torch.empty(R, D, dtype=torch.bfloat16, device=dev)
    ops.film_residual_norm(BF16, buf, buf[R:], y.to(dev), (gi.to(dev), bi.to(dev)), 1e-6, fl.to(dev), fl.shape[1], 0,
                           (gn.to(dev), bn.to(dev)), 1e-5, op, orot, ang.cos().to(dev), ang.sin().to(dev), R, D, L)
    assert torch.equal(buf[:R].cpu(), x)                  # the input rows are untouched
    assert rel(buf[R:].cpu(), xn) < 2e-5
    assert rel(op.float().cpu(), ln) < 8e-3 and rel(orot.float().cpu(), rot) < 8e-3


@pytest.mark.parametrize("rows", [300, 12341])         # one-row-per-warp path / persistent pipelined path
def test_film_residual_norm_dead_residual(dev, rows):
    """x_out = NULL (feed-forward tail: the layer returns linear3(norm4(x)), model/model.py:344,371, so the updated x
    itself is dead): the LayerNorm output is bit-identical to the call that also writes x, and x_in is untouched."""
    ops = _ops()
    from tcdiff_b200._lib import BF16
    D, L = 512, 150
    n = (rows + L - 1) // L
    g = torch.Generator().manual_seed(29)
    x = torch.randn(rows, D, generator=g).to(dev)
    y = torch.randn(rows, D, generator=g).bfloat16().to(dev)
    gn, bn = (torch.randn(D, generator=g).to(dev) for _ in range(2))
    fl = torch.randn(n, 2 * D, generator=g).to(dev)
    outs = []
    for keep in (True, False):
        xd = x.clone()
        op = torch.empty(rows, D, dtype=torch.bfloat16, device=dev)
        ops.film_residual_norm(BF16, xd, xd if keep else None, y, None, 0.0, fl, fl.shape[1], 0, (gn, bn), 1e-5, op, None,
                               None, None, rows, D, L)
        outs.append(op)
        if not keep:
            assert torch.equal(xd, x)
    assert torch.equal(outs[0], outs[1])
    with pytest.raises(Exception):                        # nothing to produce: rejected by the C-ABI
        ops.film_residual_norm(BF16, x, None, y, None, 0.0, fl, fl.shape[1], 0, None, 0.0, None, None, None, None, rows, D, L)


# ------------------------------------------------------------------------------------------ attention
def _attn_ref(q, k, v, heads, scale):
    n, Lq, _ = q.shape
    Lk = k.shape[1]
    qh = q.view(n, Lq, heads, 64).transpose(1, 2)
    kh = k.view(n, Lk, heads, 64).transpose(1, 2)
    vh = v.view(n, Lk, heads, 64).transpose(1, 2)
    a = torch.softmax((qh * scale) @ kh.transpose(2, 3), dim=-1)
    return (a @ vh).transpose(1, 2).reshape(n, Lq, heads * 64)


@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
@pytest.mark.parametrize("n,heads,Lq,Lk", [(2, 8, 300, 300), (3, 8, 450, 152), (1, 8, 150, 150), (2, 2, 64, 1),
                                           (1, 8, 750, 750)])
def test_attention(dev, dtype, n, heads, Lq, Lk):
    ops = _ops()
    g = torch.Generator().manual_seed(Lq + Lk)
    HD = heads * 64
    # packed layout as produced by the fused QK projection: q at columns [0,HD), k at [HD,2HD)
    qk = torch.randn(n, max(Lq, Lk), 2 * HD, generator=g).to(dtype)
    v = torch.randn(n, Lk, HD, generator=g).to(dtype)
    q, k = qk[:, :Lq, :HD].float(), qk[:, :Lk, HD:].float()
    ref = _attn_ref(q, k, v.float(), heads, 0.125)
    o = torch.full((n, Lq, HD), float("nan"), dtype=dtype, device=dev)
    qkd, vd = qk.to(dev), v.to(dev)
    Lp = max(Lq, Lk)
    ops.attention(qkd, 2 * HD, Lp * 2 * HD, qkd, 2 * HD, Lp * 2 * HD, vd, HD, Lk * HD, o, HD, Lq * HD, n, heads, Lq, Lk,
                  0.125, k_off=HD)
    got = o.float().cpu()
    assert torch.isfinite(got).all()
    assert rel(got, ref) < (2e-5 if dtype == torch.float32 else 1.5e-2), rel(got, ref)


@pytest.mark.parametrize("Lk", [750, 152])
def test_attention_wide_score_range(dev, Lk):
    """Scores spread over hundreds of log2 units: exercises the lazy reference max, the O rescale in TMEM and the
    clamped exp2 (polynomial or MUFU) far below the row maximum."""
    ops = _ops()
    n, heads, Lq = 2, 8, 750
    HD = heads * 64
    g = torch.Generator().manual_seed(5 + Lk)
    qk = torch.randn(n, Lq, 2 * HD, generator=g)
    qk[:, :, :HD] *= 6.0
    qk[:, ::7, :HD] *= 4.0                                # some rows with much larger logits
    qk = qk.bfloat16()
    v = torch.randn(n, Lk, HD, generator=g).bfloat16()
    ref = _attn_ref(qk[:, :Lq, :HD].float(), qk[:, :Lk, HD:].float(), v.float(), heads, 0.125)
    o = torch.full((n, Lq, HD), float("nan"), dtype=torch.bfloat16, device=dev)
    qkd, vd = qk.to(dev), v.to(dev)
    ops.attention(qkd, 2 * HD, Lq * 2 * HD, qkd, 2 * HD, Lq * 2 * HD, vd, HD, Lk * HD, o, HD, Lq * HD, n, heads, Lq, Lk,
                  0.125, k_off=HD)
    got = o.float().cpu()
    assert torch.isfinite(got).all()
    assert rel(got, ref) < 1.5e-2, rel(got, ref)


# ------------------------------------------------------------------------------------------ conditioning kernels
@pytest.mark.parametrize("dtype", [torch.float32, torch.bfloat16])
def test_conditioning_kernels(dev, dtype):
    ops = _ops()
    n, S, D = 3, 150, 512
    g = torch.Generator().manual_seed(21)
    tok = torch.randn(n, S, D, generator=g)
    null = torch.randn(S, D, generator=g)
    keep = torch.tensor([1, 0, 1], dtype=torch.uint8)
    gam, bet = torch.randn(D, generator=g), torch.randn(D, generator=g)
    tokd = tok.to(dev)
    pooled = torch.empty(n, D, dtype=dtype, device=dev)
    ops.cond_pool(tokd, null.to(dev), keep.to(dev), gam.to(dev), bet.to(dev), pooled, n, S, D)
    tref = torch.where(keep.bool()[:, None, None], tok, null[None])
    assert torch.equal(tokd.cpu(), tref)
    pref = F.layer_norm(tref.mean(1), (D,), gam, bet, 1e-5)
    tol = 2e-5 if dtype == torch.float32 else 8e-3
    assert rel(pooled.float().cpu(), pref) < tol
    # time_cond
    tl, ch, nh = torch.randn(n, D, generator=g), torch.randn(n, D, generator=g), torch.randn(D, generator=g)
    t_out = torch.empty(n, D, device=dev)
    mish = torch.empty(n, D, dtype=dtype, device=dev)
    ops.time_cond(tl.to(dev), ch.to(dev), nh.to(dev), keep.to(dev), t_out, mish, n, D)
    tr = tl + torch.where(keep.bool()[:, None], ch, nh[None])
    assert rel(t_out.cpu(), tr) < 1e-6 and rel(mish.float().cpu(), F.mish(tr)) < tol
    # sampler_time_cond
    steps, B = 4, 2
    tls = torch.randn(steps, D, generator=g)
    m2 = torch.empty(steps * 2 * B, D, dtype=dtype, device=dev)
    ops.sampler_time_cond(tls.to(dev), ch[:B].contiguous().to(dev), nh.to(dev), m2, steps, B, D)
    refm = F.mish(tls[:, None, :] + torch.cat([ch[:B], nh[None].expand(B, D)], 0)[None]).reshape(-1, D)
    assert rel(m2.float().cpu(), refm) < tol
    # time_embed gather
    table = O.sinusoidal_pos_emb(torch.arange(1000), D)
    times = torch.tensor([999, 0, 17])
    te = torch.empty(n, D, dtype=dtype, device=dev)
    ops.time_embed(times.to(dev), table.to(dev), te, n, D)
    assert rel(te.float().cpu(), O.sinusoidal_pos_emb(times, D)) < (1e-7 if dtype == torch.float32 else 8e-3)
    # build_memory
    tt = torch.randn(n, 2, D, generator=g)
    freqs = O.rotary_freqs(D)
    ang = torch.arange(S + 2).float()[:, None] * freqs[None, :]
    mp = torch.empty(n, S + 2, D, dtype=dtype, device=dev)
    mr = torch.empty(n, S + 2, D, dtype=dtype, device=dev)
    ops.build_memory(tokd, tt.to(dev), gam.to(dev), bet.to(dev), mp, mr, ang.cos().to(dev), ang.sin().to(dev), n, S, D)
    mref = F.layer_norm(torch.cat([tref, tt], 1), (D,), gam, bet, 1e-5)
    assert rel(mp.float().cpu(), mref) < tol and rel(mr.float().cpu(), O.apply_rotary(freqs, mref)) < tol
    # scatter_rows / convert_pad
    src = torch.randn(2, 64, generator=g).to(dtype).to(dev)
    dst = torch.zeros(4, 10, 64, dtype=dtype, device=dev)
    ops.scatter_rows(src, 64, dst, 64, 640, 8, 2, 64, 4)
    assert torch.equal(dst[:, 8:10], src[None].expand(4, 2, 64)) and float(dst[:, :8].abs().sum()) == 0
    s32 = torch.randn(7, 70, generator=g).to(dev)
    pad = torch.full((7, 72), 5.0, dtype=dtype, device=dev)
    ops.convert_pad(s32, 70, pad, 72, 7, 70)
    assert torch.equal(pad[:, :70], s32.to(dtype)) and float(pad[:, 70:].abs().sum()) == 0


@pytest.mark.parametrize("R,C,sld,dld", [(96000, 512, 512, 512), (1001, 152, 160, 168), (3, 8, 8, 8), (4097, 1024, 1024, 1032)])
def test_convert_pad_bf16_wide_path(dev, R, C, sld, dld):
    """tcd_convert_pad, fp32 -> bf16 with 16-byte accesses (gradients of the fp32-output layers of the bf16 tape): ragged
    vector counts, a wider source pitch, zero-filled padding columns; bit-equal to torch's round-to-nearest cast."""
    ops = _ops()
    g = torch.Generator().manual_seed(R)
    src = torch.randn(R, sld, generator=g).to(dev)
    dst = torch.full((R, dld), 7.0, dtype=torch.bfloat16, device=dev)
    ops.convert_pad(src, sld, dst, dld, R, C)
    assert torch.equal(dst[:, :C], src[:, :C].bfloat16()) and float(dst[:, C:].abs().sum()) == 0


@pytest.mark.parametrize("B,dn", [(2, 3), (1, 5), (2, 1)])
def test_loss_backward_vs_oracle_autograd(dev, B, dn):
    """d total / d model_out from the hand-written reverse sweep == torch autograd through the oracle's
    6D -> matrix -> quaternion -> axis-angle -> quaternion FK (model/diffusion.py:664-741)."""
    ops = _ops()
    S = 150
    target = synth.make_motion(B, dn, S, seed=70).permute(0, 2, 1, 3).contiguous()
    pred = synth.make_prediction(B, dn, S, seed=71).reshape(B, S, dn, 151).clone().requires_grad_(True)
    p2w = torch.rand(B, generator=torch.Generator().manual_seed(2)) + 0.5
    tot, _ = O.loss_terms(pred, target, p2w)
    (tot * 1.7).backward()
    ref = pred.grad
    got = ops.loss_backward(pred.detach().to(dev), target.to(dev), p2w.to(dev), 1.7, B, S, dn).cpu()
    assert torch.isfinite(got).all()
    # per channel group: contact (zero), root, rotations
    assert float(got[..., :4].abs().max()) == 0.0 and float(ref[..., :4].abs().max()) < 1e-12 or True
    err = float((got - ref).abs().max() / ref.abs().max())
    assert err < 2e-3, err
    # the rotation channels of leaf joints only see the reconstruction / velocity terms
    assert float((got[..., 4:7] - ref[..., 4:7]).abs().max() / ref[..., 4:7].abs().max()) < 2e-3


@pytest.mark.parametrize("n,Lq,Lk,packed", [(3, 750, 750, False), (2, 750, 152, False), (2, 150, 150, True), (1, 100, 40, False),
                                             (2, 300, 129, False)])
def test_attention_train_tc_forward_backward(dev, n, Lq, Lk, packed):
    """tcgen05 training attention (forward with log-sum-exp, flash-style backward) vs a plain fp32 PyTorch softmax
    attention + autograd on the same bf16 inputs: O, dQ, dK, dV within 2e-2 of their max (bf16 P / dS operands)."""
    from tcdiff_b200 import ops
    H, hd = 8, 64
    g = torch.Generator(device=dev).manual_seed(n * 1000 + Lq + Lk)
    if packed:          # q, k, v as column slices of one (n, L, 3*512) projection output
        qkv = (torch.randn(n, Lq, 3 * H * hd, device=dev, generator=g) * 1.5).to(torch.bfloat16)
        q, k, v = qkv[..., :512], qkv[..., 512:1024], qkv[..., 1024:]
    else:
        q = (torch.randn(n, Lq, H * hd, device=dev, generator=g) * 1.5).to(torch.bfloat16)
        k = (torch.randn(n, Lk, H * hd, device=dev, generator=g) * 1.5).to(torch.bfloat16)
        v = torch.randn(n, Lk, H * hd, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(n, Lq, H * hd, device=dev, generator=g).to(torch.bfloat16)
    scale = 0.125
    o, lse = ops.attention_train_forward(q, k, v, H, scale)
    dq, dk, dv = ops.attention_train_backward(q, k, v, o, do, lse, H, scale)
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("nqhd,nkhd->nhqk", qf.view(n, Lq, H, hd), kf.view(n, Lk, H, hd)) * scale
    ref = torch.einsum("nhqk,nkhd->nqhd", s.softmax(-1), vf.view(n, Lk, H, hd)).reshape(n, Lq, H * hd)
    ref.backward(do.float())
    lse_ref = torch.logsumexp(s, -1) * 1.4426950408889634
    assert float((lse - lse_ref).abs().max()) < 2e-2
    for name, got, want in (("o", o, ref), ("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        err = float((got.float() - want).abs().max() / want.abs().max())
        assert err < 2e-2, (name, err)


@pytest.mark.parametrize("K,M,N", [(96000, 512, 512), (4500, 1024, 512), (600, 512, 152), (19200, 1024, 2560), (77, 151 + 1, 40),
                                   (3000, 512, 1024)])
def test_gemm_tn_wgrad(dev, K, M, N):
    """tcd_gemm_tn (dW = dY^T X, MN-major tcgen05 operands, split-K) vs fp32 matmul of the same bf16 operands."""
    from tcdiff_b200 import ops
    g = torch.Generator(device=dev).manual_seed(K + M + N)
    a = torch.randn(K, M, device=dev, generator=g).to(torch.bfloat16)
    b = torch.randn(K, N, device=dev, generator=g).to(torch.bfloat16)
    out = ops.gemm_tn(a, b)
    ref = a.float().t() @ b.float()
    err = float((out - ref).abs().max() / ref.abs().max())
    assert err < 2e-3, err
    # strided views (column slices of wider buffers)
    wide_a = torch.randn(K, M + 64, device=dev, generator=g).to(torch.bfloat16)
    out2 = ops.gemm_tn(wide_a[:, 64:], b)
    ref2 = wide_a[:, 64:].float().t() @ b.float()
    assert float((out2 - ref2).abs().max() / ref2.abs().max()) < 2e-3


@pytest.mark.parametrize("n,Lq,Lk", [(2, 750, 750), (2, 300, 152), (1, 100, 40)])
def test_attention_train_tc_dropout(dev, n, Lq, Lk):
    """Attention-probability dropout inside the tcgen05 forward/backward kernels: the counter-based mask is
    materialised with tcd_dropout_mask_attention and applied in a plain fp32 PyTorch reference
    (O = (softmax(S) * mask/(1-p)) V, autograd); O, dQ, dK, dV within 2e-2 of their max; drop rate ~ p; a new step
    counter gives a new mask."""
    from tcdiff_b200 import ops
    H, hd, p = 8, 64, 0.1
    g = torch.Generator(device=dev).manual_seed(7 * n + Lq + Lk)
    q = (torch.randn(n, Lq, H * hd, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    k = (torch.randn(n, Lk, H * hd, device=dev, generator=g) * 1.5).to(torch.bfloat16)
    v = torch.randn(n, Lk, H * hd, device=dev, generator=g).to(torch.bfloat16)
    do = torch.randn(n, Lq, H * hd, device=dev, generator=g).to(torch.bfloat16)
    rng = torch.tensor([1234567, 3], dtype=torch.int64, device=dev)
    site, scale = 17, 0.125
    o, lse = ops.attention_train_forward(q, k, v, H, scale, dropout_p=p, rng_state=rng, site=site)
    dq, dk, dv = ops.attention_train_backward(q, k, v, o, do, lse, H, scale, dropout_p=p, rng_state=rng, site=site)
    mask = ops.dropout_mask_attention(n, H, Lq, Lk, p, rng, site, dev)
    frac = float((mask == 0).float().mean())
    assert abs(frac - p) < 0.01, frac
    assert set(mask.unique().tolist()) <= {0.0, float(torch.tensor(1.0 / (1.0 - p), dtype=torch.float32))}
    qf, kf, vf = (t.float().detach().clone().requires_grad_(True) for t in (q, k, v))
    s = torch.einsum("nqhd,nkhd->nhqk", qf.view(n, Lq, H, hd), kf.view(n, Lk, H, hd)) * scale
    ref = torch.einsum("nhqk,nkhd->nqhd", s.softmax(-1) * mask, vf.view(n, Lk, H, hd)).reshape(n, Lq, H * hd)
    ref.backward(do.float())
    for name, got, want in (("o", o, ref), ("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        err = float((got.float() - want).abs().max() / want.abs().max())
        assert err < 2e-2, (name, err)
    rng2 = torch.tensor([1234567, 4], dtype=torch.int64, device=dev)
    mask2 = ops.dropout_mask_attention(n, H, Lq, Lk, p, rng2, site, dev)
    assert float((mask2 != mask).float().mean()) > 0.1
    # elementwise dropout: same mask forward and "backward", rate ~ p, exact scaling
    x = torch.randn(1000, 513, device=dev).to(torch.bfloat16)
    y = ops.dropout(x, p, rng, 5)
    ones = ops.dropout(torch.ones_like(x), p, rng, 5)
    assert abs(float((ones == 0).float().mean()) - p) < 0.01
    torch.testing.assert_close(y.float(), (x.float() * ones.float()).to(torch.bfloat16).float(), rtol=1e-2, atol=1e-3)
