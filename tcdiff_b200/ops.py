"""Thin tensor->pointer wrappers over the C-ABI.  torch is used for device memory and streams only."""
import torch

from . import _lib
from ._lib import F32, BF16, ACT_NONE, ACT_RELU, ACT_GELU, ACT_MISH, ACT_SILU, check  # noqa: F401

_DT = {torch.float32: F32, torch.bfloat16: BF16}


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _ptr(t):
    return 0 if t is None else t.data_ptr()


def _cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise _lib.TcdError("tcdiff_b200 kernels need CUDA tensors (no CPU fallback); got device %s" % t.device)


def dt(t):
    return _DT[t.dtype]


def gemm(a, w, bias, act, out, M=None, N=None, K=None, lda=None, ldw=None, ldc=None):
    """out[:M,:N] = act(a[:M,:K] @ w[:N,:K]^T + bias); a, w 2-D views with unit inner stride."""
    _cuda(a, w, out)
    M = a.shape[0] if M is None else M
    K = a.shape[1] if K is None else K
    N = w.shape[0] if N is None else N
    lda = a.stride(0) if lda is None else lda
    ldw = w.stride(0) if ldw is None else ldw
    ldc = out.stride(0) if ldc is None else ldc
    assert a.dtype == w.dtype and a.stride(-1) == 1 and w.stride(-1) == 1 and out.stride(-1) == 1
    check(_lib.lib().tcd_gemm(dt(a), a.data_ptr(), lda, w.data_ptr(), ldw, _ptr(bias), act, dt(out), out.data_ptr(),
                              ldc, M, N, K, _stream()))
    return out


def gemm_tn(a, b, out=None):
    """out (M, N) fp32 = a^T @ b for bf16 a (K, M), b (K, N) row-major (pitches multiples of 8): the weight-gradient
    contraction dW = dY^T X on the tensor cores, activations read as stored."""
    _cuda(a, b)
    K, M = a.shape
    N = b.shape[1]
    assert b.shape[0] == K and a.dtype == b.dtype == torch.bfloat16 and a.stride(1) == 1 and b.stride(1) == 1
    lib = _lib.lib()
    out = torch.empty(M, N, dtype=torch.float32, device=a.device) if out is None else out
    ws = torch.empty(lib.tcd_gemm_tn_workspace_floats(M, N, K), dtype=torch.float32, device=a.device)
    check(lib.tcd_gemm_tn(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), out.data_ptr(), out.stride(0), M, N, K,
                          ws.data_ptr(), _stream()))
    return out


def layernorm_rotary(x, gamma, beta, eps, out_plain, out_rot, rot_cos, rot_sin, rows, D, tps):
    o = out_plain if out_plain is not None else out_rot
    check(_lib.lib().tcd_layernorm_rotary(dt(o), x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), eps, _ptr(out_plain),
                                          _ptr(out_rot), _ptr(rot_cos), _ptr(rot_sin), rows, D, tps, _stream()))


def rotary(x, out, rot_cos, rot_sin, rows, D, tps):
    _cuda(x, out)
    check(_lib.lib().tcd_rotary(x.data_ptr(), out.data_ptr(), rot_cos.data_ptr(), rot_sin.data_ptr(), rows, D, tps,
                                _stream()))


def film_residual_norm(op_dtype, x_in, x_out, y, ln_in, eps_in, film, film_ld, film_off, ln_next, eps_next, out_plain,
                       out_rot, rot_cos, rot_sin, rows, D, tps):
    gi, bi = ln_in if ln_in is not None else (None, None)
    gn, bn = ln_next if ln_next is not None else (None, None)
    check(_lib.lib().tcd_film_residual_norm(op_dtype, x_in.data_ptr(), _ptr(x_out), y.data_ptr(), dt(y), _ptr(gi),
                                            _ptr(bi), eps_in, _ptr(film), film_ld, film_off, _ptr(gn), _ptr(bn),
                                            eps_next, _ptr(out_plain), _ptr(out_rot), _ptr(rot_cos), _ptr(rot_sin),
                                            rows, D, tps, _stream()))


def gemm_film_residual_norm(a, w, bias, x_in, x_out, ln_in, eps_in, film, film_ld, film_off, ln_next, eps_next, out_plain,
                            out_rot, rot_cos_t, rot_sin_t, rows, tps):
    """Fused `fc` / `linear2` GEMM + FiLM residual tail (csrc/gemm_frn.cu): a (rows, K) bf16, w (512, K) bf16; rot_cos_t /
    rot_sin_t: the rotary tables TRANSPOSED, (256 angles, >= tps positions) fp32 (engine.PackedWeights.rot_cos_t)."""
    gi, bi = ln_in if ln_in is not None else (None, None)
    gn, bn = ln_next if ln_next is not None else (None, None)     # None: LN_next's affine is folded into the consumer's weights
    rot_ld = rot_cos_t.stride(0) if rot_cos_t is not None else 0
    check(_lib.lib().tcd_gemm_film_residual_norm(a.data_ptr(), a.stride(0), w.data_ptr(), w.stride(0), _ptr(bias), rows,
                                                 a.shape[1], _ptr(x_in), _ptr(x_out), _ptr(gi), _ptr(bi), eps_in,
                                                 _ptr(film), film_ld, film_off, _ptr(gn), _ptr(bn), eps_next,
                                                 _ptr(out_plain), _ptr(out_rot), _ptr(rot_cos_t), _ptr(rot_sin_t), rot_ld, tps,
                                                 _stream()))


def attention(q, ldq, qbs, k, ldk, kbs, v, ldv, vbs, o, ldo, obs, samples, heads, Lq, Lk, scale, q_off=0, k_off=0,
              v_off=0):
    """q/k/v/o are base tensors; *_off are element offsets into them (column offsets of packed projections)."""
    es = q.element_size()
    check(_lib.lib().tcd_attention(dt(q), q.data_ptr() + q_off * es, ldq, qbs, k.data_ptr() + k_off * es, ldk, kbs,
                                   v.data_ptr() + v_off * es, ldv, vbs, o.data_ptr(), ldo, obs, samples, heads, Lq, Lk,
                                   scale, _stream()))


def attention_train_forward(q, k, v, heads, scale, dropout_p=0.0, rng_state=None, site=0):
    """bf16 (n, Lq, H*64) / (n, Lk, H*64) tensors (any row pitch / batch stride that is a multiple of 8, last dim
    contiguous) -> (o bf16 contiguous, lse fp32 (n, H, Lq)) on the tcgen05 kernel."""
    n, Lq, HD = q.shape
    Lk = k.shape[1]
    o = torch.empty(n, Lq, HD, dtype=torch.bfloat16, device=q.device)
    lse = torch.empty(n, heads, Lq, dtype=torch.float32, device=q.device)
    check(_lib.lib().tcd_attention_train_forward(q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1),
                                                 k.stride(0), v.data_ptr(), v.stride(1), v.stride(0), o.data_ptr(), HD,
                                                 Lq * HD, lse.data_ptr(), n, heads, Lq, Lk, scale, dropout_p, _ptr(rng_state),
                                                 site, _stream()))
    return o, lse


def attention_train_backward(q, k, v, o, do, lse, heads, scale, dq=None, dk=None, dv=None, dropout_p=0.0, rng_state=None,
                             site=0):
    """dQ, dK, dV (bf16) of the attention above; optional preallocated outputs may be strided views."""
    n, Lq, HD = q.shape
    Lk = k.shape[1]
    lib = _lib.lib()
    dq = torch.empty(n, Lq, HD, dtype=torch.bfloat16, device=q.device) if dq is None else dq
    dk = torch.empty(n, Lk, HD, dtype=torch.bfloat16, device=q.device) if dk is None else dk
    dv = torch.empty(n, Lk, HD, dtype=torch.bfloat16, device=q.device) if dv is None else dv
    ws = torch.empty(lib.tcd_attention_train_workspace_floats(n, heads, Lq), dtype=torch.float32, device=q.device)
    check(lib.tcd_attention_train_backward(
        q.data_ptr(), q.stride(1), q.stride(0), k.data_ptr(), k.stride(1), k.stride(0), v.data_ptr(), v.stride(1),
        v.stride(0), o.data_ptr(), o.stride(1), o.stride(0), do.data_ptr(), do.stride(1), do.stride(0), lse.data_ptr(),
        dq.data_ptr(), dq.stride(1), dq.stride(0), dk.data_ptr(), dk.stride(1), dk.stride(0), dv.data_ptr(), dv.stride(1),
        dv.stride(0), ws.data_ptr(), n, heads, Lq, Lk, scale, dropout_p, _ptr(rng_state), site, _stream()))
    return dq, dk, dv


def dropout(x, p, rng_state, site, out=None):
    """y = x * keep / (1 - p) with the counter-based mask of (rng_state, site); x contiguous bf16 / fp32."""
    _cuda(x)
    out = torch.empty_like(x) if out is None else out
    check(_lib.lib().tcd_dropout(dt(x), x.data_ptr(), out.data_ptr(), x.numel(), p, rng_state.data_ptr(), site, _stream()))
    return out


def dropout_mask_attention(n, heads, Lq, Lk, p, rng_state, site, device):
    out = torch.empty(n, heads, Lq, Lk, dtype=torch.float32, device=device)
    check(_lib.lib().tcd_dropout_mask_attention(out.data_ptr(), n, heads, Lq, Lk, p, rng_state.data_ptr(), site, _stream()))
    return out


def time_embed(times, table, out, n, D):
    check(_lib.lib().tcd_time_embed(dt(out), times.data_ptr(), table.data_ptr(), out.data_ptr(), n, D, table.shape[0],
                                    _stream()))


def cond_pool(tokens, null_embed, keep, gamma, beta, pooled, n, S, D):
    check(_lib.lib().tcd_cond_pool(dt(pooled), tokens.data_ptr(), null_embed.data_ptr(), keep.data_ptr(),
                                   gamma.data_ptr(), beta.data_ptr(), pooled.data_ptr(), n, S, D, _stream()))


def time_cond(t_lin, cond_hidden, null_hidden, keep, t_out, mish_out, n, D):
    check(_lib.lib().tcd_time_cond(dt(mish_out), t_lin.data_ptr(), cond_hidden.data_ptr(), null_hidden.data_ptr(),
                                   keep.data_ptr(), _ptr(t_out), mish_out.data_ptr(), n, D, _stream()))


def sampler_time_cond(t_lin, ch_cond, ch_uncond, mish_out, steps, B, D):
    check(_lib.lib().tcd_sampler_time_cond(dt(mish_out), t_lin.data_ptr(), ch_cond.data_ptr(), ch_uncond.data_ptr(),
                                           mish_out.data_ptr(), steps, B, D, _stream()))


def build_memory(tokens, t_tokens, gamma, beta, mem_plain, mem_rot, rot_cos, rot_sin, n, S, D):
    check(_lib.lib().tcd_build_memory(dt(mem_plain), tokens.data_ptr(), t_tokens.data_ptr(), gamma.data_ptr(),
                                      beta.data_ptr(), mem_plain.data_ptr(), mem_rot.data_ptr(), rot_cos.data_ptr(),
                                      rot_sin.data_ptr(), n, S, D, _stream()))


def scatter_rows(src, src_ld, dst, dst_ld, dst_batch_stride, dst_row0, rows, cols, samples, src_off=0, dst_off=0,
                 src_batch_stride=0):
    es = src.element_size()
    check(_lib.lib().tcd_scatter_rows(dt(src), src.data_ptr() + src_off * es, src_ld, src_batch_stride,
                                      dst.data_ptr() + dst_off * es, dst_ld, dst_batch_stride, dst_row0, rows, cols,
                                      samples, _stream()))


def masked_blend(x, value, weight, B, S, dn):
    _cuda(x, value, weight)
    check(_lib.lib().tcd_masked_blend(x.data_ptr(), value.data_ptr(), weight.data_ptr(), B, S, dn, 151, _stream()))


def convert_pad(src, src_ld, dst, dst_ld, rows, cols):
    check(_lib.lib().tcd_convert_pad(dt(dst), src.data_ptr(), src_ld, dst.data_ptr(), dst_ld, rows, cols, _stream()))


def cfg_ddim_step(x, out_cond, out_uncond, noise, traj, x_out, x0_out, xpad, xpad_ld, n_tokens, w, sr, srm1, sa, c,
                  sigma, clip, last, rng=None, rng_stream=0):
    """noise: the step's pre-drawn (n_tokens, 151) tensor, or None with rng = device int64 {seed, counter} pair: the draw
    is then generated inside the kernel (Philox, draw id `rng_stream`)."""
    _cuda(x, out_cond, out_uncond, x_out)
    if noise is None and rng is not None:
        check(_lib.lib().tcd_cfg_ddim_step_rng(x.data_ptr(), out_cond.data_ptr(), out_uncond.data_ptr(), rng.data_ptr(),
                                               rng_stream, _ptr(traj), x_out.data_ptr(), _ptr(x0_out), _ptr(xpad), xpad_ld,
                                               n_tokens, 151, w, sr, srm1, sa, c, sigma, int(clip), int(last), _stream()))
        return
    check(_lib.lib().tcd_cfg_ddim_step(x.data_ptr(), out_cond.data_ptr(), out_uncond.data_ptr(), _ptr(noise), _ptr(traj),
                                       x_out.data_ptr(), _ptr(x0_out), _ptr(xpad), xpad_ld, n_tokens, 151, w, sr, srm1,
                                       sa, c, sigma, int(clip), int(last), _stream()))


def cfg_ddpm_step(x, out_cond, out_uncond, noise, x_out, xpad, xpad_ld, n_tokens, w, c1, c2, std, nonzero, mask=None,
                  value_q=None, rng=None, rng_stream=0, eps_coef=None):
    """eps_coef = (sqrt_recip_alphas_cumprod[t], sqrt_recipm1_alphas_cumprod[t]) when the network predicts the noise
    (predict_epsilon=True, model/diffusion.py:176-187), None when it predicts x0."""
    pe, sr, srm1 = (0, 0.0, 0.0) if eps_coef is None else (1, float(eps_coef[0]), float(eps_coef[1]))
    _cuda(x, out_cond, out_uncond, x_out)
    if noise is None:
        if rng is None:
            raise _lib.TcdError("cfg_ddpm_step: either a noise tensor or an rng state is required")
        check(_lib.lib().tcd_cfg_ddpm_step_rng(x.data_ptr(), out_cond.data_ptr(), out_uncond.data_ptr(), rng.data_ptr(),
                                               rng_stream, x_out.data_ptr(), _ptr(xpad), xpad_ld, n_tokens, 151, w, c1, c2,
                                               std, int(nonzero), pe, sr, srm1, _ptr(mask), _ptr(value_q), _stream()))
        return
    _cuda(noise)
    check(_lib.lib().tcd_cfg_ddpm_step(x.data_ptr(), out_cond.data_ptr(), out_uncond.data_ptr(), noise.data_ptr(),
                                       x_out.data_ptr(), _ptr(xpad), xpad_ld, n_tokens, 151, w, c1, c2, std,
                                       int(nonzero), pe, sr, srm1, _ptr(mask), _ptr(value_q), _stream()))


def philox_normal(out, rng, rng_stream):
    """out (any shape, fp32, contiguous) = draw `rng_stream` of the sampler's counter-based N(0,1) stream."""
    _cuda(out, rng)
    assert out.dtype == torch.float32 and out.is_contiguous()
    check(_lib.lib().tcd_philox_normal(out.data_ptr(), out.numel(), rng.data_ptr(), rng_stream, _stream()))
    return out


def inpaint_traj(x, traj, xpad, xpad_ld, n_tokens):
    _cuda(x)
    check(_lib.lib().tcd_inpaint_traj(x.data_ptr(), _ptr(traj), _ptr(xpad), xpad_ld, n_tokens, 151, _stream()))


def q_sample(x_start, noise, t, sqrt_ac, sqrt_1mac, x_noisy, target, xpad, xpad_ld, B, dn, S, permute, restore_traj):
    _cuda(x_start, noise, t, x_noisy)
    check(_lib.lib().tcd_q_sample(x_start.data_ptr(), noise.data_ptr(), t.data_ptr(), sqrt_ac.data_ptr(),
                                  sqrt_1mac.data_ptr(), x_noisy.data_ptr(), _ptr(target), _ptr(xpad), xpad_ld, B, dn, S,
                                  151, int(permute), int(restore_traj), _stream()))


def ax_from_6v(d6, aa, n):
    _cuda(d6, aa)
    check(_lib.lib().tcd_ax_from_6v(d6.data_ptr(), aa.data_ptr(), n, _stream()))


def smpl_fk(aa, root, pos, n):
    _cuda(aa, root, pos)
    check(_lib.lib().tcd_smpl_fk(aa.data_ptr(), root.data_ptr(), pos.data_ptr(), n, _stream()))


def motion_fk(motion, pos, n):
    _cuda(motion, pos)
    check(_lib.lib().tcd_motion_fk(motion.data_ptr(), pos.data_ptr(), n, 151, _stream()))


def _loss_code(loss_type):
    """model/diffusion.py:172: F.mse_loss if loss_type == "l2" else F.l1_loss."""
    return _lib.LOSS_L2 if loss_type == "l2" else _lib.LOSS_L1


def loss_forward(model_out, target, p2w, B, S, dn, loss_type="l2"):
    _cuda(model_out, target)
    nws = _lib.lib().tcd_loss_workspace_floats(B, S, dn)
    ws = torch.empty(nws, dtype=torch.float32, device=model_out.device)
    out = torch.empty(5, dtype=torch.float32, device=model_out.device)
    check(_lib.lib().tcd_loss_forward(model_out.data_ptr(), target.data_ptr(), _ptr(p2w), ws.data_ptr(), out.data_ptr(),
                                      B, S, dn, _loss_code(loss_type), _stream()))
    return out


def loss_backward(model_out, target, p2w, grad_total, B, S, dn, loss_type="l2"):
    _cuda(model_out, target)
    g = torch.empty_like(model_out)
    check(_lib.lib().tcd_loss_backward(model_out.data_ptr(), target.data_ptr(), _ptr(p2w), float(grad_total), g.data_ptr(),
                                       B, S, dn, _loss_code(loss_type), _stream()))
    return g
