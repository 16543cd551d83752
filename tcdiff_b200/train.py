"""Training-mode forward/backward of the hot path on the sm_100a kernels.

`torch.autograd` is used for what it is: a tape.  Every node on the tape is a `torch.autograd.Function` whose forward
AND backward are C-ABI kernel calls (tcd_gemm for the layer, dgrad and wgrad contractions; tcd_layernorm_backward,
tcd_film_backward, tcd_act_backward, tcd_attention_backward, tcd_rotary with -theta, tcd_loss_backward ...), so
parameter gradients land in `param.grad` of the drop-in DanceDecoder and any optimizer / DDP wrapper works unchanged.

Two tapes:
  * fp32 (parity mode): fp32 activations, SIMT GEMM / attention kernels; dropout must be 0;
  * bf16 (default, further down): GEMM operands and their gradients are bf16 as stored, tcgen05 GEMM / wgrad /
    attention forward+backward kernels, fp32 residual stream; training-mode dropout (the reference trains with 0.1,
    TCDiff.py:82) uses counter-based masks recomputed in the backward kernels (csrc/dropout.cuh) — torch's own
    Bernoulli stream cannot be reproduced by another implementation, so parity with dropout is tested by injecting
    THESE masks into the CPU restatement (tests/test_gpu_train.py);
  * a few small conditioning-path reshapes/selects (mean over 150 music tokens, torch.where with the keep mask,
    concatenating the two time tokens) stay as torch ops on (B,150,512)-sized tensors.
Reference: model/model.py:548-624, model/diffusion.py:636-753.
"""
import math

import torch
from torch.autograd import Function

from . import _lib, ops
from ._lib import ACT_GELU, ACT_MISH, ACT_NONE, ACT_RELU, ACT_SILU, F32, check

HEAD_DIM = 64


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _up8(v):
    return (v + 7) // 8 * 8


def _cast(x, T):
    """(R, C) fp32 -> operand dtype with C padded to a multiple of 8 (zero padded)."""
    R, C = x.shape
    if T == torch.float32:
        return x, C
    Cp = _up8(C)
    out = torch.zeros(R, Cp, dtype=T, device=x.device) if Cp != C else torch.empty(R, Cp, dtype=T, device=x.device)
    ops.convert_pad(x, x.stride(0), out, Cp, R, C)
    return out, Cp


def _cast_t(x, T):
    """(R, C) fp32 -> (C, Rp) operand dtype, transposed, Rp = R padded to a multiple of 8 with zeros."""
    R, C = x.shape
    Rp = _up8(R)
    out = torch.zeros(C, Rp, dtype=T, device=x.device) if Rp != R else torch.empty(C, Rp, dtype=T, device=x.device)
    check(_lib.lib().tcd_cast_transpose(ops._DT[T], x.data_ptr(), x.stride(0), out.data_ptr(), Rp, R, C, _stream()))
    return out, Rp


def colsum(a, b=None):
    """Column sums of a (R, C) fp32 tensor (optionally of a*b), two-level for parallelism."""
    R, C = a.shape
    dev = a.device
    lib = _lib.lib()
    G = 256
    parts = []
    full = R // G
    bp = 0 if b is None else b.data_ptr()
    if full:
        p = torch.empty(full, C, device=dev)
        check(lib.tcd_group_colsum(a.data_ptr(), bp, a.stride(0), full, G, C, p.data_ptr(), C, 0, _stream()))
        parts.append(p)
    rem = R - full * G
    if rem:
        p = torch.empty(1, C, device=dev)
        off = full * G * a.stride(0) * 4
        check(lib.tcd_group_colsum(a.data_ptr() + off, 0 if b is None else b.data_ptr() + off, a.stride(0), 1, rem, C,
                                   p.data_ptr(), C, 0, _stream()))
        parts.append(p)
    p = parts[0] if len(parts) == 1 else torch.cat(parts, 0)
    if p.shape[0] == 1:
        return p[0]
    out = torch.empty(C, device=dev)
    check(lib.tcd_group_colsum(p.data_ptr(), 0, C, 1, p.shape[0], C, out.data_ptr(), C, 0, _stream()))
    return out


class LinearFn(Function):
    """y = x W^T + b  (nn.Linear).  dgrad: dx = dy W; wgrad: dW = dy^T x; db = colsum(dy)."""

    @staticmethod
    def forward(ctx, x, W, b, T):
        x = x.contiguous()
        M, K = x.shape
        N = W.shape[0]
        a, Kp = _cast(x, T)
        w, _ = _cast(W.detach(), T)
        y = torch.empty(M, N, device=x.device)
        ops.gemm(a, w, None if b is None else b.detach(), ACT_NONE, y, M=M, N=N, K=Kp if T != torch.float32 else K)
        ctx.save_for_backward(x, W)
        ctx.has_bias, ctx.T = b is not None, T
        return y

    @staticmethod
    def backward(ctx, dy):
        x, W = ctx.saved_tensors
        T = ctx.T
        dy = dy.contiguous()
        M, K = x.shape
        N = W.shape[0]
        dx = dW = db = None
        if ctx.needs_input_grad[0]:
            g, Np = _cast(dy, T)                       # (M, Np)
            wt, _ = _cast_t(W.detach(), T)             # (K, Np')
            dx = torch.empty(M, K, device=x.device)
            ops.gemm(g, wt, None, ACT_NONE, dx, M=M, N=K, K=min(g.shape[1], wt.shape[1]))
        if ctx.needs_input_grad[1]:
            gt, Mp = _cast_t(dy, T)                    # (N, Mp)
            xt, _ = _cast_t(x, T)                      # (K, Mp)
            dW = torch.empty(N, K, device=x.device)
            ops.gemm(gt, xt, None, ACT_NONE, dW, M=N, N=K, K=Mp)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            db = colsum(dy)
        return dx, dW, db, None


class ActFn(Function):
    @staticmethod
    def forward(ctx, z, act):
        z = z.contiguous()
        y = torch.empty_like(z)
        check(_lib.lib().tcd_act_forward(act, z.data_ptr(), y.data_ptr(), z.numel(), _stream()))
        ctx.save_for_backward(z)
        ctx.act = act
        return y

    @staticmethod
    def backward(ctx, dy):
        (z,) = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(z)
        check(_lib.lib().tcd_act_backward(ctx.act, z.data_ptr(), dy.data_ptr(), dx.data_ptr(), z.numel(), _stream()))
        return dx, None


class LayerNormFn(Function):
    @staticmethod
    def forward(ctx, x, gamma, beta, eps):
        x = x.contiguous()
        R, D = x.shape
        y = torch.empty_like(x)
        ops.layernorm_rotary(x, gamma.detach(), beta.detach(), eps, y, None, None, None, R, D, 1)
        ctx.save_for_backward(x, gamma)
        ctx.eps = eps
        return y

    @staticmethod
    def backward(ctx, dy):
        x, gamma = ctx.saved_tensors
        dy = dy.contiguous()
        R, D = x.shape
        lib = _lib.lib()
        P = lib.tcd_layernorm_backward_partials(R)
        dx = torch.empty_like(x)
        pg = torch.empty(P, D, device=x.device)
        pb = torch.empty(P, D, device=x.device)
        check(lib.tcd_layernorm_backward(x.data_ptr(), gamma.detach().data_ptr(), dy.data_ptr(), ctx.eps, dx.data_ptr(),
                                         pg.data_ptr(), pb.data_ptr(), R, D, _stream()))
        return dx, colsum(pg), colsum(pb), None


class RotaryFn(Function):
    """Rotation of feature pairs by position * freq (model/rotary_embedding_torch.py:39-59); backward rotates by -theta."""

    @staticmethod
    def forward(ctx, x, cos, sin, tps):
        x = x.contiguous()
        R, D = x.shape
        y = torch.empty_like(x)
        ops.rotary(x, y, cos, sin, R, D, tps)
        ctx.save_for_backward(cos, sin)
        ctx.tps = tps
        return y

    @staticmethod
    def backward(ctx, dy):
        cos, sin = ctx.saved_tensors
        dy = dy.contiguous()
        R, D = dy.shape
        dx = torch.empty_like(dy)
        ops.rotary(dy, dx, cos, (-sin).contiguous(), R, D, ctx.tps)
        return dx, None, None, None


class AttentionFn(Function):
    """softmax(scale q k^T) v per (sample, head); q (n, Lq, H*64), k, v (n, Lk, H*64) fp32 contiguous."""

    @staticmethod
    def forward(ctx, q, k, v, heads, scale):
        q, k, v = q.contiguous(), k.contiguous(), v.contiguous()
        n, Lq, HD = q.shape
        Lk = k.shape[1]
        o = torch.empty_like(q)
        ops.attention(q, HD, Lq * HD, k, HD, Lk * HD, v, HD, Lk * HD, o, HD, Lq * HD, n, heads, Lq, Lk, scale)
        ctx.save_for_backward(q, k, v, o)
        ctx.heads, ctx.scale = heads, scale
        return o

    @staticmethod
    def backward(ctx, do):
        q, k, v, o = ctx.saved_tensors
        do = do.contiguous()
        n, Lq, HD = q.shape
        Lk = k.shape[1]
        lib = _lib.lib()
        ws = torch.empty(lib.tcd_attention_backward_workspace_floats(n, ctx.heads, Lq), device=q.device)
        dq, dk, dv = torch.empty_like(q), torch.empty_like(k), torch.empty_like(v)
        check(lib.tcd_attention_backward(q.data_ptr(), HD, Lq * HD, k.data_ptr(), HD, Lk * HD, v.data_ptr(), HD, Lk * HD,
                                         o.data_ptr(), HD, Lq * HD, do.data_ptr(), HD, Lq * HD, dq.data_ptr(), HD, Lq * HD,
                                         dk.data_ptr(), HD, Lk * HD, dv.data_ptr(), HD, Lk * HD, ws.data_ptr(), n, ctx.heads,
                                         Lq, Lk, ctx.scale, _stream()))
        return dq, dk, dv, None, None


class FiLMResidualFn(Function):
    """out = x + (1 + scale) * v + shift with per-sample (scale | shift) = film[b, off:off+2D]  (model/model.py:171-173);
    film=None is the plain residual add of the music encoder."""

    @staticmethod
    def forward(ctx, x, v, film, off, L):
        x, v = x.contiguous(), v.contiguous()
        R, D = x.shape
        out = torch.empty_like(x)
        f = None if film is None else film.contiguous()
        ops.film_residual_norm(F32, x, out, v, None, 0.0, f, 0 if f is None else f.stride(0), off, None, 0.0, None, None,
                               None, None, R, D, L)
        ctx.save_for_backward(v, f) if f is not None else ctx.save_for_backward(v)
        ctx.has_film, ctx.off, ctx.L = f is not None, off, L
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        if not ctx.has_film:
            return dout, dout, None, None, None
        v, film = ctx.saved_tensors
        R, D = v.shape
        n = R // ctx.L
        dv = torch.empty_like(v)
        dfilm = torch.zeros_like(film)
        check(_lib.lib().tcd_film_backward(dout.data_ptr(), v.data_ptr(), film.data_ptr(), film.stride(0), ctx.off,
                                           dv.data_ptr(), dfilm.data_ptr(), dfilm.stride(0), ctx.off, n, ctx.L, D, _stream()))
        return dout, dv, dfilm, None, None


class LossFn(Function):
    """The four p_losses terms (model/diffusion.py:664-741): returns the 5-vector (total, recon, vel, fk, foot);
    only `total` is differentiable (that is what the reference back-propagates, TCDiff.py:232)."""

    @staticmethod
    def forward(ctx, model_out, target, p2w, B, S, dn, loss_type="l2"):
        model_out = model_out.contiguous()
        losses = ops.loss_forward(model_out, target, p2w, B, S, dn, loss_type)
        ctx.save_for_backward(model_out, target, p2w)
        ctx.dims = (B, S, dn, loss_type)
        return losses

    @staticmethod
    def backward(ctx, g):
        model_out, target, p2w = ctx.saved_tensors
        B, S, dn, loss_type = ctx.dims
        # d/d total only (the four parts are reporting-only); scaled on the device: reading g on the host would
        # stall the stream between the forward and the backward pass
        dm = ops.loss_backward(model_out, target, p2w, 1.0, B, S, dn, loss_type)
        return dm.mul_(g[0]), None, None, None, None, None, None


# ---------------------------------------------------------------------------------------------------------------------
class _Tables:
    pass


def _tables(model):
    """rotary cos/sin and the timestep-embedding table, built on the host exactly like the inference engine's
    (engine.PackedWeights) but cached independently of the weights (which change every optimizer step)."""
    dev = model.input_projection.weight.device
    tb = getattr(model._cache, "train_tables", None)       # derived state lives in the model's _Cache (dropped on deepcopy/pickle)
    if tb is None or tb.device != dev:
        D = model.latent_dim
        tb = _Tables()
        tb.device = dev
        freqs = model.rotary.freqs.detach().float().cpu()
        Lmax = max(model.seq_len * model.required_dancer_num, model.seq_len + 2)
        ang = torch.arange(Lmax).type(freqs.dtype)[:, None] * freqs[None, :]
        tb.rot_cos, tb.rot_sin = ang.cos().to(dev).contiguous(), ang.sin().to(dev).contiguous()
        half = D // 2
        e = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
        e = torch.arange(getattr(model, "time_table_rows", 1000))[:, None] * e[None, :]
        tb.time_table = torch.cat((e.sin(), e.cos()), dim=-1).to(dev).contiguous()
        model._cache.train_tables = tb
    return tb


def _lin(P, name, x, T, bias=True):
    return LinearFn.apply(x, P[name + ".weight"], P[name + ".bias"] if bias else None, T)


def _ln(P, name, x, eps=1e-5):
    return LayerNormFn.apply(x, P[name + ".weight"], P[name + ".bias"], eps)


def _forward_fp32(model, x, cond_embed, times, keep):
    """DanceDecoder.forward (model/model.py:548-624) on the fp32 autograd tape.  x (B, L, 151) fp32, keep (B,) bool."""
    if model.dropout_p > 0 and model.training:
        raise NotImplementedError("dropout > 0 is implemented on the bf16 tape only (the fp32 tape is the parity mode); "
                                  "build the model with dropout=0.0 or dtype='bf16'")
    P = dict(model.named_parameters())
    T = model.compute_dtype
    w = _tables(model)                                         # host-built rotary / timestep tables (weight independent)
    D, dn, S, H, NL = model.latent_dim, model.required_dancer_num, model.seq_len, model.num_heads, model.num_layers
    B = x.shape[0]
    L = S * dn
    Mm = S + 2
    scale = 1.0 / math.sqrt(HEAD_DIM)
    keep = keep.to(torch.bool)
    # front (model.py:560-561)
    h = _lin(P, "input_projection", x.reshape(B * L, 151), T)
    g = ActFn.apply(_lin(P, "relative_projection_layer.0", h.view(B * S, dn * D), T), ACT_RELU)
    g = ActFn.apply(_lin(P, "relative_projection_layer.2", g, T), ACT_RELU)
    xr = _lin(P, "relative_projection_layer.4", g, T).view(B * L, D)
    # music path (model.py:572-581)
    c = cond_embed[:, : 2 * S, :].reshape(B * S, -1).float().contiguous()
    c = _lin(P, "cond_projection.2", ActFn.apply(_lin(P, "cond_projection.0", c, T), ACT_RELU), T)
    for i in range(2):
        p = f"cond_encoder.{i}"
        nrm = _ln(P, p + ".norm1", c)
        qk = RotaryFn.apply(nrm, w.rot_cos, w.rot_sin, S)
        Wi, bi = P[p + ".self_attn.in_proj_weight"], P[p + ".self_attn.in_proj_bias"]
        q = LinearFn.apply(qk, Wi[:D], bi[:D], T)
        k = LinearFn.apply(qk, Wi[D:2 * D], bi[D:2 * D], T)
        v = LinearFn.apply(nrm, Wi[2 * D:], bi[2 * D:], T)
        a = AttentionFn.apply(q.view(B, S, D), k.view(B, S, D), v.view(B, S, D), H, 1.0 / math.sqrt(D // H))
        c = FiLMResidualFn.apply(c, _lin(P, p + ".self_attn.out_proj", a.view(B * S, D), T), None, 0, S)
        f = _lin(P, p + ".linear2", ActFn.apply(_lin(P, p + ".linear1", _ln(P, p + ".norm2", c), T), ACT_GELU), T)
        c = FiLMResidualFn.apply(c, f, None, 0, S)
    tokens = torch.where(keep[:, None, None], c.view(B, S, D), P["null_cond_embed"])          # model.py:589
    pooled = tokens.mean(dim=-2)                                                              # :593
    ch = _ln(P, "non_attn_cond_projection.0", pooled)
    ch = _lin(P, "non_attn_cond_projection.3", ActFn.apply(_lin(P, "non_attn_cond_projection.1", ch, T), ACT_SILU), T)
    # time path (model.py:601-612)
    te = w.time_table[times.clamp(0, w.time_table.shape[0] - 1)]
    th = ActFn.apply(_lin(P, "time_mlp.1", te, T), ACT_MISH)
    t = _lin(P, "to_time_cond.0", th, T)
    tt = _lin(P, "to_time_tokens.0", th, T).view(B, 2, D)
    t = t + torch.where(keep[:, None], ch, P["null_cond_hidden"])
    mt = ActFn.apply(t, ACT_MISH)
    mem = _ln(P, "norm_cond", torch.cat((tokens, tt), dim=-2).reshape(B * Mm, D))             # :615-616
    mem_rot = RotaryFn.apply(mem, w.rot_cos, w.rot_sin, Mm)
    for i in range(NL):
        p = f"seqTransDecoder.stack.{i}"
        film = torch.cat([_lin(P, f"{p}.film{j}.block.1", mt, T) for j in (1, 2, 3)], dim=1)  # (B, 3*2D)
        # self-attention block (model.py:326-327)
        n1 = _ln(P, p + ".norm1", xr)
        qk = RotaryFn.apply(n1, w.rot_cos, w.rot_sin, L)
        q = _lin(P, p + ".self_attn.w_qs", qk, T, bias=False)
        k = _lin(P, p + ".self_attn.w_ks", qk, T, bias=False)
        v = _lin(P, p + ".self_attn.w_vs", n1, T, bias=False)
        a = AttentionFn.apply(q.view(B, L, -1), k.view(B, L, -1), v.view(B, L, -1), H, scale)
        o = _ln(P, p + ".self_attn.layer_norm", _lin(P, p + ".self_attn.fc", a.view(B * L, -1), T, bias=False), 1e-6)
        xr = FiLMResidualFn.apply(xr, o, film, 0, L)
        # cross-attention block (model.py:331-334)
        q = _lin(P, p + ".multihead_attn.w_qs", RotaryFn.apply(_ln(P, p + ".norm2", xr), w.rot_cos, w.rot_sin, L), T, bias=False)
        k = _lin(P, p + ".multihead_attn.w_ks", mem_rot, T, bias=False)
        v = _lin(P, p + ".multihead_attn.w_vs", mem, T, bias=False)
        a = AttentionFn.apply(q.view(B, L, -1), k.view(B, Mm, -1), v.view(B, Mm, -1), H, scale)
        o = _ln(P, p + ".multihead_attn.layer_norm", _lin(P, p + ".multihead_attn.fc", a.view(B * L, -1), T, bias=False), 1e-6)
        xr = FiLMResidualFn.apply(xr, o, film, 2 * D, L)
        # feed-forward block (model.py:338-339) and the layer's return value linear3(norm4(x)) (:344)
        f = _lin(P, p + ".linear2", ActFn.apply(_lin(P, p + ".linear1", _ln(P, p + ".norm3", xr), T), ACT_GELU), T)
        xr = FiLMResidualFn.apply(xr, f, film, 4 * D, L)
        xr = _lin(P, p + ".linear3", _ln(P, p + ".norm4", xr), T)
    return _lin(P, "final_layer", xr, T).view(B, L, 151)


# =====================================================================================================================
# bf16 tape: GEMM operands and their gradients are bf16 tensors (no casts around the contractions), the residual
# stream, LayerNorm statistics, FiLM tables, losses and all parameter gradients are fp32.
#   * every nn.Linear: forward tcd_gemm (tcgen05), dgrad tcd_gemm against a per-step transposed bf16 copy of the
#     (small) weight, wgrad tcd_gemm_tn reading dY and X as stored (MN-major operands, split-K), bias grad
#     tcd_colsum_bf16.  Projections that share their input (w_qs | w_ks; the three FiLM generators) are one GEMM.
#   * attention: tcgen05 forward with log-sum-exp + tcgen05 flash backward (attention_bwd_tc.cu).
#   * LayerNorm(+rotary) forward in one kernel, backward with the rotary branch fused.
BF = torch.bfloat16


class _Pack:
    """bf16 operand copies of one (possibly concatenated) nn.Linear weight: W (sum N, Kp) for the forward, W^T (Kp, sum Np)
    for dgrad, fp32 bias.  The buffers are allocated once (pad rows / columns zero) and refreshed from the fp32 parameters
    whenever those change — all packs of the model together, in ONE tcd_pack_weights launch at the head of a step
    (refresh_packs); a pack used before its first batched refresh fills itself."""

    def __init__(self, parts, biases):
        self.parts = parts                      # [(param, row_lo, row_hi)]
        self.biases = biases                    # [(param, lo, hi)] or None
        self.sig = None
        dev = parts[0][0].device
        self.K = parts[0][0].shape[1]
        self.Kp = _up8(self.K)
        self.N = sum(hi - lo for _, lo, hi in parts)
        self.Np = _up8(self.N)
        self.w = torch.zeros(self.N, self.Kp, dtype=BF, device=dev)
        self.wt = torch.zeros(self.Kp, self.Np, dtype=BF, device=dev)
        self.b = None
        if biases is not None and len(biases) > 1:
            self.b = torch.empty(sum(hi - lo for _, lo, hi in biases), dtype=torch.float32, device=dev)

    def signature(self):
        sig = tuple((p.data_ptr(), p._version) for p, _, _ in self.parts)
        if self.biases is not None:
            sig += tuple((p.data_ptr(), p._version) for p, _, _ in self.biases)
        return sig

    def segments(self):
        """(src_ptr, w_ptr, wt_ptr, src_ld, w_ld, wt_ld, rows, cols) per part, the record layout of tcd_pack_weights."""
        out, off = [], 0
        for p, lo, hi in self.parts:
            src = p.detach()[lo:hi]
            out.append((src.data_ptr(), self.w.data_ptr() + 2 * off * self.Kp, self.wt.data_ptr() + 2 * off, src.stride(0), self.Kp,
                        self.Np, hi - lo, self.K))
            off += hi - lo
        return out

    def refresh_bias(self):
        if self.biases is None:
            return
        if len(self.biases) > 1:
            torch.cat([p.detach()[lo:hi] for p, lo, hi in self.biases], out=self.b)
        else:
            p, lo, hi = self.biases[0]
            self.b = p.detach()[lo:hi]              # a view of the fp32 parameter: always current

    def get(self):
        sig = self.signature()
        if sig != self.sig:
            _pack_launch([self])
            self.refresh_bias()
            self.sig = sig
        return self


def _pack_launch(packs, cache=None):
    """One tcd_pack_weights launch for the given packs; `cache` (a dict) keeps the device segment table between steps."""
    import struct
    segs = [sg for pk in packs for sg in pk.segments()]
    key = tuple(segs)
    ent = cache.get("table") if cache is not None else None
    if ent is None or ent[0] != key:
        raw = b"".join(struct.pack("<QQQqqqii", *sg) for sg in segs)
        table = torch.frombuffer(bytearray(raw), dtype=torch.uint8).to(packs[0].w.device)
        ent = (key, table)
        if cache is not None:
            cache["table"] = ent
    blocks = max(1, min(256, max((sg[6] + 31) // 32 * ((sg[7] + 31) // 32) for sg in segs)))
    check(_lib.lib().tcd_pack_weights(ent[1].data_ptr(), len(segs), blocks, _stream()))


def refresh_packs(model):
    """Head of a training step: every bf16 weight copy the tape will ask for is rebuilt from the fp32 parameters in one
    launch if any parameter changed since the last refresh (the optimizer step bumps the version counters)."""
    cache = model._cache.__dict__
    packs = list(cache.get("train_packs", {}).values())
    if not packs or all(pk.sig == pk.signature() for pk in packs):
        return
    _pack_launch(packs, cache.setdefault("train_pack_table", {}))
    for pk in packs:
        pk.refresh_bias()
        pk.sig = pk.signature()


def _pack(model, key, parts, biases=None):
    packs = model._cache.__dict__.setdefault("train_packs", {})
    pk = packs.get(key)
    if pk is None or any(a[0] is not b[0] for a, b in zip(pk.parts, parts)):
        pk = packs[key] = _Pack(parts, biases)
    return pk.get()


def _to_bf16(x, width=None):
    """(R, C) fp32 -> (R, up8(C) or width) bf16, zero padded."""
    R, C = x.shape
    Cp = _up8(C) if width is None else width
    out = torch.zeros(R, Cp, dtype=BF, device=x.device) if Cp != C else torch.empty(R, Cp, dtype=BF, device=x.device)
    ops.convert_pad(x, x.stride(0), out, Cp, R, C)
    return out


def _colsum16(a):
    R, C = a.shape
    lib = _lib.lib()
    out = torch.empty(C, device=a.device)
    ws = torch.empty(max(1, lib.tcd_colsum_bf16_workspace_floats(R, C)), device=a.device)
    check(lib.tcd_colsum_bf16(a.data_ptr(), a.stride(0), R, C, out.data_ptr(), ws.data_ptr(), _stream()))
    return out


class BLinearFn(Function):
    """y = x [W_1; W_2; ...]^T + [b_1; b_2; ...] with x (M, Kp) bf16; tensors = weights then (optionally) biases, each a
    Parameter or a row slice of one (their gradients are returned per tensor)."""

    @staticmethod
    def forward(ctx, x, pk, out_dtype, n_w, *tensors):
        M = x.shape[0]
        if out_dtype == BF and pk.Np != pk.N:        # bf16 outputs feed other GEMMs: keep the 16-byte pitch, zero padded
            y = torch.zeros(M, pk.Np, dtype=BF, device=x.device)
        else:
            y = torch.empty(M, pk.N, dtype=out_dtype, device=x.device)
        ops.gemm(x, pk.w, pk.b, ACT_NONE, y, M=M, N=pk.N, K=pk.Kp)
        ctx.save_for_backward(x)
        ctx.pk, ctx.n_w, ctx.n_t = pk, n_w, len(tensors)
        ctx.rows = [t.shape[0] for t in tensors[:n_w]]
        return y

    @staticmethod
    def backward(ctx, dy):
        (x,) = ctx.saved_tensors
        pk, n_w = ctx.pk, ctx.n_w
        M = x.shape[0]
        if dy.dtype != BF:                       # fp32-output layers (residual stream, FiLM tables, final_layer)
            dy = _to_bf16(dy.contiguous(), pk.Np)
        elif dy.shape[1] != pk.Np:
            dy = _to_bf16(dy.float().contiguous(), pk.Np)
        elif not dy.is_contiguous():
            dy = dy.contiguous()
        dx = None
        if ctx.needs_input_grad[0]:
            dx = torch.empty(M, pk.Kp, dtype=BF, device=x.device)
            ops.gemm(dy, pk.wt, None, ACT_NONE, dx, M=M, N=pk.Kp, K=pk.Np)
        grads = [None] * ctx.n_t
        if any(ctx.needs_input_grad[4:4 + n_w]):
            dW = ops.gemm_tn(dy, x)                                      # (Np, Kp) fp32
            off = 0
            for i, r in enumerate(ctx.rows):
                g = dW[off:off + r]
                grads[i] = g if pk.Kp == pk.K else g[:, :pk.K].contiguous()
                off += r
        if ctx.n_t > n_w and any(ctx.needs_input_grad[4 + n_w:]):
            db = _colsum16(dy)
            off = 0
            for i, r in enumerate(ctx.rows):
                grads[n_w + i] = db[off:off + r]
                off += r
        return (dx, None, None, None, *grads)


def _blin(model, P, names, x, out_dtype=BF, bias=True, rows=None):
    """names: parameter prefixes sharing the input x; rows: optional (lo, hi) row slice of a packed weight
    (nn.MultiheadAttention.in_proj_weight)."""
    if rows is None:
        Ws = [P[n + ".weight"] for n in names]
        bs = [P[n + ".bias"] for n in names] if bias else []
        parts = [(w, 0, w.shape[0]) for w in Ws]
        bparts = [(b, 0, b.shape[0]) for b in bs] if bias else None
        key = tuple(names)
    else:
        lo, hi = rows
        Wfull, bfull = P[names[0] + "_weight"], P[names[0] + "_bias"]
        Ws, bs = [Wfull[lo:hi]], [bfull[lo:hi]]
        parts, bparts = [(Wfull, lo, hi)], [(bfull, lo, hi)]
        key = (names[0], lo, hi)
    pk = _pack(model, key, parts, bparts)
    return BLinearFn.apply(x, pk, out_dtype, len(Ws), *Ws, *bs)


class BActFn(Function):
    """y = act(z) on bf16, optionally followed by a fused nn.Dropout site (p, st, site)."""

    @staticmethod
    def forward(ctx, z, act, p=0.0, st=None, site=0):
        z = z.contiguous()
        y = torch.empty_like(z)
        check(_lib.lib().tcd_act_forward_bf16(act, z.data_ptr(), y.data_ptr(), z.numel(), p, ops._ptr(st), site, _stream()))
        ctx.save_for_backward(z, st)
        ctx.act, ctx.p, ctx.site = act, p, site
        return y

    @staticmethod
    def backward(ctx, dy):
        z, st = ctx.saved_tensors
        dy = dy.contiguous()
        dx = torch.empty_like(z)
        check(_lib.lib().tcd_act_backward_bf16(ctx.act, z.data_ptr(), dy.data_ptr(), dx.data_ptr(), z.numel(), ctx.p,
                                               ops._ptr(st), ctx.site, _stream()))
        return dx, None, None, None, None


def _ln_partials_reduce(pgb, P_, D):
    """(2, P, D) per-warp partials -> (dgamma, dbeta) with ONE launch (two groups of P rows)."""
    out = torch.empty(2, D, device=pgb.device)
    check(_lib.lib().tcd_group_colsum(pgb.data_ptr(), 0, D, 2, P_, D, out.data_ptr(), D, 0, _stream()))
    return out[0], out[1]


class BLayerNormFn(Function):
    """(x_alias, plain, rotated): bf16 copies of LayerNorm(x) for x (R, D) fp32 or bf16; either copy may be skipped
    (returned as None).  With alias=True the first output is x itself, to be used for the residual connection around
    the norm: its gradient is then added to dx inside the backward kernel instead of by a separate autograd add."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, want_plain, want_rot, cos, sin, tps, alias, p=0.0, st=None, site=0):
        x = x.contiguous()
        R, D = x.shape
        yp = torch.empty(R, D, dtype=BF, device=x.device) if want_plain else None
        yr = torch.empty(R, D, dtype=BF, device=x.device) if want_rot else None
        if x.dtype == BF:                          # (p, st, site): nn.Dropout on the norm's INPUT, fused
            assert want_plain and not want_rot
            check(_lib.lib().tcd_layernorm_bf16(x.data_ptr(), gamma.detach().data_ptr(), beta.detach().data_ptr(), eps,
                                                yp.data_ptr(), R, D, p, ops._ptr(st), site, _stream()))
        else:
            assert p == 0.0
            ops.layernorm_rotary(x, gamma.detach(), beta.detach(), eps, yp, yr, cos, sin, R, D, tps)
        ctx.save_for_backward(x, gamma, cos, sin, st)
        ctx.eps, ctx.tps, ctx.p, ctx.site = eps, tps, p, site
        ctx.set_materialize_grads(False)
        return (x.view_as(x) if alias else None), yp, yr

    @staticmethod
    def backward(ctx, dres, dyp, dyr):
        x, gamma, cos, sin, st = ctx.saved_tensors
        R, D = x.shape
        lib = _lib.lib()
        if dyp is None and dyr is None:
            return (dres,) + (None,) * 12
        dyp = None if dyp is None else dyp.contiguous()
        dyr = None if dyr is None else dyr.contiguous()
        dres = None if dres is None else dres.contiguous()
        P_ = lib.tcd_layernorm_backward_mixed_partials(R)
        dx = torch.empty_like(x)
        pgb = torch.empty(2, P_, D, device=x.device)
        check(lib.tcd_layernorm_backward_mixed(ops._DT[x.dtype], _lib.BF16, x.data_ptr(), gamma.detach().data_ptr(),
                                               0 if dyp is None else dyp.data_ptr(), 0 if dyr is None else dyr.data_ptr(),
                                               cos.data_ptr(), sin.data_ptr(), ctx.tps, ctx.eps,
                                               0 if dres is None else dres.data_ptr(), dx.data_ptr(), pgb[0].data_ptr(),
                                               pgb[1].data_ptr(), R, D, ctx.p, ops._ptr(st), ctx.site, _stream()))
        dg, db = _ln_partials_reduce(pgb, P_, D)
        return dx, dg, db, None, None, None, None, None, None, None, None, None, None


# ----------------------------------------------------------------------------------------------------- dropout
# Sites (model/model.py:98,103,240-245,383,396,400-401; nn.MultiheadAttention dropout), per layer:
#   music encoder layer i: 0 attention probabilities, 1 dropout1 (after out_proj), 2 FFN inner, 3 dropout2
#   decoder layer i: 0 / 3 self / cross attention probabilities, 1 / 4 after fc, 2 / 5 dropout1 / dropout2, 6 FFN inner,
#   7 dropout3
def site_id(kind, layer, k):
    return (0 if kind == "enc" else 100) + 16 * layer + k


def dropout_state(model, advance=True):
    """Device-resident {seed, step counter} of the model's dropout masks (csrc/dropout.cuh).  Returns the snapshot this
    forward pass uses (the backward pass re-derives the masks from it) and advances the live counter; both are device
    ops, so a CUDA-graph replay draws new masks every step.  The seed comes from torch's generator (torch.manual_seed)."""
    dev = model.input_projection.weight.device
    st = getattr(model._cache, "dropout_rng", None)
    if st is None or st.device != dev:
        st = torch.zeros(2, dtype=torch.int64, device=dev)
        seed = int(torch.randint(0, 2 ** 62, (1,)))
        if torch.distributed.is_available() and torch.distributed.is_initialized():
            seed ^= (torch.distributed.get_rank() + 1) * 0x9E3779B97F4A7C15 & (2 ** 62 - 1)    # decorrelate the ranks' masks
        st[0] = seed
        model._cache.dropout_rng = st
    snap = st.clone()
    if advance:
        st[1:].add_(1)
    return snap


class BAttentionFn(Function):
    """softmax(scale q k^T) v per (sample, head) on the tcgen05 kernels.  layout "qk|v": a = packed (n, L, 2*H*64)
    projections [q | k], b = v; layout "q|k|v": a, b, c separate."""

    @staticmethod
    def forward(ctx, a, b, c, heads, scale, p=0.0, st=None, site=0):
        HD = heads * HEAD_DIM
        if c is None:
            q, k, v = a[..., :HD], a[..., HD:], b
        else:
            q, k, v = a, b, c
        o, lse = ops.attention_train_forward(q, k, v, heads, scale, p, st, site)
        ctx.save_for_backward(a, b, c, o, lse, st)
        ctx.heads, ctx.scale, ctx.p, ctx.site = heads, scale, p, site
        return o

    @staticmethod
    def backward(ctx, do):
        a, b, c, o, lse, st = ctx.saved_tensors
        heads, HD = ctx.heads, ctx.heads * HEAD_DIM
        do = do.contiguous()
        kw = dict(dropout_p=ctx.p, rng_state=st, site=ctx.site)
        if c is None:
            da = torch.empty_like(a)
            dv = torch.empty_like(b)
            ops.attention_train_backward(a[..., :HD], a[..., HD:], b, o, do, lse, heads, ctx.scale, dq=da[..., :HD],
                                         dk=da[..., HD:], dv=dv, **kw)
            return da, dv, None, None, None, None, None, None
        dq, dk, dv = ops.attention_train_backward(a, b, c, o, do, lse, heads, ctx.scale, **kw)
        return dq, dk, dv, None, None, None, None, None


class BFiLMResidualFn(Function):
    """out = x + (1 + scale) v + shift, x / out fp32, v bf16 (film=None: out = x + v), optionally with the nn.Dropout
    site that precedes the block's residual connection fused on v (p, st, site)."""

    @staticmethod
    def forward(ctx, x, v, film, off, L, p=0.0, st=None, site=0):
        x, v = x.contiguous(), v.contiguous()
        R, D = x.shape
        out = torch.empty_like(x)
        f = None if film is None else film.contiguous()
        check(_lib.lib().tcd_film_residual_bf16(x.data_ptr(), v.data_ptr(), ops._ptr(f), 0 if f is None else f.stride(0), off,
                                                out.data_ptr(), R, L, D, p, ops._ptr(st), site, _stream()))
        ctx.save_for_backward(v, f, st)
        ctx.off, ctx.L, ctx.p, ctx.site = off, L, p, site
        return out

    @staticmethod
    def backward(ctx, dout):
        dout = dout.contiguous()
        v, film, st = ctx.saved_tensors
        if film is None:
            dv = _to_bf16(dout)
            if ctx.p > 0:
                dv = ops.dropout(dv, ctx.p, st, ctx.site, out=dv)
            return dout, dv, None, None, None, None, None, None
        R, D = v.shape
        n = R // ctx.L
        lib = _lib.lib()
        dv = torch.empty_like(v)
        dfilm = torch.zeros_like(film)
        ws = torch.empty(lib.tcd_film_backward_workspace_floats(n, ctx.L, D), device=v.device)
        check(lib.tcd_film_backward_bf16(dout.data_ptr(), v.data_ptr(), film.data_ptr(), film.stride(0), ctx.off, dv.data_ptr(),
                                         dfilm.data_ptr(), dfilm.stride(0), ctx.off, ws.data_ptr(), n, ctx.L, D, ctx.p,
                                         ops._ptr(st), ctx.site, _stream()))
        return dout, dv, dfilm, None, None, None, None, None


def _bln(P, name, x, w, eps=1e-5, plain=True, rot=False, tps=1, alias=False, drop=(0.0, None, 0)):
    """-> (x alias or None, plain or None, rotated or None); drop = (p, state, site) of a dropout fused on a bf16 input"""
    return BLayerNormFn.apply(x, P[name + ".weight"], P[name + ".bias"], eps, plain, rot, w.rot_cos, w.rot_sin, tps, alias,
                              *drop)


def _forward_bf16(model, x, cond_embed, times, keep):
    """DanceDecoder.forward (model/model.py:548-624) on the bf16 autograd tape."""
    P = dict(model.named_parameters())
    w = _tables(model)
    D, dn, S, H, NL = model.latent_dim, model.required_dancer_num, model.seq_len, model.num_heads, model.num_layers
    B = x.shape[0]
    L = S * dn
    Mm = S + 2
    scale = 1.0 / math.sqrt(HEAD_DIM)
    keep = keep.to(torch.bool)
    F32_ = torch.float32
    pdrop = float(model.dropout_p) if model.training else 0.0
    dr = (pdrop, dropout_state(model)) if pdrop > 0 else None       # (p, {seed, counter} snapshot of this pass)
    ap = (pdrop, dr[1]) if dr else (0.0, None)

    def ds(kind, layer, k):                                         # (p, state, site) of a fused dropout site
        return (pdrop, dr[1], site_id(kind, layer, k)) if dr else (0.0, None, 0)
    # front (model.py:560-561): input projection + fusion MLP over the dancers of a frame
    h = _blin(model, P, ["input_projection"], _to_bf16(x.reshape(B * L, 151)))
    g = BActFn.apply(_blin(model, P, ["relative_projection_layer.0"], h.view(B * S, dn * D)), ACT_RELU)
    g = BActFn.apply(_blin(model, P, ["relative_projection_layer.2"], g), ACT_RELU)
    xr = _blin(model, P, ["relative_projection_layer.4"], g, out_dtype=F32_).view(B * L, D)
    # music path (model.py:572-581)
    c = _to_bf16(cond_embed[:, : 2 * S, :].reshape(B * S, -1).float().contiguous())
    c = _blin(model, P, ["cond_projection.2"], BActFn.apply(_blin(model, P, ["cond_projection.0"], c), ACT_RELU), out_dtype=F32_)
    for i in range(2):
        p = f"cond_encoder.{i}"
        c, nrm, qk = _bln(P, p + ".norm1", c, w, plain=True, rot=True, tps=S, alias=True)
        qkp = _blin(model, P, [p + ".self_attn.in_proj"], qk, rows=(0, 2 * D))
        v = _blin(model, P, [p + ".self_attn.in_proj"], nrm, rows=(2 * D, 3 * D))
        a = BAttentionFn.apply(qkp.view(B, S, 2 * D), v.view(B, S, D), None, H, 1.0 / math.sqrt(D // H), *ap,
                               site_id("enc", i, 0))
        c = BFiLMResidualFn.apply(c, _blin(model, P, [p + ".self_attn.out_proj"], a.view(B * S, D)), None, 0, S,
                                  *ds("enc", i, 1))
        c, n2, _ = _bln(P, p + ".norm2", c, w, alias=True)
        f = BActFn.apply(_blin(model, P, [p + ".linear1"], n2), ACT_GELU, *ds("enc", i, 2))
        c = BFiLMResidualFn.apply(c, _blin(model, P, [p + ".linear2"], f), None, 0, S, *ds("enc", i, 3))
    tokens = torch.where(keep[:, None, None], c.view(B, S, D), P["null_cond_embed"])          # model.py:589
    pooled = tokens.mean(dim=-2)                                                              # :593
    _, ch, _ = _bln(P, "non_attn_cond_projection.0", pooled, w)
    ch = _blin(model, P, ["non_attn_cond_projection.3"],
               BActFn.apply(_blin(model, P, ["non_attn_cond_projection.1"], ch), ACT_SILU), out_dtype=F32_)
    # time path (model.py:601-612)
    te = w.time_table[times.clamp(0, w.time_table.shape[0] - 1)]
    th = BActFn.apply(_blin(model, P, ["time_mlp.1"], te.to(BF)), ACT_MISH)
    tt2 = _blin(model, P, ["to_time_cond.0", "to_time_tokens.0"], th, out_dtype=F32_)          # (B, D + 2D) in one GEMM
    t = tt2[:, :D] + torch.where(keep[:, None], ch, P["null_cond_hidden"])
    tt = tt2[:, D:].reshape(B, 2, D)
    mt = ActFn.apply(t, ACT_MISH).to(BF)
    _, mem, mem_rot = _bln(P, "norm_cond", torch.cat((tokens, tt), dim=-2).reshape(B * Mm, D), w, plain=True, rot=True, tps=Mm)
    for i in range(NL):
        p = f"seqTransDecoder.stack.{i}"
        film = _blin(model, P, [f"{p}.film{j}.block.1" for j in (1, 2, 3)], mt, out_dtype=F32_)   # (B, 3*2D), one GEMM
        # self-attention block (model.py:326-327)
        xr, n1, qk = _bln(P, p + ".norm1", xr, w, plain=True, rot=True, tps=L, alias=True)
        qkp = _blin(model, P, [p + ".self_attn.w_qs", p + ".self_attn.w_ks"], qk, bias=False)
        v = _blin(model, P, [p + ".self_attn.w_vs"], n1, bias=False)
        a = BAttentionFn.apply(qkp.view(B, L, 2 * D), v.view(B, L, D), None, H, scale, *ap, site_id("dec", i, 0))
        fo = _blin(model, P, [p + ".self_attn.fc"], a.view(B * L, D), bias=False)
        _, o, _ = _bln(P, p + ".self_attn.layer_norm", fo, w, eps=1e-6, drop=ds("dec", i, 1))
        xr = BFiLMResidualFn.apply(xr, o, film, 0, L, *ds("dec", i, 2))
        # cross-attention block (model.py:331-334)
        xr, _, n2r = _bln(P, p + ".norm2", xr, w, plain=False, rot=True, tps=L, alias=True)
        q = _blin(model, P, [p + ".multihead_attn.w_qs"], n2r, bias=False)
        k = _blin(model, P, [p + ".multihead_attn.w_ks"], mem_rot, bias=False)
        v = _blin(model, P, [p + ".multihead_attn.w_vs"], mem, bias=False)
        a = BAttentionFn.apply(q.view(B, L, D), k.view(B, Mm, D), v.view(B, Mm, D), H, scale, *ap, site_id("dec", i, 3))
        fo = _blin(model, P, [p + ".multihead_attn.fc"], a.view(B * L, D), bias=False)
        _, o, _ = _bln(P, p + ".multihead_attn.layer_norm", fo, w, eps=1e-6, drop=ds("dec", i, 4))
        xr = BFiLMResidualFn.apply(xr, o, film, 2 * D, L, *ds("dec", i, 5))
        # feed-forward block (model.py:338-339) and the layer's return value linear3(norm4(x)) (:344)
        xr, n3, _ = _bln(P, p + ".norm3", xr, w, alias=True)
        f = BActFn.apply(_blin(model, P, [p + ".linear1"], n3), ACT_GELU, *ds("dec", i, 6))
        xr = BFiLMResidualFn.apply(xr, _blin(model, P, [p + ".linear2"], f), film, 4 * D, L, *ds("dec", i, 7))
        _, n4, _ = _bln(P, p + ".norm4", xr, w)
        xr = _blin(model, P, [p + ".linear3"], n4, out_dtype=BF if i == NL - 1 else F32_)
    return _blin(model, P, ["final_layer"], xr, out_dtype=F32_).view(B, L, 151)


def denoiser_forward_train(model, x, cond_embed, times, keep):
    """DanceDecoder.forward (model/model.py:548-624) on the autograd tape.  x (B, L, 151) fp32, keep (B,) bool."""
    if model.compute_dtype == torch.bfloat16:
        refresh_packs(model)
        return _forward_bf16(model, x, cond_embed, times, keep)
    return _forward_fp32(model, x, cond_embed, times, keep)


def p_losses_train(diffusion, x_start, cond, t, noise=None, keep_mask=None):
    """GaussianDiffusion.p_losses (model/diffusion.py:636-741) with gradients: (total, (recon, vel, fk, foot))."""
    model = diffusion.model
    dev = diffusion.betas.device
    B, dn, S, C = x_start.shape
    xs = x_start.to(device=dev, dtype=torch.float32).contiguous()
    noise = torch.randn(B, S, dn, C, device=dev) if noise is None else noise.to(dev).float().contiguous()
    t = t.to(dev).long().contiguous()
    x_noisy = torch.empty(B, S, dn, C, device=dev)
    target = torch.empty(B, S, dn, C, device=dev)
    ops.q_sample(xs, noise, t, diffusion.sqrt_alphas_cumprod, diffusion.sqrt_one_minus_alphas_cumprod, x_noisy, target,
                 None, 0, B, dn, S, True, True)
    if keep_mask is None:
        keep_mask = torch.zeros(B, device=dev).float().uniform_(0, 1) < (1 - diffusion.cond_drop_prob)
    out = denoiser_forward_train(model, x_noisy.view(B, S * dn, C), cond.to(dev), t, keep_mask.to(dev))
    p2w = diffusion.p2_loss_weight.gather(-1, t).contiguous()
    if diffusion.predict_epsilon:                              # the target is the noise (model/diffusion.py:657-658)
        target = noise
    losses = LossFn.apply(out.reshape(B, S, dn, C), target, p2w, B, S, dn, diffusion.loss_type)
    return losses[0], (losses[1], losses[2], losses[3], losses[4])


class GraphedTrainStep:
    """The body of the reference's training loop (TCDiff.py:227-245: zero_grad, loss, backward, Adan step, EMA) captured
    in CUDA graphs and replayed, because eagerly the step is bound by ~1000 kernel launches of host work, not by the GPU.

        step = GraphedTrainStep(diffusion, optimizer, x_first, cond_first)     # runs `warmup` real steps, then captures
        total, (recon, vel, fk, foot) = step(x, cond)                          # device tensors, overwritten every call

    Single GPU: one graph.  Data parallel: [zero_grad, loss, backward] | NCCL all-reduce of the gradient arena (eager)
    | [Adan + EMA].  Batch shapes, lr and the other hyper-parameters are frozen at capture time (the step counter and
    bias corrections live on the device).  Timesteps, noise and the classifier-free keep mask are drawn inside the
    graph from torch's CUDA generator, as in the reference's `GaussianDiffusion.loss`.
    """

    def __init__(self, diffusion, optimizer, x, cond, warmup=3):
        self.diffusion, self.opt = diffusion, optimizer
        dev = diffusion.betas.device
        self.x = x.to(device=dev, dtype=torch.float32).clone()
        self.cond = cond.to(device=dev).clone()
        optimizer.set_capturable(True)
        self.world = optimizer._world()
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream()
        side.wait_stream(cur)
        # The eager warm-up below (arenas, packs, kernel attributes, allocator pools) runs real optimisation steps.  They
        # must not count: parameters, EMA copy, optimizer state, step counters and the dropout counter are snapshotted
        # here and restored after it, so the first replay is step 1 on the caller's weights, like the reference's loop.
        model, master = diffusion.model, diffusion.master_model
        snap_p = [p.detach().clone() for p in model.parameters()]
        snap_ma = [p.detach().clone() for p in master.parameters()]
        snap_state = {id(p): {k: (v.detach().clone() if torch.is_tensor(v) else v) for k, v in optimizer.state[p].items()}
                      for g in optimizer.param_groups for p in g["params"] if len(optimizer.state[p])}
        snap_steps = {gi: f["step"] for gi, f in optimizer._flat.items()}
        drng = getattr(model._cache, "dropout_rng", None)
        snap_drng = None if drng is None else drng.clone()
        with torch.cuda.stream(side):
            for _ in range(max(1, warmup)):
                optimizer.zero_grad()
                total, _ = diffusion.loss(self.x, self.cond)
                total.backward()
                optimizer.step()
                del total
        cur.wait_stream(side)
        torch.cuda.synchronize()
        with torch.no_grad():
            torch._foreach_copy_([p.data for p in model.parameters()], snap_p)        # in place: the arena views stay
            torch._foreach_copy_([p.data for p in master.parameters()], snap_ma)
            for gi, f in optimizer._flat.items():
                step0 = snap_steps.get(gi, 0)
                ar = f["arena"]
                for p, pg, m_, v_, n_ in zip(f["live"], ar.views(f["PG"]), ar.views(f["M"]), ar.views(f["V"]), ar.views(f["N"])):
                    st0 = snap_state.get(id(p))
                    if st0 is None:
                        pg.zero_(); m_.zero_(); v_.zero_(); n_.zero_()
                    else:
                        pg.copy_(st0["prev_grad"]); m_.copy_(st0["m"]); v_.copy_(st0["v"]); n_.copy_(st0["n"])
                        step0 = int(st0["step"])
                    optimizer.state[p]["step"] = step0
                f["step"] = step0
                if "step_dev" in f:
                    f["step_dev"].fill_(step0)
            drng = getattr(model._cache, "dropout_rng", None)
            if drng is not None:
                if snap_drng is not None:
                    drng.copy_(snap_drng)
                else:
                    drng[1:].zero_()
        from .adan import _bump
        _bump(list(model.parameters()))
        _bump(list(master.parameters()))
        torch.cuda.synchronize()
        for pk in diffusion.model._cache.__dict__.get("train_packs", {}).values():
            pk.sig = None                                                  # force the weight re-packing launch into the graph
        reducers = [f["reducer"] for f in optimizer._flat.values() if "reducer" in f]
        for r in reducers:
            r.enabled = False
        l0 = _lib.LAUNCHES[0]
        self.g1 = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g1):
            optimizer.zero_grad()
            total, parts = diffusion.loss(self.x, self.cond)
            total.backward()
            if self.world == 1:
                optimizer.step()
        self.g2 = None
        if self.world > 1:
            self.g2 = torch.cuda.CUDAGraph()
            with torch.cuda.graph(self.g2):
                optimizer.step(reduce=False)
            # the capture ran the bookkeeping once without executing anything: undo it
        for f in optimizer._flat.values():
            f["step"] -= 1
        # the capture recorded the re-packing launch without executing it: an eager p_losses between construction and the
        # first replay must re-pack, not trust a signature recorded during capture
        for pk in diffusion.model._cache.__dict__.get("train_packs", {}).values():
            pk.sig = None
        self.launches = _lib.LAUNCHES[0] - l0        # C-ABI kernel-launching calls recorded in the graphs
        self.total = total.detach()
        self.parts = tuple(p.detach() for p in parts)

    def __call__(self, x, cond):
        self.x.copy_(x, non_blocking=True)
        self.cond.copy_(cond, non_blocking=True)
        self.g1.replay()
        if self.g2 is not None:
            self.opt.reduce_gradients(hooks_ran=False)
            self.g2.replay()
        self.opt.after_replay()
        _lib.LAUNCHES[0] += self.launches
        return self.total, self.parts
