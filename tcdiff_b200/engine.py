"""Host-side launch sequence of the DanceDecoder denoiser over the C-ABI kernels.

The reference evaluates model/model.py:548-624 as ~1500 eager aten ops per pass; here one pass is
~140 kernel launches of hand-written sm_100a kernels (captured into a CUDA graph by the sampler):

    front      input_projection + fusion projection (3 GEMMs)                 model.py:560-561
    music      pairing, cond_projection, 2 encoder layers, pool, cond_hidden  model.py:572-597
    time       sinusoidal table gather, time_mlp, to_time_cond/tokens, FiLM   model.py:601-612,154-168
    memory     norm_cond(cat(tokens, t_tokens)) + rotary, cross-attn K/V      model.py:615-616,386-390
    layers     8 x (self-attn, cross-attn, FFN, linear3(norm4))               model.py:308-344
    head       final_layer                                                    model.py:623

`Denoiser` is stateless w.r.t. activations (buffers come from a `Workspace`) so the sampler can run the
step-invariant parts once and replay only `front` + `layers` per step.
"""
import math

import torch

from . import ops
from ._lib import F32, BF16, ACT_NONE, ACT_RELU, ACT_GELU, ACT_MISH, ACT_SILU

HEAD_DIM = 64  # model/model.py:55,532

# The feed-forward tail's updated residual is dead: the layer returns linear3(norm4(x)) (model/model.py:344,371),
# so only norm4(x) is consumed and the 4 B/element x write is skipped (x_out = NULL in the C-ABI; r01: bit-identical
# samples, 113.6 -> 115.3 clips/s).
SKIP_DEAD_X = True


def fuse_tails():
    """Which `fc` / `linear2` GEMM + FiLM + residual + LayerNorm tails run as ONE fused kernel (csrc/gemm_frn.cu; bf16 mode,
    D = 512): bit 0 self-attention tail, bit 1 cross-attention tail, bit 2 feed-forward tail.  A compile-time choice of the
    library (csrc/tuning.cuh, TCD_TUNE_FUSE_TAILS) read back through tcd_tuning(); nothing is switchable at run time."""
    return ops._lib.lib().tcd_tuning(b"fuse_tails")


def fold_ln():
    """Whether norm3 / norm4's affine is folded into linear1 / linear3's weights where their tails run fused (csrc/tuning.cuh,
    TCD_TUNE_FOLD_LN; a compile-time choice of the library)."""
    return max(0, ops._lib.lib().tcd_tuning(b"fold_ln"))


def _round_up(v, m):
    return (v + m - 1) // m * m


class Workspace:
    """Named device buffers, allocated once and reused (static addresses for CUDA-graph replay)."""

    def __init__(self, device):
        self.device = device
        self.bufs = {}

    def get(self, name, shape, dtype, zero=False):
        key = (name, tuple(shape), dtype)
        t = self.bufs.get(key)
        if t is None:
            t = (torch.zeros if zero else torch.empty)(shape, dtype=dtype, device=self.device)
            self.bufs[key] = t
        return t

    def nbytes(self):
        return sum(t.numel() * t.element_size() for t in self.bufs.values() if torch.is_tensor(t))


class PackedWeights:
    """Kernel-side copies of a DanceDecoder state: operand dtype (fp32 | bf16), K padded to a multiple
    of 8 elements (TMA needs 16-byte pitches), projections that share an input concatenated, biases and
    LayerNorm affines in fp32, host-built rotary and timestep tables (derived cache; see model.py)."""

    def __init__(self, sd, cfg, dtype, device, n_timestep=1000):
        self.cfg = cfg
        self.dtype = dtype
        self.device = device
        D, NL = cfg["latent_dim"], cfg["num_layers"]
        f32 = lambda k: sd[k].detach().to(device=device, dtype=torch.float32).contiguous()

        def W(t):
            t = t.detach().to(device=device, dtype=torch.float32)
            kp = _round_up(t.shape[1], 8)
            if kp != t.shape[1]:
                t = torch.nn.functional.pad(t, (0, kp - t.shape[1]))
            return t.to(dtype).contiguous()

        w = lambda k: W(sd[k + ".weight"])
        b = lambda k: f32(k + ".bias")
        ln = lambda k: (f32(k + ".weight"), f32(k + ".bias"))
        self.in_w, self.in_b = w("input_projection"), b("input_projection")
        self.fus = [(w(f"relative_projection_layer.{i}"), b(f"relative_projection_layer.{i}")) for i in (0, 2, 4)]
        self.cp0 = (w("cond_projection.0"), b("cond_projection.0"))
        self.cp2 = (w("cond_projection.2"), b("cond_projection.2"))
        self.enc = []
        for i in range(2):
            p = f"cond_encoder.{i}"
            wi, bi = sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"].detach().float().to(device)
            self.enc.append(dict(
                qk_w=W(wi[: 2 * D]), qk_b=bi[: 2 * D].contiguous(), v_w=W(wi[2 * D:]), v_b=bi[2 * D:].contiguous(),
                out=(w(p + ".self_attn.out_proj"), b(p + ".self_attn.out_proj")),
                l1=(w(p + ".linear1"), b(p + ".linear1")), l2=(w(p + ".linear2"), b(p + ".linear2")),
                n1=ln(p + ".norm1"), n2=ln(p + ".norm2")))
        self.null_embed = f32("null_cond_embed").reshape(-1, D).contiguous()
        self.null_hidden = f32("null_cond_hidden").reshape(D).contiguous()
        self.nacp_ln = ln("non_attn_cond_projection.0")
        self.nacp1 = (w("non_attn_cond_projection.1"), b("non_attn_cond_projection.1"))
        self.nacp3 = (w("non_attn_cond_projection.3"), b("non_attn_cond_projection.3"))
        self.tm1 = (w("time_mlp.1"), b("time_mlp.1"))
        self.ttc = (w("to_time_cond.0"), b("to_time_cond.0"))
        self.ttt = (w("to_time_tokens.0"), b("to_time_tokens.0"))
        self.norm_cond = ln("norm_cond")
        self.layers = []
        film_w, film_b, ck, cv = [], [], [], []
        for i in range(NL):
            p = f"seqTransDecoder.stack.{i}"
            sa, ca = p + ".self_attn", p + ".multihead_attn"
            self.layers.append(dict(
                sa_qk=W(torch.cat([sd[sa + ".w_qs.weight"], sd[sa + ".w_ks.weight"]], 0)),
                sa_v=w(sa + ".w_vs"), sa_fc=w(sa + ".fc"), sa_ln=ln(sa + ".layer_norm"),
                ca_q=w(ca + ".w_qs"), ca_fc=w(ca + ".fc"), ca_ln=ln(ca + ".layer_norm"),
                l1=(w(p + ".linear1"), b(p + ".linear1")), l2=(w(p + ".linear2"), b(p + ".linear2")),
                l3=(w(p + ".linear3"), b(p + ".linear3")),
                n1=ln(p + ".norm1"), n2=ln(p + ".norm2"), n3=ln(p + ".norm3"), n4=ln(p + ".norm4")))
            if dtype == torch.bfloat16:
                # LayerNorm affine folded downstream (exact in real arithmetic): W (g * n + b) + c = (W diag(g)) n + (W b + c)
                Ly = self.layers[-1]
                for name, nrm in (("l1", "n3"), ("l3", "n4")):
                    Wf = sd[f"{p}.linear{name[1]}.weight"].detach().to(device=device, dtype=torch.float32)
                    g_, b_ = Ly[nrm]
                    Ly[name + "n"] = (W(Wf * g_[None, :]), (Ly[name][1] + Wf @ b_).contiguous())
            ck.append(sd[ca + ".w_ks.weight"])
            cv.append(sd[ca + ".w_vs.weight"])
            for f in ("film1", "film2", "film3"):
                film_w.append(sd[f"{p}.{f}.block.1.weight"])
                film_b.append(sd[f"{p}.{f}.block.1.bias"])
        self.ca_k_all = W(torch.cat(ck, 0))                      # (NL*H*64, D): layer i at rows i*D..
        self.ca_v_all = W(torch.cat(cv, 0))
        self.film_w = W(torch.cat(film_w, 0))                    # (NL*3*2D, D)
        self.film_b = torch.cat(film_b, 0).detach().float().to(device).contiguous()
        self.fin = (w("final_layer"), b("final_layer"))
        # rotary angle tables exactly as the reference builds them (fp32 positions x fp32 freqs, then cos/sin
        # on the host; model/rotary_embedding_torch.py:115-130)
        freqs = sd["rotary.freqs"].detach().float().cpu()
        Lmax = max(cfg["seq_len"] * cfg["dancers"], cfg["seq_len"] + 2)
        ang = torch.arange(Lmax).type(freqs.dtype)[:, None] * freqs[None, :]
        self.rot_cos, self.rot_sin = ang.cos().to(device).contiguous(), ang.sin().to(device).contiguous()
        # angle-major copies for the fused GEMM + tail kernel (a warp's 32 rows read them coalesced)
        self.rot_cos_t, self.rot_sin_t = self.rot_cos.t().contiguous(), self.rot_sin.t().contiguous()
        # timestep embedding table (model/utils.py:41-48), fp32 on the host for every integer timestep
        half = D // 2
        e = torch.exp(torch.arange(half) * -(math.log(10000) / (half - 1)))
        e = torch.arange(n_timestep)[:, None] * e[None, :]
        self.time_table = torch.cat((e.sin(), e.cos()), dim=-1).to(device).contiguous()


class Denoiser:
    def __init__(self, packed: PackedWeights):
        self.w = packed
        self.cfg = packed.cfg
        self.T = packed.dtype                                     # operand dtype
        self.tcd = F32 if self.T == torch.float32 else BF16

    # ---------------------------------------------------------------- small helpers
    def _lin(self, ws, name, a, wb, act, out_dtype, M, K=None, out=None, N=None):
        w_, b_ = wb
        N = w_.shape[0] if N is None else N
        if out is None:
            out = ws.get(name, (M, N), out_dtype)
        K = (a.shape[1] if K is None else K)
        ops.gemm(a, w_, b_, act, out, M=M, N=N, K=K)
        return out

    # ---------------------------------------------------------------- music (step-invariant)
    def music_encode(self, ws, cond_embed, keep, tag="m"):
        """model.py:572-597.  cond_embed (n, 2S or 2S+1, Fm) float; keep (n,) uint8.
        Returns tokens (n,S,D) fp32 (null-substituted where keep==0) and cond_hidden (n,D) fp32."""
        w, cfg, T = self.w, self.cfg, self.T
        n, Tm, Fm = cond_embed.shape
        S, D, H = cfg["seq_len"], cfg["latent_dim"], cfg["num_heads"]
        if Tm // 2 != S:
            raise ValueError(f"music length {Tm} does not match seq_len {S} (model/model.py:573-576)")
        R = n * S
        c = cond_embed[:, : 2 * S, :].reshape(R, 2 * Fm)
        if c.dtype != torch.float32 or not c.is_contiguous():
            c = c.float().contiguous()                           # .float() of model.py:578
        K0 = w.cp0[0].shape[1]                                   # 2Fm rounded up to 8
        a0 = ws.get(tag + "a0", (R, K0), T, zero=True)
        ops.convert_pad(c, 2 * Fm, a0, K0, R, 2 * Fm)
        K1 = w.cp2[0].shape[1]
        h0 = ws.get(tag + "h0", (R, K1), T, zero=True)            # pad columns stay zero
        ops.gemm(a0, w.cp0[0], w.cp0[1], ACT_RELU, h0, M=R, N=Fm, K=K0)
        tok = ws.get(tag + "tok", (R, D), torch.float32)
        ops.gemm(h0, w.cp2[0], w.cp2[1], ACT_NONE, tok, M=R, N=D, K=K1)
        plain = ws.get(tag + "plain", (R, D), T)
        rot = ws.get(tag + "rot", (R, D), T)
        qk = ws.get(tag + "qk", (R, 2 * D), T)
        v = ws.get(tag + "v", (R, D), T)
        ctx = ws.get(tag + "ctx", (R, D), T)
        o = ws.get(tag + "o", (R, D), torch.float32)
        ff = ws.get(tag + "ff", (R, cfg["ff_size"]), T)
        hd = D // H
        assert hd == HEAD_DIM, "music encoder head dim must be 64"
        for i, L in enumerate(w.enc):
            if i == 0:
                ops.layernorm_rotary(tok, L["n1"][0], L["n1"][1], 1e-5, plain, rot, w.rot_cos, w.rot_sin, R, D, S)
            ops.gemm(rot, L["qk_w"], L["qk_b"], ACT_NONE, qk, M=R)
            ops.gemm(plain, L["v_w"], L["v_b"], ACT_NONE, v, M=R)
            ops.attention(qk, 2 * D, S * 2 * D, qk, 2 * D, S * 2 * D, v, D, S * D, ctx, D, S * D, n, H, S, S,
                          1.0 / math.sqrt(hd), k_off=D)
            ops.gemm(ctx, L["out"][0], L["out"][1], ACT_NONE, o, M=R)
            ops.film_residual_norm(self.tcd, tok, tok, o, None, 0.0, None, 0, 0, L["n2"], 1e-5, plain, None, None, None,
                                   R, D, S)
            ops.gemm(plain, L["l1"][0], L["l1"][1], ACT_GELU, ff, M=R)
            ops.gemm(ff, L["l2"][0], L["l2"][1], ACT_NONE, o, M=R)
            nxt = w.enc[i + 1]["n1"] if i + 1 < len(w.enc) else None
            ops.film_residual_norm(self.tcd, tok, tok, o, None, 0.0, None, 0, 0, nxt, 1e-5,
                                   plain if nxt is not None else None, rot if nxt is not None else None,
                                   w.rot_cos, w.rot_sin, R, D, S)
        pooled = ws.get(tag + "pooled", (n, D), T)
        ops.cond_pool(tok, w.null_embed, keep, w.nacp_ln[0], w.nacp_ln[1], pooled, n, S, D)
        c1 = self._lin(ws, tag + "c1", pooled, w.nacp1, ACT_SILU, T, n)
        ch = self._lin(ws, tag + "ch", c1, w.nacp3, ACT_NONE, torch.float32, n)
        return tok.view(n, S, D), ch

    # ---------------------------------------------------------------- time
    def time_path(self, ws, times, tag="t"):
        """model.py:601-605: times (n,) int64 -> t_lin (n,D) fp32, t_tokens (n,2,D) fp32."""
        w, D, T = self.w, self.cfg["latent_dim"], self.T
        n = times.shape[0]
        te = ws.get(tag + "te", (n, D), T)
        ops.time_embed(times, w.time_table, te, n, D)
        th = self._lin(ws, tag + "th", te, w.tm1, ACT_MISH, T, n)
        t_lin = self._lin(ws, tag + "tl", th, w.ttc, ACT_NONE, torch.float32, n)
        tt = self._lin(ws, tag + "tt", th, w.ttt, ACT_NONE, torch.float32, n)
        return t_lin, tt.view(n, 2, D)

    def film_table(self, ws, t_lin, cond_hidden, keep, tag="f"):
        """model.py:609-612 + every DenseFiLM (model.py:164-168): (n, NL*3*2D) fp32, block (layer i, film j)
        at columns (3i+j)*2D holding [scale | shift]."""
        w, D, T = self.w, self.cfg["latent_dim"], self.T
        n = t_lin.shape[0]
        mt = ws.get(tag + "mish", (n, D), T)
        ops.time_cond(t_lin, cond_hidden, w.null_hidden, keep, None, mt, n, D)
        film = ws.get(tag + "film", (n, w.film_w.shape[0]), torch.float32)
        ops.gemm(mt, w.film_w, w.film_b, ACT_NONE, film, M=n)
        return film

    # ---------------------------------------------------------------- cross-attention memory
    def memory_kv(self, ws, tokens, t_tokens, tag="kv"):
        """model.py:615-616 and the K/V projections of every layer's cross-attention (model.py:79-80,388):
        returns Kc, Vc (n, S+2, NL*D) in operand dtype; layer i occupies columns i*D..(i+1)*D."""
        w, cfg, T = self.w, self.cfg, self.T
        n, S, D = tokens.shape
        Mm = S + 2
        mp = ws.get(tag + "mp", (n * Mm, D), T)
        mr = ws.get(tag + "mr", (n * Mm, D), T)
        ops.build_memory(tokens, t_tokens, w.norm_cond[0], w.norm_cond[1], mp, mr, w.rot_cos, w.rot_sin, n, S, D)
        NLD = w.ca_k_all.shape[0]
        Kc = ws.get(tag + "K", (n * Mm, NLD), T)
        Vc = ws.get(tag + "V", (n * Mm, NLD), T)
        ops.gemm(mr, w.ca_k_all, None, ACT_NONE, Kc, M=n * Mm)
        ops.gemm(mp, w.ca_v_all, None, ACT_NONE, Vc, M=n * Mm)
        return Kc.view(n, Mm, NLD), Vc.view(n, Mm, NLD)

    # ---------------------------------------------------------------- front
    def front(self, ws, x, n, xres, xpad=None, tag="fr"):
        """model.py:553-561: x (n, L, 151) fp32 (or its bf16 K-padded copy `xpad`) -> xres (n*L, D) fp32."""
        w, cfg, T = self.w, self.cfg, self.T
        D, dn, S = cfg["latent_dim"], cfg["dancers"], cfg["seq_len"]
        L = S * dn
        R = n * L
        h = ws.get(tag + "h", (R, D), T)
        if T == torch.float32:
            ops.gemm(x.view(R, 151), w.in_w, w.in_b, ACT_NONE, h, M=R, K=151)
        else:
            Kp = w.in_w.shape[1]
            if xpad is None:
                xpad = ws.get(tag + "xpad", (R, Kp), T, zero=True)
                ops.convert_pad(x, 151, xpad, Kp, R, 151)
            ops.gemm(xpad, w.in_w, w.in_b, ACT_NONE, h, M=R, K=Kp)
        hv = h.view(n * S, dn * D)                                # tokens are frame-major: pure view
        g1 = ws.get(tag + "g1", (n * S, 2 * D), T)
        g2 = ws.get(tag + "g2", (n * S, 2 * D), T)
        ops.gemm(hv, w.fus[0][0], w.fus[0][1], ACT_RELU, g1, M=n * S)
        ops.gemm(g1, w.fus[1][0], w.fus[1][1], ACT_RELU, g2, M=n * S)
        ops.gemm(g2, w.fus[2][0], w.fus[2][1], ACT_NONE, xres.view(-1, dn * D)[: n * S], M=n * S)
        return xres

    # ---------------------------------------------------------------- decoder stack + head
    def _pair(self, a, wgt, bias, xres, keep_x, y, ln_in, eps_in, film, fld, foff, ln_next, plain, rot, n, L, D):
        """`fc` / `linear2` GEMM followed by its FiLM + residual + LayerNorm tail over n samples (model.py:327,334,339)."""
        w = self.w
        rows = n * L
        ops.gemm(a[:rows], wgt, bias, ACT_NONE, y, M=rows)
        ops.film_residual_norm(self.tcd, xres, xres if keep_x else None, y, ln_in, eps_in, film, fld, foff,
                               ln_next, 1e-5, plain, rot, None if rot is None else w.rot_cos,
                               None if rot is None else w.rot_sin, rows, D, L)

    def layers(self, ws, xres, n, Kc, Vc, film, out, tag="ly", shared_front=0):
        """model.py:308-344 x NL + final_layer (model.py:623).  xres (n*L, D) fp32 is consumed in place;
        Kc/Vc (n, S+2, NL*D); film (n, NL*3*2D) fp32 (row pitch film.stride(0)); out (n*L, 151) fp32.
        shared_front = B > 0 (sampler, n == 2B): only xres[:B*L] is filled and samples b and B+b see the same
        x (conditional / unconditional pass), so layer 0's norm1, Q/K/V projections, self-attention and fc are
        evaluated once on B samples and only the FiLM modulation differs (model.py:326-327)."""
        w, cfg, T, tcd = self.w, self.cfg, self.T, self.tcd
        D, dn, S, H, NL = cfg["latent_dim"], cfg["dancers"], cfg["seq_len"], cfg["num_heads"], cfg["num_layers"]
        L, Mm = S * dn, S + 2
        R = n * L
        NLD = Kc.shape[-1]
        plain = ws.get(tag + "plain", (R, D), T)
        rot = ws.get(tag + "rot", (R, D), T)
        qk = ws.get(tag + "qk", (R, 2 * H * HEAD_DIM), T)
        v = ws.get(tag + "v", (R, H * HEAD_DIM), T)
        ctx = ws.get(tag + "ctx", (R, H * HEAD_DIM), T)
        y = ws.get(tag + "y", (R, D), T)                          # block outputs before FiLM (operand dtype)
        ff = ws.get(tag + "ff", (R, cfg["ff_size"]), T)
        fld = film.stride(0)
        scale = 1.0 / math.sqrt(HEAD_DIM)                         # q / temperature, model.py:69,97
        HD = H * HEAD_DIM
        Ly0 = w.layers[0]
        n0 = shared_front if shared_front else n              # samples that go through layer 0's attention block
        R0 = n0 * L
        assert not shared_front or n == 2 * shared_front
        ops.layernorm_rotary(xres, Ly0["n1"][0], Ly0["n1"][1], 1e-5, plain, rot, w.rot_cos, w.rot_sin, R0, D, L)
        for i, Ly in enumerate(w.layers):
            # --- self-attention block (model.py:326-327,374-383)
            ni, Ri = (n0, R0) if i == 0 else (n, R)
            ops.gemm(rot, Ly["sa_qk"], None, ACT_NONE, qk, M=Ri)
            ops.gemm(plain, Ly["sa_v"], None, ACT_NONE, v, M=Ri)
            ops.attention(qk, 2 * HD, L * 2 * HD, qk, 2 * HD, L * 2 * HD, v, HD, L * HD, ctx, HD, L * HD, ni, H, L, L,
                          scale, k_off=HD)
            fuse = fuse_tails() if (T == torch.bfloat16 and D == 512) else 0
            fold = fuse if fold_ln() else 0                   # bit 1: norm3 folded into linear1, bit 2: norm4 into linear3
            if i == 0 and shared_front:
                ops.gemm(ctx, Ly["sa_fc"], None, ACT_NONE, y, M=Ri)
                # unconditional half first (reads the shared x, writes rows [R0, 2*R0)), then the conditional half in place
                ops.film_residual_norm(tcd, xres, xres[R0:], y, Ly["sa_ln"], 1e-6, film[n0:], fld, 0, Ly["n2"], 1e-5,
                                       None, rot[R0:], w.rot_cos, w.rot_sin, R0, D, L)
                ops.film_residual_norm(tcd, xres, xres, y, Ly["sa_ln"], 1e-6, film, fld, 0, Ly["n2"], 1e-5,
                                       None, rot, w.rot_cos, w.rot_sin, R0, D, L)
            elif fuse & 1:
                ops.gemm_film_residual_norm(ctx, Ly["sa_fc"], None, xres, xres, Ly["sa_ln"], 1e-6, film, fld, (3 * i) * 2 * D,
                                            Ly["n2"], 1e-5, None, rot, w.rot_cos_t, w.rot_sin_t, R, L)
            else:
                self._pair(ctx, Ly["sa_fc"], None, xres, True, y, Ly["sa_ln"], 1e-6, film, fld, (3 * i) * 2 * D, Ly["n2"],
                           None, rot, n, L, D)
            # --- cross-attention block (model.py:331-334,386-396)
            ops.gemm(rot, Ly["ca_q"], None, ACT_NONE, v, M=R)    # reuse `v` as the cross-attention query buffer
            ops.attention(v, HD, L * HD, Kc, NLD, Mm * NLD, Vc, NLD, Mm * NLD, ctx, HD, L * HD, n, H, L, Mm, scale,
                          k_off=i * HD, v_off=i * HD)
            if fuse & 2:
                ops.gemm_film_residual_norm(ctx, Ly["ca_fc"], None, xres, xres, Ly["ca_ln"], 1e-6, film, fld,
                                            (3 * i + 1) * 2 * D, None if fold & 2 else Ly["n3"], 1e-5, plain, None, None, None, R, L)
            else:
                self._pair(ctx, Ly["ca_fc"], None, xres, True, y, Ly["ca_ln"], 1e-6, film, fld, (3 * i + 1) * 2 * D, Ly["n3"],
                           plain, None, n, L, D)
            # --- feed-forward block (model.py:338-339,399-401)
            l1 = Ly["l1n"] if fold & 2 else Ly["l1"]
            ops.gemm(plain, l1[0], l1[1], ACT_GELU, ff, M=R)
            if fuse & 4:
                ops.gemm_film_residual_norm(ff, Ly["l2"][0], Ly["l2"][1], xres, None if SKIP_DEAD_X else xres, None, 0.0,
                                            film, fld, (3 * i + 2) * 2 * D, None if fold & 4 else Ly["n4"], 1e-5, plain, None, None,
                                            None, R, L)
            else:
                self._pair(ff, Ly["l2"][0], Ly["l2"][1], xres, not SKIP_DEAD_X, y, None, 0.0, film, fld, (3 * i + 2) * 2 * D,
                           Ly["n4"], plain, None, n, L, D)
            # --- x = linear3(norm4(x)) is the layer's return value (model.py:344,371)
            l3 = Ly["l3n"] if fold & 4 else Ly["l3"]
            if i + 1 < NL:
                ops.gemm(plain, l3[0], l3[1], ACT_NONE, xres, M=R)
                nx = w.layers[i + 1]["n1"]
                ops.layernorm_rotary(xres, nx[0], nx[1], 1e-5, plain, rot, w.rot_cos, w.rot_sin, R, D, L)
            else:
                ops.gemm(plain, l3[0], l3[1], ACT_NONE, y, M=R)
        ops.gemm(y, w.fin[0], w.fin[1], ACT_NONE, out, M=R, N=151, ldc=151)
        return out

    # ---------------------------------------------------------------- one full pass (general API)
    def forward(self, ws, x, cond_embed, times, keep):
        """DanceDecoder.forward (model.py:548-624) for per-sample times / keep mask."""
        cfg = self.cfg
        n = x.shape[0]
        D, L = cfg["latent_dim"], cfg["seq_len"] * cfg["dancers"]
        tok, ch = self.music_encode(ws, cond_embed, keep)
        t_lin, tt = self.time_path(ws, times)
        film = self.film_table(ws, t_lin, ch, keep)
        Kc, Vc = self.memory_kv(ws, tok, tt)
        xres = ws.get("xres", (n * L, D), torch.float32)
        self.front(ws, x, n, xres)
        out = torch.empty((n, L, 151), dtype=torch.float32, device=x.device)
        self.layers(ws, xres, n, Kc, Vc, film, out.view(n * L, 151))
        return out
