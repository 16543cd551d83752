"""TEST INFRASTRUCTURE — functional CPU restatement (torch fp32) of the TCDiff hot path.

Stateless functions over a plain ``state_dict`` (reference key names).  Every
function cites the reference lines it restates (paths relative to /root/reference).
Validated against the unmodified reference by ``oracle/make_golden.py`` and
``tests/test_oracle_vs_reference.py`` (in the build container) and against the committed
``tests/golden/*.pt`` everywhere.  The pytorch3d boundary is **parity unpinned** (oracle/p3d.py).

Eval-mode semantics only (dropout is the identity): train-mode dropout streams cannot be
matched across implementations (SURVEY §7 "Randomness for parity").
"""
import math

import numpy as np
import torch
import torch.nn.functional as F

from . import p3d

# ----------------------------------------------------------------------------- constants
# vis.py:48-73 (kinematic tree) and vis.py:76-101 (bone offsets, metres)
SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]
SMPL_OFFSETS = [
    [0.0, 0.0, 0.0], [0.05858135, -0.08228004, -0.01766408], [-0.06030973, -0.09051332, -0.01354254],
    [0.00443945, 0.12440352, -0.03838522], [0.04345142, -0.38646945, 0.008037],
    [-0.04325663, -0.38368791, -0.00484304], [0.00448844, 0.1379564, 0.02682033],
    [-0.01479032, -0.42687458, -0.037428], [0.01905555, -0.4200455, -0.03456167],
    [-0.00226458, 0.05603239, 0.00285505], [0.04105436, -0.06028581, 0.12204243],
    [-0.03483987, -0.06210566, 0.13032329], [-0.0133902, 0.21163553, -0.03346758],
    [0.07170245, 0.11399969, -0.01889817], [-0.08295366, 0.11247234, -0.02370739],
    [0.01011321, 0.08893734, 0.05040987], [0.12292141, 0.04520509, -0.019046],
    [-0.11322832, 0.04685326, -0.00847207], [0.2553319, -0.01564902, -0.02294649],
    [-0.26012748, -0.01436928, -0.03126873], [0.26570925, 0.01269811, -0.00737473],
    [-0.26910836, 0.00679372, -0.00602676], [0.08669055, -0.01063603, -0.01559429],
    [-0.0887537, -0.00865157, -0.01010708],
]
FOOT_JOINTS = [7, 8, 10, 11]                       # model/diffusion.py:720
LOSS_WEIGHTS = (0.636, 2.964, 0.646, 10.942)       # model/diffusion.py:735-740
HEAD_DIM = 64                                      # model/model.py:55,532 (d_k hard-coded)


# ----------------------------------------------------------------------------- schedule
def make_betas(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """model/utils.py:67-99 — float64 numpy betas."""
    if schedule == "linear":
        b = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=torch.float64) ** 2
    elif schedule == "cosine":
        ts = torch.arange(n_timestep + 1, dtype=torch.float64) / n_timestep + cosine_s
        a = torch.cos(ts / (1 + cosine_s) * np.pi / 2).pow(2)
        a = a / a[0]
        b = np.clip(1 - a[1:] / a[:-1], a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        b = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64)
    elif schedule == "sqrt":
        b = torch.linspace(linear_start, linear_end, n_timestep, dtype=torch.float64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return b.numpy() if isinstance(b, torch.Tensor) else np.asarray(b)


def make_schedule(schedule="cosine", n_timestep=1000, use_p2=False):
    """model/diffusion.py:109-169 — the 13 fp32 buffers, in the reference's precision order
    (float64 betas -> fp32 tensor -> fp32 cumprod ...)."""
    betas = torch.Tensor(make_betas(schedule, n_timestep))
    alphas = 1.0 - betas
    ac = torch.cumprod(alphas, 0)
    ac_prev = torch.cat([torch.ones(1), ac[:-1]])
    pv = betas * (1.0 - ac_prev) / (1.0 - ac)
    gamma = 0.5 if use_p2 else 0
    return {
        "betas": betas,
        "alphas_cumprod": ac,
        "alphas_cumprod_prev": ac_prev,
        "sqrt_alphas_cumprod": torch.sqrt(ac),
        "sqrt_one_minus_alphas_cumprod": torch.sqrt(1.0 - ac),
        "log_one_minus_alphas_cumprod": torch.log(1.0 - ac),
        "sqrt_recip_alphas_cumprod": torch.sqrt(1.0 / ac),
        "sqrt_recipm1_alphas_cumprod": torch.sqrt(1.0 / ac - 1),
        "posterior_variance": pv,
        "posterior_log_variance_clipped": torch.log(torch.clamp(pv, min=1e-20)),
        "posterior_mean_coef1": betas * np.sqrt(ac_prev) / (1.0 - ac),
        "posterior_mean_coef2": (1.0 - ac_prev) * np.sqrt(alphas) / (1.0 - ac),
        "p2_loss_weight": (1 + ac / (1 - ac)) ** -gamma,
    }


def ddim_times(n_timestep=1000, sampling_timesteps=50):
    """model/diffusion.py:389-391 — [(999,979),...,(19,-1)]."""
    t = torch.linspace(-1, n_timestep - 1, steps=sampling_timesteps + 1)
    t = list(reversed(t.int().tolist()))
    return list(zip(t[:-1], t[1:]))


def ddim_coeffs(sched, time, time_next, eta=1.0):
    """model/diffusion.py:415-419 — fp32 0-d tensor arithmetic exactly as the reference does it."""
    a = sched["alphas_cumprod"][time]
    an = sched["alphas_cumprod"][time_next]
    sigma = eta * ((1 - a / an) * (1 - an) / (1 - a)).sqrt()
    c = (1 - an - sigma ** 2).sqrt()
    return an.sqrt(), c, sigma


# ----------------------------------------------------------------------------- embeddings
def sinusoidal_pos_emb(times, dim):
    """model/utils.py:41-48."""
    half = dim // 2
    e = math.log(10000) / (half - 1)
    e = torch.exp(torch.arange(half) * -e)
    e = times[:, None] * e[None, :]
    return torch.cat((e.sin(), e.cos()), dim=-1)


def rotary_angles(freqs, seq_len):
    """model/rotary_embedding_torch.py:115-130 — (L, dim) angle table, each freq repeated twice."""
    ang = torch.arange(seq_len).type(freqs.dtype)[:, None] * freqs[None, :]
    return ang.repeat_interleave(2, dim=-1)


def rotary_freqs(dim, theta=10000):
    """model/rotary_embedding_torch.py:89-92."""
    return 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))


def apply_rotary(freqs, t):
    """model/rotary_embedding_torch.py:39-59,107-113 — rotate the full feature vector of every token
    by its flat position index; pairs are (2i, 2i+1)."""
    ang = rotary_angles(freqs, t.shape[-2]).to(t)
    pair = t.reshape(t.shape[:-1] + (-1, 2))
    swapped = torch.stack((-pair[..., 1], pair[..., 0]), dim=-1).reshape(t.shape)
    return t * ang.cos() + swapped * ang.sin()


# ----------------------------------------------------------------------------- denoiser
def _lin(sd, name, x, bias=True):
    return F.linear(x, sd[name + ".weight"], sd[name + ".bias"] if bias else None)


def _ln(sd, name, x, eps=1e-5):
    return F.layer_norm(x, (x.shape[-1],), sd[name + ".weight"], sd[name + ".bias"], eps)


def _heads(x, nh):
    b, l, _ = x.shape
    return x.view(b, l, nh, HEAD_DIM).transpose(1, 2)


# Training-mode dropout (model/model.py:98,103,240-245,383,396,400-401): torch's Bernoulli stream cannot be shared with
# another implementation, so the oracle takes the MASKS from a hook: hook(kind, layer, k, x) -> x * mask / (1-p), with
# kind "enc" (music encoder layer: k = 0 attention probabilities, 1 dropout1, 2 FFN inner, 3 dropout2) or "dec"
# (decoder layer: k = 0 / 3 self / cross attention probabilities, 1 / 4 after fc, 2 / 5 dropout1 / dropout2, 6 FFN
# inner, 7 dropout3).  No hook = eval mode.
_DROP_HOOK = None


class dropout_hook:
    def __init__(self, fn):
        self.fn = fn

    def __enter__(self):
        global _DROP_HOOK
        self.prev, _DROP_HOOK = _DROP_HOOK, self.fn

    def __exit__(self, *a):
        global _DROP_HOOK
        _DROP_HOOK = self.prev


def _drop(p, k, x):
    if _DROP_HOOK is None:
        return x
    parts = p.split(".")
    kind = "enc" if parts[0] == "cond_encoder" else "dec"
    layer = int(parts[1] if kind == "enc" else parts[2])
    return _DROP_HOOK(kind, layer, k, x)


def sbi_msa(sd, p, q_in, k_in, v_in, nh, k0=0):
    """model/model.py:71-107 with trj_dist=None: bias-free projections, softmax((q/8) k^T) v, fc,
    LayerNorm(eps=1e-6); no residual; the q.emb^T 'indexed matrix' never reaches the output."""
    q = _heads(_lin(sd, p + ".w_qs", q_in, False), nh)
    k = _heads(_lin(sd, p + ".w_ks", k_in, False), nh)
    v = _heads(_lin(sd, p + ".w_vs", v_in, False), nh)
    att = torch.softmax(torch.matmul(q / (HEAD_DIM ** 0.5), k.transpose(2, 3)), dim=-1)
    att = _drop(p, k0, att)                                                  # model.py:98
    o = torch.matmul(att, v).transpose(1, 2).reshape(q_in.shape[0], q_in.shape[1], -1)
    return _ln(sd, p + ".layer_norm", _drop(p, k0 + 1, _lin(sd, p + ".fc", o, False)), eps=1e-6)    # :103-106


def film(sd, p, t):
    """model/model.py:164-173 — (scale, shift) = chunk(Linear(Mish(t))); scale first."""
    ss = _lin(sd, p + ".block.1", F.mish(t))[:, None, :]
    return ss.chunk(2, dim=-1)


def decoder_layer(sd, p, x, memory, t, freqs, nh):
    """model/model.py:308-371 (norm_first) — note the returned value is linear3(norm4(x)); the
    traj_Modulation MLP result `out` is discarded by the reference (model.py:344-355,371)."""
    n1 = _ln(sd, p + ".norm1", x)
    qk = apply_rotary(freqs, n1)                                            # model.py:375
    s, b = film(sd, p + ".film1", t)
    x = x + (s + 1) * _drop(p, 2, sbi_msa(sd, p + ".self_attn", qk, qk, n1, nh, 0)) + b       # model.py:326-327,383
    n2 = _ln(sd, p + ".norm2", x)
    s, b = film(sd, p + ".film2", t)
    x = x + (s + 1) * _drop(p, 5, sbi_msa(sd, p + ".multihead_attn", apply_rotary(freqs, n2),
                                          apply_rotary(freqs, memory), memory, nh, 3)) + b   # model.py:331-334,386-396
    n3 = _ln(sd, p + ".norm3", x)
    s, b = film(sd, p + ".film3", t)
    ff = _drop(p, 7, _lin(sd, p + ".linear2", _drop(p, 6, F.gelu(_lin(sd, p + ".linear1", n3)))))   # :400-401
    x = x + (s + 1) * ff + b                                                  # :338-339
    return _lin(sd, p + ".linear3", _ln(sd, p + ".norm4", x))                 # model.py:344


def music_encoder_layer(sd, p, x, freqs, nh):
    """model/model.py:211-245 (norm_first) with nn.MultiheadAttention(batch_first) written out:
    packed in_proj with bias, 1/sqrt(64) scaling, out_proj with bias; q,k get the rotary input."""
    n = _ln(sd, p + ".norm1", x)
    qk = apply_rotary(freqs, n)
    W, bias = sd[p + ".self_attn.in_proj_weight"], sd[p + ".self_attn.in_proj_bias"]
    d = x.shape[-1]
    hd = d // nh
    B, L, _ = x.shape
    q = F.linear(qk, W[:d], bias[:d]).view(B, L, nh, hd).transpose(1, 2)
    k = F.linear(qk, W[d:2 * d], bias[d:2 * d]).view(B, L, nh, hd).transpose(1, 2)
    v = F.linear(n, W[2 * d:], bias[2 * d:]).view(B, L, nh, hd).transpose(1, 2)
    att = torch.softmax(torch.matmul(q, k.transpose(2, 3)) / math.sqrt(hd), dim=-1)
    att = _drop(p, 0, att)                                                    # nn.MultiheadAttention(dropout=...)
    o = torch.matmul(att, v).transpose(1, 2).reshape(B, L, d)
    x = x + _drop(p, 1, _lin(sd, p + ".self_attn.out_proj", o))                # model.py:240
    n = _ln(sd, p + ".norm2", x)
    return x + _drop(p, 3, _lin(sd, p + ".linear2", _drop(p, 2, F.gelu(_lin(sd, p + ".linear1", n)))))   # :244-245


def infer_config(sd):
    """Recover the constructor hyper-parameters from tensor shapes (model/model.py:417-540)."""
    d = sd["input_projection.weight"].shape[0]
    nfeats = sd["input_projection.weight"].shape[1]
    dn = sd["relative_projection_layer.0.weight"].shape[1] // d
    fm = sd["cond_projection.0.weight"].shape[0]
    seq_len = sd["null_cond_embed"].shape[1]
    layers = 1 + max(int(k.split(".")[2]) for k in sd if k.startswith("seqTransDecoder.stack."))
    nh = sd["embeddings_table.weight"].shape[1] // HEAD_DIM
    ff = sd["seqTransDecoder.stack.0.linear1.weight"].shape[0]
    return dict(nfeats=nfeats, latent_dim=d, dancers=dn, cond_feature_dim=fm, seq_len=seq_len,
                num_layers=layers, num_heads=nh, ff_size=ff)


def music_tokens(sd, cond_embed, cfg=None):
    """model/model.py:572-581 — pair consecutive 60 fps frames, project, 2 encoder layers."""
    cfg = cfg or infer_config(sd)
    B, T, _ = cond_embed.shape
    if T % 2 == 1:
        cond_embed = cond_embed[:, :-1, :]
    c = cond_embed.reshape(B, T // 2, -1).float()
    c = _lin(sd, "cond_projection.2", F.relu(_lin(sd, "cond_projection.0", c)))
    freqs = sd.get("rotary.freqs", rotary_freqs(cfg["latent_dim"]))
    for i in range(2):
        c = music_encoder_layer(sd, f"cond_encoder.{i}", c, freqs, cfg["num_heads"])
    return c


def dance_decoder_forward(sd, x, cond_embed, times, keep_mask=None, cond_drop_prob=0.0):
    """model/model.py:548-624 — one denoiser pass.  ``keep_mask`` (B,) bool overrides the Bernoulli
    draw of model.py:567; with cond_drop_prob in {0,1} the mask is all-True / all-False."""
    cfg = infer_config(sd)
    D, dn, S, nh = cfg["latent_dim"], cfg["dancers"], cfg["seq_len"], cfg["num_heads"]
    B = x.shape[0]
    x = x.reshape(B, -1, 151)                                               # model.py:553
    h = _lin(sd, "input_projection", x)                                     # :560
    g = h.reshape(B, S, D * dn)                                             # :561 fusion projection
    g = F.relu(_lin(sd, "relative_projection_layer.0", g))
    g = F.relu(_lin(sd, "relative_projection_layer.2", g))
    h = _lin(sd, "relative_projection_layer.4", g).reshape(B, dn * S, D)
    if keep_mask is None:
        if cond_drop_prob == 0:
            keep_mask = torch.ones(B, dtype=torch.bool)
        elif cond_drop_prob == 1:
            keep_mask = torch.zeros(B, dtype=torch.bool)
        else:
            keep_mask = torch.zeros(B).float().uniform_(0, 1) < (1 - cond_drop_prob)
    tokens = music_tokens(sd, cond_embed, cfg)
    tokens = torch.where(keep_mask[:, None, None], tokens, sd["null_cond_embed"].to(tokens.dtype))  # :589
    pooled = tokens.mean(dim=-2)                                            # :593
    ch = _ln(sd, "non_attn_cond_projection.0", pooled)                      # :597
    ch = _lin(sd, "non_attn_cond_projection.3", F.silu(_lin(sd, "non_attn_cond_projection.1", ch)))
    th = F.mish(_lin(sd, "time_mlp.1", sinusoidal_pos_emb(times, D)))       # :601
    t = _lin(sd, "to_time_cond.0", th)                                      # :604
    tt = _lin(sd, "to_time_tokens.0", th).reshape(B, 2, D)                  # :605
    t = t + torch.where(keep_mask[:, None], ch, sd["null_cond_hidden"].to(t.dtype))  # :609-612
    mem = _ln(sd, "norm_cond", torch.cat((tokens, tt), dim=-2))             # :615-616
    freqs = sd.get("rotary.freqs", rotary_freqs(D))
    for i in range(cfg["num_layers"]):
        h = decoder_layer(sd, f"seqTransDecoder.stack.{i}", h, mem, t, freqs, nh)
    return _lin(sd, "final_layer", h)                                       # :623


def guided_forward(sd, x, cond_embed, times, w):
    """model/model.py:542-546."""
    unc = dance_decoder_forward(sd, x, cond_embed, times, cond_drop_prob=1)
    con = dance_decoder_forward(sd, x, cond_embed, times, cond_drop_prob=0)
    return unc + (con - unc) * w


# ----------------------------------------------------------------------------- diffusion
def _ex(a, t, nd):
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (nd - 1)))          # model/utils.py:61-64


def q_sample(sched, x_start, t, noise):
    """model/diffusion.py:625-634."""
    return (_ex(sched["sqrt_alphas_cumprod"], t, x_start.dim()) * x_start
            + _ex(sched["sqrt_one_minus_alphas_cumprod"], t, x_start.dim()) * noise)


def _inpaint_traj(x, x0r, dn):
    """model/diffusion.py:396-403,427-431 — overwrite channels 4,5 from x_0[...,0:2]."""
    B = x.shape[0]
    xv = x.reshape(B, 150, dn, 151)
    xv[:, :, :, [4, 5]] = x0r[:, :, :, [0, 1]]
    return xv.reshape(B, 150 * dn, 151)


FOOT_IDX = (1, 2, 3, 4, 5, 7, 8, 10, 11)            # model/diffusion.py:308


def _foot_copy(x, x0f, dn):
    """model/diffusion.py:307-309,343-344 — impose the 6-D rotation channels 7+(i-1)*6..7+i*6 on frames 75:120."""
    B = x.shape[0]
    xv = x.reshape(B, 150, dn, 151)
    for i in FOOT_IDX:
        a, b = 4 + 3 + (i - 1) * 6, 4 + 3 + i * 6
        xv[:, 75:120, :, a:b] = x0f[:, 75:120, :, a:b]
    return xv.reshape(B, 150 * dn, 151)


def ddim_sample(sd, sched, shape, cond, x_0, noise_bank, guidance_weight=2.0, clip=True,
                n_timestep=1000, sampling_timesteps=50, eta=1.0, trace=None, long_mode=False, seq_len=150,
                footwork=False):
    """model/diffusion.py:385-442.  ``noise_bank[0]`` is x_T, ``noise_bank[1:]`` the per-step draws
    (one per step with time_next >= 0) in call order.  ``trace`` (list) collects (x_t, x_start).
    long_mode=True restates long_ddim_sample (:445-515): per-step guidance ramp and window hand-over.
    footwork=True restates ddim_sample_Footwork (:288-383): x_0 is a full (B, S*dn, 151) motion."""
    B = shape[0]
    x = noise_bank[0].clone()
    dn = shape[1] // 150
    x0r = x0f = None
    if x_0 is not None:
        if footwork:
            x0f = x_0.reshape(-1, 150, dn, 151)
            x0r = x0f[..., :3]
        else:
            x0r = x_0.reshape(-1, 150, dn, 3)
        x = _inpaint_traj(x, x0r, dn)
        if footwork:
            x = _foot_copy(x, x0f, dn)
    step_w = [guidance_weight] * sampling_timesteps
    if long_mode:
        step_w = list(np.clip(np.linspace(0, guidance_weight * 2, sampling_timesteps), None, guidance_weight))  # :454
    half = seq_len // 2
    k = 1
    for step, (time, time_next) in enumerate(ddim_times(n_timestep, sampling_timesteps)):
        tc = torch.full((B,), time, dtype=torch.long)
        x_start = guided_forward(sd, x, cond, tc, step_w[step])             # :195-204
        if clip:
            x_start = x_start.clamp(-1.0, 1.0)
        pred_noise = ((_ex(sched["sqrt_recip_alphas_cumprod"], tc, 3) * x - x_start)
                      / _ex(sched["sqrt_recipm1_alphas_cumprod"], tc, 3))   # :189-193
        if trace is not None:
            trace.append((x.clone(), x_start.clone()))
        if time_next < 0:
            x = x_start
            continue
        sa, c, sigma = ddim_coeffs(sched, time, time_next, eta)
        x = x_start * sa + c * pred_noise + sigma * noise_bank[k]           # :423-425
        k += 1
        if x0r is not None:
            x = _inpaint_traj(x, x0r, dn)
        if footwork and x0f is not None:
            x = _foot_copy(x, x0f, dn)
        if long_mode and time > 0:                                          # :502-506
            xv = x.reshape(B, seq_len, shape[1] // seq_len, shape[2])
            xv[1:, :half] = xv[:-1, half:].clone()
            x = xv.reshape(B, -1, shape[2])
    if x0r is not None:
        x = _inpaint_traj(x, x0r, dn)                                       # :434-440
    if footwork and x0f is not None:                                        # :356-381 linear-interpolation fusion
        width = 10
        w = torch.from_numpy(np.linspace(0, 1, width)).to(x)[None, :, None, None]
        xv = x.reshape(B, 150, dn, 151)
        for i in FOOT_IDX:
            a, b = 4 + 3 + (i - 1) * 6, 4 + 3 + i * 6
            xv[:, 75:75 + width, :, a:b] = w * x0f[:, 75:75 + width, :, a:b] + (1 - w) * xv[:, 75:75 + width, :, a:b]
            xv[:, 75 + width:-width, :, a:b] = x0f[:, 75 + width:-width, :, a:b]
            xv[:, 120 - width:120, :, a:b] = (1 - w) * x0f[:, 120 - width:120, :, a:b] + w * xv[:, 120 - width:120, :, a:b]
        x = xv.reshape(B, 150 * dn, 151)
    return x


def _x_recon(sched, x, t, out, predict_epsilon):
    """predict_start_from_noise + clamp (model/diffusion.py:176-187,226-231)."""
    if predict_epsilon:
        out = _ex(sched["sqrt_recip_alphas_cumprod"], t, 3) * x - _ex(sched["sqrt_recipm1_alphas_cumprod"], t, 3) * out
    return out.clamp(-1.0, 1.0)


def p_sample_loop(sd, sched, shape, cond, noise_bank, guidance_weight=2.0, n_timestep=1000,
                  start_point=None, long_mode=False, constraint=None, predict_epsilon=False):
    """model/diffusion.py:206-286 with clip_denoised=True.  predict_epsilon=True: the network output is the noise and
    x_recon = sqrt_recip_alphas_cumprod x - sqrt_recipm1_alphas_cumprod out (:176-187,226-228).
    ``noise_bank[0]`` = x_T; ``noise_bank[1+j]`` = draw of the j-th p_sample call.
    long_mode=True restates long_inpaint_loop (:559-609): x[1:, :half] = x[:-1, half:] after every step i > 0.
    constraint={"mask", "value"} restates inpaint_loop (:518-557): after every p_sample the masked entries are replaced
    by q_sample(value, i - 1) (by x itself at i == 0); the bank then holds the draws in the reference's call order,
    i.e. p_sample's randn_like followed by q_sample's randn_like for every step with i > 0."""
    B = shape[0]
    x = noise_bank[0].clone()
    start_point = n_timestep if start_point is None else start_point
    if constraint is not None:
        assert not long_mode
        k = 1
        for i in reversed(range(0, start_point)):
            t = torch.full((B,), i, dtype=torch.long)
            w = min(guidance_weight, 0) if i > 1.0 * n_timestep else (min(guidance_weight, 1) if i < 0.1 * n_timestep
                                                                     else guidance_weight)
            x_recon = _x_recon(sched, x, t, guided_forward(sd, x, cond, t, w), predict_epsilon)
            mean = (_ex(sched["posterior_mean_coef1"], t, 3) * x_recon + _ex(sched["posterior_mean_coef2"], t, 3) * x)
            logvar = _ex(sched["posterior_log_variance_clipped"], t, 3)
            nz = (1 - (t == 0).float()).reshape(B, 1, 1)
            x = mean + nz * (0.5 * logvar).exp() * noise_bank[k]
            k += 1
            if i > 0:                                                       # :547
                value_ = q_sample(sched, constraint["value"], t - 1, noise_bank[k])
                k += 1
            else:
                value_ = x
            x = value_ * constraint["mask"] + (1.0 - constraint["mask"]) * x   # :549
        return x
    for j, i in enumerate(reversed(range(0, start_point))):
        t = torch.full((B,), i, dtype=torch.long)
        if i > 1.0 * n_timestep:                                            # :219-224
            w = min(guidance_weight, 0)
        elif i < 0.1 * n_timestep:
            w = min(guidance_weight, 1)
        else:
            w = guidance_weight
        x_recon = _x_recon(sched, x, t, guided_forward(sd, x, cond, t, w), predict_epsilon)   # :226-231
        mean = (_ex(sched["posterior_mean_coef1"], t, 3) * x_recon
                + _ex(sched["posterior_mean_coef2"], t, 3) * x)             # :206-210
        logvar = _ex(sched["posterior_log_variance_clipped"], t, 3)
        nz = (1 - (t == 0).float()).reshape(B, 1, 1)
        x = mean + nz * (0.5 * logvar).exp() * noise_bank[1 + j]            # :246-251
        if long_mode and i > 0:                                             # :599-601
            half = x.shape[1] // 2
            x[1:, :half] = x[:-1, half:].clone()
    return x


# ----------------------------------------------------------------------------- kinematics / loss
def ax_from_6v(q):
    """dataset/quaternion.py:28-32."""
    assert q.shape[-1] == 6
    return p3d.matrix_to_axis_angle(p3d.rotation_6d_to_matrix(q))


def smpl_forward(rotations, root_positions):
    """vis.py:358-406 — axis-angle -> quaternion, walk the parent chain."""
    assert rotations.dim() == 4 and root_positions.dim() == 3
    q = p3d.axis_angle_to_quaternion(rotations)
    off = torch.tensor(SMPL_OFFSETS, dtype=q.dtype)
    has_child = [False] * 24
    for j, pa in enumerate(SMPL_PARENTS):
        if pa >= 0:
            has_child[pa] = True
    pos, rot = [], []
    for j, pa in enumerate(SMPL_PARENTS):
        if pa < 0:
            pos.append(root_positions)
            rot.append(q[:, :, 0])
        else:
            o = off[j].expand(q.shape[0], q.shape[1], 3)
            pos.append(p3d.quaternion_apply(rot[pa], o) + pos[pa])
            rot.append(p3d.quaternion_multiply(rot[pa], q[:, :, j]) if has_child[j] else None)
    return torch.stack(pos, dim=2)


def loss_terms(model_out, target, p2w=None, loss_type="l2"):
    """model/diffusion.py:664-741 — the four weighted losses from (B,S,dn,151) prediction and target; loss_type "l2" =
    F.mse_loss, "l1" = F.l1_loss (:172; the reference constructor's default is "l1", TCDiff.py passes "l2").
    Returns (total, (recon, vel, fk, foot)) already weighted like the reference."""
    B, S, dn, C = model_out.shape
    if p2w is None:
        p2w = torch.ones(B)
    el = (lambda d: d ** 2) if loss_type == "l2" else (lambda d: d.abs())
    rec = el(model_out - target).reshape(B, -1).mean(1) * p2w
    mc, mo = model_out[..., :4], model_out[..., 4:]
    tg = target[..., 4:]
    vel = el((mo[:, 1:] - mo[:, :-1]) - (tg[:, 1:] - tg[:, :-1])).reshape(B, -1).mean(1) * p2w
    mq = ax_from_6v(mo[..., 3:].reshape(B, S * dn, -1, 6))
    tq = ax_from_6v(tg[..., 3:].reshape(B, S * dn, -1, 6))
    mxp = smpl_forward(mq, mo[..., :3].reshape(B, S * dn, 3))
    txp = smpl_forward(tq, tg[..., :3].reshape(B, S * dn, 3))
    fk = el((mxp[:, :, 1:] - mxp[:, :, 0:1]) - (txp[:, :, 1:] - txp[:, :, 0:1])).reshape(B, -1).mean(1) * p2w
    feet = mxp.reshape(B, S, dn, 24, 3)[:, :, :, FOOT_JOINTS]
    fv = torch.zeros_like(feet)
    fv[:, :-1] = feet[:, 1:] - feet[:, :-1]
    fv = torch.where((mc > 0.95)[..., None], fv, torch.zeros_like(fv))      # :722,729
    foot = el(fv).reshape(B, -1).mean(1)
    w = LOSS_WEIGHTS
    losses = (w[0] * rec.mean(), w[1] * vel.mean(), w[2] * fk.mean(), w[3] * foot.mean())
    return sum(losses), losses


def p_losses(sd, sched, x_start, cond, t, noise, keep_mask, loss_type="l2", predict_epsilon=False):
    """model/diffusion.py:636-741, eval-mode network.  predict_epsilon=True: the target is the noise (:657-658)."""
    B, dn, S, C = x_start.shape
    xs = x_start.permute(0, 2, 1, 3)
    xn = q_sample(sched, xs, t, noise)
    xn[:, :, :, [4, 5]] = xs[:, :, :, [4, 5]]                                # :650
    out = dance_decoder_forward(sd, xn.reshape(B, S * dn, C), cond, t, keep_mask=keep_mask)
    p2w = sched["p2_loss_weight"].gather(-1, t)
    target = noise if predict_epsilon else xs
    return loss_terms(out.reshape(B, S, dn, C), target.reshape(B, S, dn, C), p2w, loss_type)


# ---------------------------------------------------------------------------------------------------------------------
# Optimizer step ("next" row N1): Adan (model/adan.py:33-123, restart_cond=None) and the EMA blend
# (model/diffusion.py:61-76), restated functionally over lists of tensors.
def adan_init(params):
    """Per-parameter state as the reference creates it lazily (model/adan.py:56-61)."""
    return [dict(step=0, prev_grad=torch.zeros_like(p), m=torch.zeros_like(p), v=torch.zeros_like(p),
                 n=torch.zeros_like(p)) for p in params]


def adan_step(params, grads, state, lr=1e-3, betas=(0.02, 0.08, 0.01), eps=1e-8, weight_decay=0.0):
    """One update in place on `params` / `state`; parameters whose grad is None are skipped (model/adan.py:47-49)."""
    b1, b2, b3 = betas
    for p, g, st in zip(params, grads, state):
        if g is None:
            continue
        step = st["step"]
        if step > 0:                                                  # :70 — the first call leaves m, v, n at zero
            st["m"].mul_(1 - b1).add_(g, alpha=b1)                    # :75
            diff = g - st["prev_grad"]                                # :77
            st["v"].mul_(1 - b2).add_(diff, alpha=b2)                 # :79
            nxt = (g + (1 - b2) * diff) ** 2                          # :81
            st["n"].mul_(1 - b3).add_(nxt, alpha=b3)                  # :83
        step += 1                                                     # :87
        cm, cv, cn = (1 / (1 - (1 - b) ** step) for b in (b1, b2, b3))        # :89-91
        wss = lr / (st["n"] * cn).sqrt().add_(eps)                    # :96
        p.addcmul_(wss, st["m"] * cm + (1 - b2) * st["v"] * cv, value=-1.0).div_(1 + weight_decay * lr)   # :98-104
        st["prev_grad"].copy_(g)                                      # :120
        st["step"] = step


def ema_update(ma_params, cur_params, beta=0.9999):
    """EMA.update_model_average (model/diffusion.py:66-76): ma = ma*beta + (1-beta)*cur, per tensor."""
    for ma, cur in zip(ma_params, cur_params):
        ma.copy_(ma * beta + (1 - beta) * cur)


# ---------------------------------------------------------------------------------------------------------------------
# Post-sampling stage ("next" row N2): un-normalise, split contact, 6D -> axis-angle, FK to joint positions
# (model/diffusion.py:811-838,942-955) and long-mode stitching (:841-909; dataset/quaternion.py:35-71).
def unnormalize(x, min_, scale_):
    """Normalizer.unnormalize (dataset/preprocess.py:39-43) + MinMaxScaler.inverse_transform (dataset/scaler.py:80-83)."""
    x = torch.clip(x, -1, 1)
    x = x - min_[-x.shape[-1]:]
    return x / scale_[-x.shape[-1]:]


def quat_slerp(x, y, a):
    """dataset/quaternion.py:35-71 (functional: inputs are not modified)."""
    ln = torch.sum(x * y, dim=-1)
    neg = ln < 0.0
    ln = torch.where(neg, -ln, ln)
    y = torch.where(neg[..., None], -y, y)
    a = torch.zeros_like(x[..., 0]) + a
    linear = (1.0 - ln) < 0.01
    omegas = torch.arccos(ln)
    sinoms = torch.sin(omegas)
    amount0 = torch.where(linear, 1.0 - a, torch.sin((1.0 - a) * omegas) / sinoms)
    amount1 = torch.where(linear, a, torch.sin(a * omegas) / sinoms)
    return amount0[..., None] * x + amount1[..., None] * y


def samples_to_poses(samples, min_, scale_, dn, mode="normal"):
    """samples (b, 150*dn, 151) normalised, frame-major tokens.
    normal -> dict(contact (b, dn, 150, 4), smpl_trans (b, 150*dn, 3), smpl_poses (b, 150*dn, 24, 3) axis-angle,
                   full_pose (b, dn, 150, 24, 3))
    long   -> the b windows of ONE song stitched with half-window overlap: dict(smpl_trans (F, dn, 3),
              smpl_poses (F, dn, 24, 3), full_pose (dn, F, 24, 3)), F = 150 + 75 (b - 1)."""
    b, sl, _ = samples.shape
    S = 150
    x = unnormalize(samples.reshape(-1, 151), min_, scale_).reshape(b, S, sl // S, 151)          # :812-815
    contact, x = x[..., :4], x[..., 4:]                                                          # :819-821
    x = x.reshape(b, -1, 147)
    pos = x[:, :, :3]                                                                            # :836
    q = ax_from_6v(x[:, :, 3:].reshape(b, -1, 24, 6))                                            # :837-839
    if mode == "normal":
        poses = smpl_forward(q, pos)                                                             # :942
        return dict(contact=contact.permute(0, 2, 1, 3), smpl_trans=pos, smpl_poses=q,
                    full_pose=poses.reshape(b, S, dn, 24, 3).permute(0, 2, 1, 3, 4))             # :944-955
    half = S // 2
    pos = pos.reshape(b, S, dn, 3)
    q = q.reshape(b, S, dn, 24, 3)
    fade_out = torch.ones(1, S, 1)
    fade_in = torch.ones(1, S, 1)
    fade_out[:, half:, :] = torch.linspace(1, 0, half)[None, :, None]
    fade_in[:, :half, :] = torch.linspace(0, 1, half)[None, :, None]
    w = torch.linspace(0, 1, half)[None, :, None]
    F_ = S + half * (b - 1)
    pos_all, q_all = [], []
    for d in range(dn):                                                                          # :855
        cur_pos = pos[:, :, d, :].clone()
        cur_q = q[:, :, d]
        cur_pos[:-1] *= fade_out                                                                 # :868-869
        cur_pos[1:] *= fade_in
        full_pos = torch.zeros(F_, 3)
        for i in range(b):
            full_pos[i * half: i * half + S] += cur_pos[i]                                       # :871-875
        left, right = p3d.axis_angle_to_quaternion(cur_q[:-1, half:]), p3d.axis_angle_to_quaternion(cur_q[1:, :half])
        merged = p3d.quaternion_to_axis_angle(quat_slerp(left, right, w))                        # :880-888
        full_q = torch.zeros(F_, 24, 3)
        full_q[:half] += cur_q[0, :half]
        for i in range(b - 1):
            full_q[half * (i + 1): half * (i + 2)] += merged[i]
        full_q[half * b: half * (b + 1)] += cur_q[-1, half:]                                     # :890-896
        pos_all.append(full_pos)
        q_all.append(full_q)
    full_pos = torch.stack(pos_all, dim=1)                                                       # (F, dn, 3)
    full_q = torch.stack(q_all, dim=1)                                                           # (F, dn, 24, 3)
    pose = smpl_forward(full_q.reshape(1, -1, 24, 3), full_pos.reshape(1, -1, 3))                # :911-915
    return dict(smpl_trans=full_pos, smpl_poses=full_q, full_pose=pose.reshape(F_, dn, 24, 3).permute(1, 0, 2, 3))
