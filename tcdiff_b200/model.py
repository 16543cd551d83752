"""Drop-in DanceDecoder (reference: model/model.py:416-624) running on the sm_100a kernels.

The module owns the same fp32 ``nn.Parameter`` set under the same names as the reference (446
state_dict entries at the TCDiff.py:76-87 hyper-parameters, including the 11 aliased
``*.rotary.freqs`` buffers and the dead ``traj_Modulation`` / ``traj_embedding`` /
``embeddings_table`` weights), so reference checkpoints load key-for-key (``module.``-prefixed too).
The parameters stay the source of truth: optimizers/EMA mutate them in place, and the kernel-side
packed copies (bf16 / padded / concatenated) are a derived cache that is rebuilt whenever a parameter's
(data_ptr, version) signature changes and that does not survive ``copy.deepcopy``
(GaussianDiffusion deep-copies the model into ``master_model``, model/diffusion.py:101).

Behaviour differences from the reference, by design:
  * .eval() calls (the sampling path) use the graph-free inference engine; train-mode calls with gradients enabled
    run on the autograd tape of kernel calls (tcdiff_b200/train.py); dropout > 0 needs the bf16 tape;
  * ``trj_dist`` is unsupported (it is never passed by the reference and fails there);
  * no CPU execution: inputs must be CUDA tensors and the CUDA library must be present.
"""
import os
from typing import Callable

import torch
import torch.nn as nn
import torch.nn.functional as F
from torch import Tensor

from . import ops
from .engine import Denoiser, PackedWeights, Workspace
from .rotary import RotaryEmbedding
from .utils import SinusoidalPosEmb

_DTYPES = {"fp32": torch.float32, "float32": torch.float32, "bf16": torch.bfloat16, "bfloat16": torch.bfloat16,
           torch.float32: torch.float32, torch.bfloat16: torch.bfloat16}


class _Cache:
    """Derived kernel-side state; deliberately empty after deepcopy / pickling."""

    def __init__(self):
        self.sig = None
        self.packed = None
        self.denoiser = None
        self.ws = None

    def __deepcopy__(self, memo):
        return _Cache()

    def __getstate__(self):
        return {}

    def __setstate__(self, state):
        self.__init__()


def _gated(d_in, d_out, d_ctx):
    # parameter container for the reference's ConcatSquashLinear (dead code w.r.t. the output)
    m = nn.Module()
    m._layer = nn.Linear(d_in, d_out)
    m._hyper_bias = nn.Linear(d_ctx, d_out, bias=False)
    m._hyper_gate = nn.Linear(d_ctx, d_out)
    return m


def _msa(n_head, d_model):
    m = nn.Module()
    for name in ("w_qs", "w_ks", "w_vs"):
        setattr(m, name, nn.Linear(d_model, n_head * 64, bias=False))
    m.fc = nn.Linear(n_head * 64, d_model, bias=False)
    m.layer_norm = nn.LayerNorm(d_model, eps=1e-6)
    return m


def _film(d):
    m = nn.Module()
    m.block = nn.Sequential(nn.Mish(), nn.Linear(d, 2 * d))
    return m


def _encoder_layer(d, heads, ff, rotary):
    m = nn.Module()
    m.self_attn = nn.MultiheadAttention(d, heads, batch_first=True)   # parameter container only
    m.linear1, m.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
    m.norm1, m.norm2 = nn.LayerNorm(d), nn.LayerNorm(d)
    m.rotary = rotary
    return m


def _decoder_layer(d, heads, ff, rotary):
    m = nn.Module()
    m.self_attn, m.multihead_attn = _msa(heads, d), _msa(heads, d)
    m.linear1, m.linear2 = nn.Linear(d, ff), nn.Linear(ff, d)
    m.norm1, m.norm2, m.norm3 = nn.LayerNorm(d), nn.LayerNorm(d), nn.LayerNorm(d)
    m.film1, m.film2, m.film3 = _film(d), _film(d), _film(d)
    m.rotary = rotary
    m.linear3, m.norm4 = nn.Linear(d, d), nn.LayerNorm(d)
    m.traj_Modulation = nn.ModuleList([_gated(d, 128, 512), _gated(128, 128, 512), _gated(128, d, 512)])
    return m


class DanceDecoder(nn.Module):
    def __init__(self, nfeats: int, seq_len: int = 150, latent_dim: int = 256, ff_size: int = 1024,
                 num_layers: int = 4, num_heads: int = 4, dropout: float = 0.1, cond_feature_dim: int = 4800,
                 activation: Callable[[Tensor], Tensor] = F.gelu, use_rotary=True, required_dancer_num=4,
                 **kwargs) -> None:
        super().__init__()
        if not use_rotary:
            raise NotImplementedError("use_rotary=False (absolute positional encoding) is not on the supported path")
        if activation not in (F.gelu, "gelu"):
            raise NotImplementedError("only the reference's activation=F.gelu is implemented (TCDiff.py:85)")
        if nfeats != 151:
            raise NotImplementedError("the reference hard-codes 151 motion channels (model/model.py:553)")
        if latent_dim % 128 or latent_dim // num_heads != 64:
            raise NotImplementedError("kernels need latent_dim % 128 == 0 and head dim 64 (reference: 512 / 8)")
        self.nfeats, self.latent_dim, self.seq_len = nfeats, latent_dim, seq_len
        self.ff_size, self.num_layers, self.num_heads = ff_size, num_layers, num_heads
        self.required_dancer_num = required_dancer_num
        self.cond_feature_dim = cond_feature_dim
        self.dropout_p = dropout
        self.compute_dtype = _DTYPES[kwargs.pop("dtype", os.environ.get("TCDIFF_DTYPE", "bf16"))]
        D = latent_dim
        self.rotary = RotaryEmbedding(dim=D)
        self.abs_pos_encoding = nn.Identity()
        self.time_mlp = nn.Sequential(SinusoidalPosEmb(D), nn.Linear(D, D * 4), nn.Mish())
        self.to_time_cond = nn.Sequential(nn.Linear(D * 4, D))
        self.to_time_tokens = nn.Sequential(nn.Linear(D * 4, D * 2), nn.Identity())
        self.null_cond_embed = nn.Parameter(torch.randn(1, seq_len, D))
        self.null_cond_hidden = nn.Parameter(torch.randn(1, D))
        self.norm_cond = nn.LayerNorm(D)
        self.input_projection = nn.Linear(nfeats, D)
        self.cond_encoder = nn.Sequential(*[_encoder_layer(D, num_heads, ff_size, self.rotary) for _ in range(2)])
        self.cond_projection = nn.Sequential(nn.Linear(cond_feature_dim * 2, cond_feature_dim), nn.ReLU(),
                                             nn.Linear(cond_feature_dim, D))
        self.non_attn_cond_projection = nn.Sequential(nn.LayerNorm(D), nn.Linear(D, D), nn.SiLU(), nn.Linear(D, D))
        stack = nn.Module()
        stack.stack = nn.ModuleList([_decoder_layer(D, num_heads, ff_size, self.rotary) for _ in range(num_layers)])
        self.seqTransDecoder = stack
        self.final_layer = nn.Linear(D, nfeats)
        self.relative_projection_layer = nn.Sequential(nn.Linear(D * required_dancer_num, D * 2), nn.ReLU(),
                                                       nn.Linear(D * 2, D * 2), nn.ReLU(),
                                                       nn.Linear(D * 2, D * required_dancer_num))
        self.d_k = 64
        self.embeddings_table = nn.Embedding(10, self.d_k * num_heads)
        self.traj_embedding = nn.Sequential(nn.Linear(2, 64), nn.ReLU(), nn.Linear(64, D))
        self._cache = _Cache()
        # rows of the host-built timestep-embedding table (SinusoidalPosEmb of every integer timestep, model/utils.py:41-48);
        # GaussianDiffusion raises it to its n_timestep
        self.time_table_rows = 1000

    def set_time_table_rows(self, n):
        if int(n) > self.time_table_rows:
            self.time_table_rows = int(n)
            self._cache.sig = None
            self._cache.__dict__.pop("train_tables", None)

    # ------------------------------------------------------------------ derived kernel-side state
    def kernel_config(self):
        return dict(nfeats=self.nfeats, latent_dim=self.latent_dim, dancers=self.required_dancer_num,
                    cond_feature_dim=self.cond_feature_dim, seq_len=self.seq_len, num_layers=self.num_layers,
                    num_heads=self.num_heads, ff_size=self.ff_size)

    def _signature(self):
        sig = [self.compute_dtype]
        for p in self.parameters():
            sig.append((p.data_ptr(), p._version))
        return tuple(sig)

    def denoiser(self):
        """(Denoiser, Workspace) for the current parameters; repacks when they changed."""
        c = self._cache
        sig = self._signature()
        if c.sig != sig:
            dev = self.input_projection.weight.device
            if dev.type != "cuda":
                raise ops._lib.TcdError("tcdiff_b200.DanceDecoder runs on CUDA only; move the module with .cuda()")
            sd = {k: v for k, v in self.state_dict().items()}
            c.packed = PackedWeights(sd, self.kernel_config(), self.compute_dtype, dev, n_timestep=self.time_table_rows)
            c.denoiser = Denoiser(c.packed)
            if c.ws is None or c.ws.device != dev:
                c.ws = Workspace(dev)
            c.sig = sig
        return c.denoiser, c.ws

    def mark_weights_dirty(self):
        """Force a repack on next use (needed only after raw ``param.data`` in-place edits, which do not
        advance the version counters the cache signature watches)."""
        self._cache.sig = None

    def set_compute_dtype(self, dtype):
        self.compute_dtype = _DTYPES[dtype]
        return self

    def load_state_dict(self, state_dict, strict=True, **kw):
        # accept DDP-wrapped checkpoints ("module." prefix, TCDiff.py:31-36)
        if state_dict and all(k.startswith("module.") for k in state_dict):
            state_dict = {k[len("module."):]: v for k, v in state_dict.items()}
        return super().load_state_dict(state_dict, strict=strict, **kw)

    # ------------------------------------------------------------------ reference API
    def _keep_mask(self, batch, cond_drop_prob, device, keep_mask):
        if keep_mask is not None:
            return keep_mask.to(device=device, dtype=torch.uint8)
        if cond_drop_prob == 0:                      # prob_mask_like(.., 1 - p): p == 0 keeps all
            return torch.ones(batch, dtype=torch.uint8, device=device)
        if cond_drop_prob == 1:
            return torch.zeros(batch, dtype=torch.uint8, device=device)
        return (torch.zeros(batch, device=device).float().uniform_(0, 1) < (1 - cond_drop_prob)).to(torch.uint8)

    def forward(self, x: Tensor, cond_embed: Tensor, times: Tensor, cond_drop_prob: float = 0.0, trj_dist=None, *,
                keep_mask=None):
        if trj_dist is not None:
            raise NotImplementedError("trj_dist is unsupported (it fails in the reference too, SURVEY §8b)")
        batch = x.shape[0]
        x = x.reshape(batch, -1, 151)
        if x.shape[1] != self.seq_len * self.required_dancer_num:
            raise ValueError(f"expected {self.seq_len * self.required_dancer_num} tokens, got {x.shape[1]}")
        if self.training and self.dropout_p > 0 and (self.compute_dtype != torch.bfloat16 or not torch.is_grad_enabled()):
            raise NotImplementedError("training-mode dropout runs on the bf16 autograd tape only; call .eval(), use "
                                      "dtype='bf16' with gradients enabled, or build the model with dropout=0.0")
        if x.device.type != "cuda":
            raise ops._lib.TcdError("tcdiff_b200.DanceDecoder runs on CUDA only; move the inputs with .cuda()")
        x = x.to(torch.float32).contiguous()
        keep = self._keep_mask(batch, cond_drop_prob, x.device, keep_mask)
        times = times.to(device=x.device, dtype=torch.int64).contiguous()
        if self.training and torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in self.parameters())):
            # differentiable call (train mode only; .eval() is always the graph-free inference engine):
            # the autograd tape of kernel calls (tcdiff_b200/train.py)
            from . import train
            return train.denoiser_forward_train(self, x, cond_embed.to(x.device), times, keep)
        with torch.no_grad():
            den, ws = self.denoiser()
            return den.forward(ws, x, cond_embed.to(x.device), times, keep)

    @torch.no_grad()
    def guided_forward(self, x, cond_embed, times, guidance_weight):
        unc = self.forward(x, cond_embed, times, cond_drop_prob=1)
        con = self.forward(x, cond_embed, times, cond_drop_prob=0)
        out = torch.empty_like(con)
        # unc + (con - unc) * w through the fused step kernel's blend (no clamp, x' = x0)
        ops.cfg_ddim_step(con, con, unc, None, None, out, None, None, 0, con.numel() // 151, float(guidance_weight),
                          0.0, 1.0, 0.0, 0.0, 0.0, clip=False, last=True)
        return out
