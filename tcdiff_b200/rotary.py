"""Drop-in RotaryEmbedding (reference: model/rotary_embedding_torch.py:75-130), 'lang' frequencies only.

freqs_i = theta^(-2i/dim); a token at position p has its feature pairs (2i, 2i+1) rotated by p*freqs_i.
TCDiff applies this to the whole 512-d token vector BEFORE the Q/K projections, with p = flat token index
(model/model.py:375,387-388).  On the denoiser path the rotation is fused into the LayerNorm kernels;
this class keeps the reference's public method for callers that use it directly.
"""
import torch
from torch import nn

from . import ops


class RotaryEmbedding(nn.Module):
    def __init__(self, dim, custom_freqs=None, freqs_for="lang", theta=10000, max_freq=10, num_freqs=1,
                 learned_freq=False):
        super().__init__()
        if freqs_for != "lang" or learned_freq:
            raise NotImplementedError("only freqs_for='lang', learned_freq=False is used by TCDiff")
        if custom_freqs is not None:
            freqs = custom_freqs
        else:
            freqs = 1.0 / (theta ** (torch.arange(0, dim, 2)[: dim // 2].float() / dim))
        self.register_buffer("freqs", freqs)
        self.cache = dict()

    def tables(self, seq_len, device):
        """(cos, sin) of shape (seq_len, dim/2): fp32 positions x fp32 freqs evaluated on the host, exactly
        the reference's einsum-then-cos (rotary_embedding_torch.py:115-130), then uploaded."""
        key = (seq_len, str(device))
        if key not in self.cache:
            f = self.freqs.detach().float().cpu()
            ang = torch.arange(seq_len).type(f.dtype)[:, None] * f[None, :]
            self.cache[key] = (ang.cos().to(device).contiguous(), ang.sin().to(device).contiguous())
        return self.cache[key]

    def rotate_queries_or_keys(self, t, seq_dim=-2):
        if seq_dim not in (-2, t.dim() - 2):
            raise NotImplementedError("seq_dim must be -2")
        L, D = t.shape[-2], t.shape[-1]
        cos, sin = self.tables(L, t.device)
        x = t.float().contiguous()
        out = torch.empty_like(x)
        ops.rotary(x, out, cos, sin, x.numel() // D, D, L)
        return out.to(t.dtype)

    def forward(self, t, cache_key=None):
        """Angle table (…, dim) with every frequency repeated twice, as the reference returns."""
        if callable(t):
            t = t()
        ang = t.type(self.freqs.dtype)[..., None] * self.freqs
        return ang.repeat_interleave(2, dim=-1)
