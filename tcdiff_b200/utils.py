"""Host-side scalar machinery of the diffusion wrapper: noise schedules, per-timestep coefficient lookup, the
classifier-free-guidance keep mask and the module form of the timestep embedding.

Behavioural contract = the reference's model/utils.py:36-99.  What matters for parity is the PRECISION ORDER of the
schedule (float64 all the way to the betas, fp32 only afterwards — GaussianDiffusion's 13 buffers are bit-identical to the
reference's, tests/test_abi_and_host.py) and that the degenerate keep-probabilities 0 and 1 consume no random numbers.
"""
import math

import numpy as np
import torch
from torch import nn

_F64 = torch.float64


def _cosine_betas(n, s):
    # alphas_cumprod(t) = cos^2(((t/n + s) / (1 + s)) * pi/2), normalised to 1 at t = 0; beta_t = 1 - acp_t / acp_{t-1}
    steps = torch.arange(n + 1, dtype=_F64) / n + s
    acp = torch.cos(steps / (1 + s) * np.pi / 2).pow(2)
    acp = acp / acp[0]
    return np.clip(1 - acp[1:] / acp[:-1], a_min=0, a_max=0.999)


_SCHEDULES = {
    "linear": lambda n, lo, hi, s: torch.linspace(lo ** 0.5, hi ** 0.5, n, dtype=_F64) ** 2,
    "cosine": lambda n, lo, hi, s: _cosine_betas(n, s),
    "sqrt_linear": lambda n, lo, hi, s: torch.linspace(lo, hi, n, dtype=_F64),
    "sqrt": lambda n, lo, hi, s: torch.linspace(lo, hi, n, dtype=_F64) ** 0.5,
}


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """numpy float64 betas of length n_timestep for one of the four schedules the reference knows."""
    try:
        build = _SCHEDULES[schedule]
    except KeyError:
        raise ValueError(f"schedule '{schedule}' unknown.") from None
    betas = build(n_timestep, linear_start, linear_end, cosine_s)
    return betas.numpy() if torch.is_tensor(betas) else np.asarray(betas)


def extract(a, t, x_shape):
    """Per-sample coefficient a[t_b], shaped (B, 1, ..., 1) to broadcast against a tensor of shape x_shape."""
    picked = a.gather(-1, t)
    return picked.reshape((t.shape[0],) + (1,) * (len(x_shape) - 1))


def prob_mask_like(shape, prob, device):
    """Boolean mask that is True with probability `prob`; prob 0 / 1 are answered without touching the RNG."""
    if prob in (0, 1):
        return torch.full(shape, bool(prob), device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


class SinusoidalPosEmb(nn.Module):
    """(B,) timesteps -> (B, dim): [sin(t e_j) | cos(t e_j)], e_j = exp(-j ln(1e4) / (dim/2 - 1)).  The denoiser itself
    gathers rows of a host-built (1000, dim) table (engine.PackedWeights.time_table) holding the same expression."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half = self.dim // 2
        freq = torch.exp(torch.arange(half, device=x.device) * -(math.log(10000) / (half - 1)))
        ang = x[:, None] * freq[None, :]
        return torch.cat((ang.sin(), ang.cos()), dim=-1)
