"""Host-side helpers mirroring model/utils.py (schedule construction, coefficient lookup, CFG mask)."""
import math

import numpy as np
import torch
from torch import nn


class SinusoidalPosEmb(nn.Module):
    """Timestep embedding (reference model/utils.py:36-48).  On the denoiser path this is a gather from a
    host-built table (engine.PackedWeights.time_table); the module form evaluates the same expression."""

    def __init__(self, dim):
        super().__init__()
        self.dim = dim

    def forward(self, x):
        half = self.dim // 2
        k = math.log(10000) / (half - 1)
        f = torch.exp(torch.arange(half, device=x.device) * -k)
        a = x[:, None] * f[None, :]
        return torch.cat((a.sin(), a.cos()), dim=-1)


def prob_mask_like(shape, prob, device):
    """Bernoulli keep-mask with the degenerate cases kept RNG-free (reference model/utils.py:52-58)."""
    if prob == 1:
        return torch.ones(shape, device=device, dtype=torch.bool)
    if prob == 0:
        return torch.zeros(shape, device=device, dtype=torch.bool)
    return torch.zeros(shape, device=device).float().uniform_(0, 1) < prob


def extract(a, t, x_shape):
    """a[t] broadcastable against x_shape (reference model/utils.py:61-64)."""
    return a.gather(-1, t).reshape(t.shape[0], *((1,) * (len(x_shape) - 1)))


def make_beta_schedule(schedule, n_timestep, linear_start=1e-4, linear_end=2e-2, cosine_s=8e-3):
    """float64 numpy betas (reference model/utils.py:67-99); precision order matters for parity."""
    f64 = torch.float64
    if schedule == "linear":
        betas = torch.linspace(linear_start ** 0.5, linear_end ** 0.5, n_timestep, dtype=f64) ** 2
    elif schedule == "cosine":
        grid = torch.arange(n_timestep + 1, dtype=f64) / n_timestep + cosine_s
        acp = torch.cos(grid / (1 + cosine_s) * np.pi / 2).pow(2)
        acp = acp / acp[0]
        betas = np.clip(1 - acp[1:] / acp[:-1], a_min=0, a_max=0.999)
    elif schedule == "sqrt_linear":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=f64)
    elif schedule == "sqrt":
        betas = torch.linspace(linear_start, linear_end, n_timestep, dtype=f64) ** 0.5
    else:
        raise ValueError(f"schedule '{schedule}' unknown.")
    return betas.numpy() if isinstance(betas, torch.Tensor) else np.asarray(betas)
