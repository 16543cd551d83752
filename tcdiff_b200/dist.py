"""Batch-sharded sampling across GPUs (one process per GPU, torch.distributed for the plumbing).

Every batch row of ddim_sample / p_sample_loop is independent (no cross-sample op anywhere in the
denoiser), so rank r takes a contiguous slice of the batch, runs its own CUDA graph with no data-path
collective, and ONE all_gather at the end assembles the result.  `long_ddim_sample`
couples neighbouring rows and must not be sharded below song granularity.
"""
import torch
import torch.distributed as dist


def shard_bounds(batch, world, rank):
    """Contiguous, balanced [lo, hi) slice of `batch` rows for `rank` (first `batch % world` ranks get +1)."""
    base, extra = divmod(batch, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def sharded_sample(sample_fn, shape, cond, x_0=None, noise_bank=None, group=None, **kw):
    """Run `sample_fn(shape_local, cond_local, x_0=..., noise_bank=...)` on this rank's rows and gather.

    sample_fn is GaussianDiffusion.ddim_sample (or p_sample_loop wrapped to the same signature).
    Inputs are the FULL-batch tensors (every rank holds them, e.g. from a replicated data loader);
    returns the full (B, L, C) result on every rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    B = shape[0]
    lo, hi = shard_bounds(B, world, rank)
    local_bank = None if noise_bank is None else [n[lo:hi] for n in noise_bank]
    out = sample_fn((hi - lo,) + tuple(shape[1:]), cond[lo:hi], x_0=None if x_0 is None else x_0[lo:hi],
                    noise_bank=local_bank, **kw)
    if world == 1:
        return out
    sizes = [shard_bounds(B, world, r) for r in range(world)]
    maxn = max(h - l for l, h in sizes)
    pad = out.new_zeros((maxn,) + tuple(out.shape[1:]))
    pad[: hi - lo] = out
    parts = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(parts, pad, group=group)                 # the single collective of the sampling path
    return torch.cat([p[: h - l] for p, (l, h) in zip(parts, sizes)], dim=0)
