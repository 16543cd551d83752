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
    if all(h - l == maxn for l, h in sizes):                  # even shards: gather straight into the result, no copies
        full = out.new_empty((B,) + tuple(out.shape[1:]))
        dist.all_gather_into_tensor(full, out.contiguous(), group=group)     # the single collective of the sampling path
        return full
    pad = out.new_zeros((maxn,) + tuple(out.shape[1:]))
    pad[: hi - lo] = out
    full = out.new_empty((world * maxn,) + tuple(out.shape[1:]))
    dist.all_gather_into_tensor(full, pad, group=group)
    return torch.cat([full[r * maxn: r * maxn + (h - l)] for r, (l, h) in enumerate(sizes)], dim=0)


class GradReducer:
    """Data-parallel gradient averaging over a flat gradient arena (the reference gets this from
    accelerate/DDP with find_unused_parameters=True, TCDiff.py:51,108).

    The arena is cut into contiguous buckets of ~bucket_mb; a post-accumulate-grad hook on every parameter counts its
    bucket down and, when the bucket is complete, launches an asynchronous SUM all-reduce of that slice while the
    backward pass keeps running (NCCL runs on its own stream).  `finish()` waits for the outstanding buckets (and
    reduces synchronously whatever the hooks did not cover); the division by the world size is left to the consumer
    (the optimizer kernel multiplies by 1/world).  One backward pass per step (gradient accumulation over several
    backward passes would need the hooks disabled: `enabled = False`, then `finish(all_in_hooks=False)`).  Device
    agnostic: the same logic runs over gloo on CPU tensors.
    """

    def __init__(self, params, grad_arena, offsets, group=None, bucket_mb=32):
        self.group = group
        self.views = {id(p): grad_arena[o:o + p.numel()].view(p.shape) for p, o in zip(params, offsets)}
        self.world = dist.get_world_size(group)
        self.arena = grad_arena
        limit = max(1, int(bucket_mb * (1 << 20)) // 4)
        # buckets in REVERSE parameter order (gradients arrive roughly back to front)
        self.buckets = []            # [lo, hi, n_params]
        self.bucket_of = {}
        ends = list(offsets[1:]) + [grad_arena.numel()]
        hi = None
        for i in range(len(params) - 1, -1, -1):
            if hi is None:
                hi, cnt = ends[i], 0
            cnt += 1
            self.bucket_of[id(params[i])] = len(self.buckets)
            if hi - offsets[i] >= limit or i == 0:
                self.buckets.append([offsets[i], hi, cnt])
                hi = None
        self.pending = [b[2] for b in self.buckets]
        self.works = [None] * len(self.buckets)
        self.launched = 0
        self.enabled = True          # False: hooks only keep the gradients in the arena (CUDA-graph capture)
        self._handles = [p.register_post_accumulate_grad_hook(self._hook) for p in params]

    def _hook(self, p):
        v = self.views[id(p)]
        if p.grad.data_ptr() != v.data_ptr():        # .grad was cleared/replaced from outside: bring it home
            v.copy_(p.grad)
            p.grad = v
        if not self.enabled:
            return
        b = self.bucket_of[id(p)]
        self.pending[b] -= 1
        if self.pending[b] == 0:
            lo, hi, _ = self.buckets[b]
            self.works[b] = dist.all_reduce(self.arena[lo:hi], op=dist.ReduceOp.SUM, group=self.group, async_op=True)
            self.launched += 1

    def reset(self):
        self.pending = [b[2] for b in self.buckets]
        self.works = [None] * len(self.buckets)

    def finish(self, all_in_hooks=True):
        """Make the arena hold the SUM over ranks.  all_in_hooks=False: the backward pass ran before the hooks
        existed (the step that builds the arena) -> one synchronous all-reduce of the whole arena."""
        if not all_in_hooks:
            dist.all_reduce(self.arena, op=dist.ReduceOp.SUM, group=self.group)
            self.reset()
            return
        for b, w in enumerate(self.works):
            if w is not None:
                w.wait()
            else:                    # bucket with a parameter that got no gradient this step
                lo, hi, _ = self.buckets[b]
                dist.all_reduce(self.arena[lo:hi], op=dist.ReduceOp.SUM, group=self.group)
        self.reset()

    def remove(self):
        for h in self._handles:
            h.remove()
