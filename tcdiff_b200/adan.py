"""Adan optimizer + EMA of the master model, fused over flat arenas (reference model/adan.py:11-123,
model/diffusion.py:61-76, TCDiff.py:110,231-245).

Same constructor, same per-parameter state keys ("step", "prev_grad", "m", "v", "n") and the same update as the
reference class, but the parameters that receive gradients are re-homed (once, at the first `step()`) into one
contiguous fp32 arena per param group, their `.grad`s into a second one, and the whole group is updated by ONE
`tcd_adan_ema_step` launch.  `p.data`, `p.grad` and the state tensors are views into the arenas, so state_dict(),
checkpoints and anything else that walks the parameters keep working.

Data parallel: `Adan(..., data_parallel=True)` (or a process group) averages the gradient arena over the ranks with
NCCL.  From the second step on the all-reduce is bucketed and launched from post-accumulate-grad hooks while the
backward pass is still running (tcdiff_b200.dist.GradReducer); the 1/world factor is applied inside the update kernel.

There is no CPU path: parameters must be CUDA fp32.
"""
import torch
from torch.optim import Optimizer

from . import _lib
from ._lib import check

_ALIGN = 64            # floats; every parameter starts on a 256-byte boundary (TMA / vector-load friendly)


def _stream():
    return torch.cuda.current_stream().cuda_stream


def _bump(tensors):
    """Advance the autograd version counters after a raw-pointer update so caches keyed on them (the packed bf16
    weights of DanceDecoder) notice the change."""
    torch.autograd.graph.increment_version(tensors)


class _Arena:
    """A flat fp32 buffer holding `params` back to back (each aligned to _ALIGN floats)."""

    def __init__(self, shapes, device):
        self.offsets, off = [], 0
        for s in shapes:
            self.offsets.append(off)
            n = 1
            for d in s:
                n *= d
            off += (n + _ALIGN - 1) // _ALIGN * _ALIGN
        self.shapes = list(shapes)
        self.total = off
        self.device = device

    def new_buffer(self):
        return torch.zeros(self.total, dtype=torch.float32, device=self.device)

    def views(self, buf):
        out = []
        for off, s in zip(self.offsets, self.shapes):
            n = 1
            for d in s:
                n *= d
            out.append(buf[off:off + n].view(s))
        return out


def _rehome(tensors_data_owner, views):
    """Copy each parameter's data into its arena view and point the parameter at the view."""
    with torch.no_grad():
        torch._foreach_copy_(views, [p.data for p in tensors_data_owner])
    for p, v in zip(tensors_data_owner, views):
        p.data = v


class Adan(Optimizer):
    def __init__(self, params, lr=1e-3, betas=(0.02, 0.08, 0.01), eps=1e-8, weight_decay=0, restart_cond=None,
                 data_parallel=None, bucket_mb=32):
        assert len(betas) == 3
        if restart_cond is not None:
            raise NotImplementedError("restart_cond is not implemented on the fused path (the reference never passes it, "
                                      "TCDiff.py:110)")
        defaults = dict(lr=lr, betas=betas, eps=eps, weight_decay=weight_decay, restart_cond=restart_cond)
        super().__init__(params, defaults)
        self._flat = {}                    # group index -> dict(arena, live, P, G, PG, M, V, N, step, reducer, ema...)
        self._ema = None                   # (ma_model, cur_model, beta)
        self._dp = data_parallel
        self._bucket_mb = bucket_mb
        _lib.lib()                         # fail now, loudly, if the CUDA library is missing

    # ------------------------------------------------------------------------------------------------ EMA fusion
    def attach_ema(self, ma_model, current_model, beta=0.9999):
        """Fuse `EMA(beta).update_model_average(ma_model, current_model)` (model/diffusion.py:66-71) into every
        `step()`: the averaged copy of each updated parameter is written by the same kernel pass; parameters the
        optimizer does not touch (no gradient) are blended by one extra `tcd_ema_update` launch."""
        self._ema = (ma_model, current_model, float(beta))
        for f in self._flat.values():
            f.pop("ema_ready", None)

    def _setup_ema(self, gi, f):
        ma_model, cur_model, _ = self._ema
        cur = list(cur_model.parameters())
        ma = list(ma_model.parameters())
        if len(cur) != len(ma):
            raise ValueError("attach_ema: the two models have different parameter lists")
        pair = {id(c): m for c, m in zip(cur, ma)}
        live_ma = []
        for p in f["live"]:
            if id(p) not in pair:
                raise ValueError("attach_ema: an optimised parameter does not belong to current_model")
            live_ma.append(pair[id(p)])
        f["E"] = f["arena"].new_buffer()
        _rehome(live_ma, f["arena"].views(f["E"]))
        f["ema_live"] = live_ma
        if gi == 0:
            # everything the optimizer never updates (dead branches, frozen rotary tables): two small arenas
            live_ids = {id(p) for g in self._flat.values() for p in g["live"]}
            rest = [c for c in cur if id(c) not in live_ids]
            if rest:
                ar = _Arena([tuple(c.shape) for c in rest], f["arena"].device)
                f["rest_arena"], f["rest_P"], f["rest_E"] = ar, ar.new_buffer(), ar.new_buffer()
                _rehome(rest, ar.views(f["rest_P"]))
                rest_ma = [pair[id(c)] for c in rest]
                _rehome(rest_ma, ar.views(f["rest_E"]))
                f["rest_ma"] = rest_ma
        f["ema_ready"] = True

    # ------------------------------------------------------------------------------------------------ arenas
    def _build(self, gi, group):
        live = [p for p in group["params"] if p.grad is not None]
        if not live:
            return None
        for p in live:
            if p.device.type != "cuda" or p.dtype != torch.float32:
                raise _lib.TcdError("tcdiff_b200.Adan updates CUDA fp32 parameters only (no CPU path)")
            if p.grad.is_sparse:
                raise RuntimeError("sparse gradients are not supported")
        dev = live[0].device
        arena = _Arena([tuple(p.shape) for p in live], dev)
        f = dict(arena=arena, live=live, step=0)
        for name in ("P", "G", "PG", "M", "V", "N"):
            f[name] = arena.new_buffer()
        f["pviews"] = arena.views(f["P"])
        _rehome(live, f["pviews"])
        gviews = arena.views(f["G"])
        with torch.no_grad():
            torch._foreach_copy_(gviews, [p.grad for p in live])
        for p, g in zip(live, gviews):
            p.grad = g
        f["gviews"] = gviews
        for p, pg, m, v, n in zip(live, arena.views(f["PG"]), arena.views(f["M"]), arena.views(f["V"]),
                                  arena.views(f["N"])):
            st = self.state[p]
            if len(st):                    # state loaded from a checkpoint before the first step
                pg.copy_(st["prev_grad"]); m.copy_(st["m"]); v.copy_(st["v"]); n.copy_(st["n"])
                f["step"] = int(st["step"])
            st.update(step=f["step"], prev_grad=pg, m=m, v=v, n=n)
        world = self._world()
        if world > 1:
            from .dist import GradReducer
            f["reducer"] = GradReducer(live, f["G"], arena.offsets, group=self._group(), bucket_mb=self._bucket_mb)
        self._flat[gi] = f
        return f

    def _group(self):
        return None if self._dp in (None, True, False) else self._dp

    def _world(self):
        if not self._dp:
            return 1
        import torch.distributed as dist
        if not (dist.is_available() and dist.is_initialized()):
            return 1
        return dist.get_world_size(self._group())

    def _grads_in_arena(self, f):
        """Autograd accumulates in place into the arena views; if someone replaced/cleared a .grad (zero_grad with
        set_to_none=True from outside), copy it back and re-point."""
        fix_p, fix_g = [], []
        for p, g in zip(f["live"], f["gviews"]):
            if p.grad is None:
                g.zero_()
                p.grad = g
            elif p.grad.data_ptr() != g.data_ptr():
                fix_p.append(p)
                fix_g.append(g)
        if fix_p:
            with torch.no_grad():
                torch._foreach_copy_(fix_g, [p.grad for p in fix_p])
            for p, g in zip(fix_p, fix_g):
                p.grad = g
            return False
        return True

    def _params_in_arena(self, f):
        """The update kernel writes the arena: a parameter whose .data was replaced since the arenas were built (model.to(),
        .float(), p.data = ...) would silently stop training.  Re-home it (same shape / dtype / device) or raise."""
        moved = [(p, v) for p, v in zip(f["live"], f["pviews"]) if p.data_ptr() != v.data_ptr()]
        if moved and getattr(self, "_capturable", False) and torch.cuda.is_current_stream_capturing():
            raise RuntimeError("tcdiff_b200.Adan: a parameter left the optimizer's arena while its step is being captured")
        for p, v in moved:
            if p.data.shape != v.shape or p.data.dtype != v.dtype or p.data.device != v.device:
                raise RuntimeError("tcdiff_b200.Adan: a parameter's storage was replaced by one of another shape, dtype or "
                                   "device after the first step; rebuild the optimizer (load_state_dict keeps its state)")
        if moved:
            _rehome([p for p, _ in moved], [v for _, v in moved])
        return not moved

    def zero_grad(self, set_to_none=False):
        """One memset per arena; the .grad views stay in place (set_to_none is accepted for API compatibility)."""
        if not self._flat:
            return super().zero_grad(set_to_none=True)
        for f in self._flat.values():
            f["G"].zero_()
            if "reducer" in f:
                f["reducer"].reset()
        flat_ids = {id(p) for f in self._flat.values() for p in f["live"]}
        for group in self.param_groups:
            for p in group["params"]:
                if id(p) not in flat_ids and p.grad is not None:
                    p.grad = None

    # ------------------------------------------------------------------------------------------------ step
    def set_capturable(self, flag=True):
        """Keep the step counter on the device (tcd_adan_ema_step_device) so that `step()` can be captured in a CUDA
        graph and replayed; call `after_replay()` after each replay to keep the host-side bookkeeping in sync."""
        self._capturable = bool(flag)
        for f in self._flat.values():
            f.pop("step_dev", None)

    def reduce_gradients(self, hooks_ran=True):
        """Data parallel: make the gradient arenas hold the SUM over ranks (the 1/world factor is applied by the update
        kernel).  Called by step(); separately by GraphedTrainStep between its two graphs."""
        for f in self._flat.values():
            if "reducer" in f:
                f["reducer"].finish(all_in_hooks=hooks_ran)

    def after_replay(self):
        """Host-side bookkeeping of one replayed (graph-captured) step: step counters and version counters."""
        for f in self._flat.values():
            f["step"] += 1
            for p in f["live"]:
                self.state[p]["step"] = f["step"]
            _bump(f["live"])
            if self._ema is not None and f.get("ema_ready"):
                _bump(f["ema_live"])
                if "rest_ma" in f:
                    _bump(f["rest_ma"])

    @torch.no_grad()
    def step(self, closure=None, reduce=True):
        loss = None
        if closure is not None:
            with torch.enable_grad():
                loss = closure()
        lib = _lib.lib()
        for gi, group in enumerate(self.param_groups):
            f = self._flat.get(gi)
            first = f is None
            if first:
                f = self._build(gi, group)
                if f is None:
                    continue
            if not first and len(f["live"]) != len(group["params"]):
                live_ids = f.setdefault("live_ids", {id(p) for p in f["live"]})
                late = [p for p in group["params"] if p.grad is not None and id(p) not in live_ids]
                if late:
                    raise RuntimeError("tcdiff_b200.Adan: %d parameter(s) received a gradient for the first time after the "
                                       "arenas were built (the set of trained parameters is fixed at the first step)" % len(late))
            self._params_in_arena(f)
            world = 1
            if "reducer" in f:                       # its hooks keep the gradients in the arena
                world = f["reducer"].world
                if reduce:
                    f["reducer"].finish(all_in_hooks=not first)
            else:
                self._grads_in_arena(f)
            ema_ptr, ema_beta = 0, 0.0
            if self._ema is not None:
                if not f.get("ema_ready"):
                    self._setup_ema(gi, f)
                ema_ptr, ema_beta = f["E"].data_ptr(), self._ema[2]
            b1, b2, b3 = group["betas"]
            ptrs = (f["P"].data_ptr(), f["G"].data_ptr(), f["PG"].data_ptr(), f["M"].data_ptr(), f["V"].data_ptr(),
                    f["N"].data_ptr(), ema_ptr, f["arena"].total)
            hyper = (1.0 / world, group["lr"], b1, b2, b3, group["eps"], group["weight_decay"], ema_beta, _stream())
            if getattr(self, "_capturable", False):
                if "step_dev" not in f:
                    f["step_dev"] = torch.tensor([f["step"]], dtype=torch.int64, device=f["P"].device)
                    f["scalars"] = torch.zeros(32, dtype=torch.float32, device=f["P"].device)
                check(lib.tcd_adan_ema_step_device(*ptrs, f["step_dev"].data_ptr(), f["scalars"].data_ptr(), *hyper))
            else:
                check(lib.tcd_adan_ema_step(*ptrs, f["step"], *hyper))
            f["step"] += 1
            for p in f["live"]:
                self.state[p]["step"] = f["step"]
            _bump(f["live"])
            if self._ema is not None:
                _bump(f["ema_live"])
                if "rest_E" in f:
                    check(lib.tcd_ema_update(f["rest_E"].data_ptr(), f["rest_P"].data_ptr(), f["rest_arena"].total,
                                             ema_beta, _stream()))
                    _bump(f["rest_ma"])
        return loss

    def load_state_dict(self, state_dict):
        """Loaded state tensors are copied into the arenas (or kept until the arenas are built at the first step)."""
        flat, self._flat = self._flat, {}
        super().load_state_dict(state_dict)
        for gi, f in flat.items():
            for p, pg, m, v, n in zip(f["live"], f["arena"].views(f["PG"]), f["arena"].views(f["M"]),
                                      f["arena"].views(f["V"]), f["arena"].views(f["N"])):
                st = self.state[p]
                if len(st):
                    pg.copy_(st["prev_grad"]); m.copy_(st["m"]); v.copy_(st["v"]); n.copy_(st["n"])
                    f["step"] = int(st["step"])
                st.update(step=f["step"], prev_grad=pg, m=m, v=v, n=n)
        self._flat = flat
