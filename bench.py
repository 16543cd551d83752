#!/usr/bin/env python
"""bench.py — TCDiff denoising hot path on B200: 5 s group-dance clips/s with DDIM-50.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B] [--dtype bf16|fp32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One bench "step" = one full DDIM-50 sampling call (50 guided network evaluations + 50 fused update
kernels, prologue included) over one batch of synthetic AIOZ-GDance-shaped input per GPU:
BASELINE.json configs[1] — batch 64, 5 dancers, 150 frames, 151-dim motion, 438-dim music features,
dataset-style trajectory conditioning, bf16 tensor cores, CUDA-graph captured loop, random-init weights.
N > 1 is weak scaling (batch 64 per GPU, batch-sharded, one final all_gather inside the timed region).

Prints ONE JSON line (rank 0).  Keys beyond the base contract: roofline (dominant kernel = the tcgen05
GEMM, timed live with CUDA events in an instrumented eager denoise step), cpu_baseline (the oracle port on
the host cores, bounded sample), e2e (public API with pinned host buffers, H2D + D2H inside the timing).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "clips_per_sec_ddim50"
UNIT = "clips/s"
WORKLOAD = "c2: DDIM-50, batch 64/GPU, 5 dancers, 150 frames, 151-d motion, 438-d music, traj in-painting, CFG w=2"


def peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        d = json.load(open(p))
        return dict(tflops=float(d.get("bf16_tflops_sustained", d.get("bf16_tflops", 1400.0))),
                    tflops_burst=float(d.get("bf16_tflops", 1590.0)), hbm=float(d.get("hbm_gbs", 6650.0)), src="measured")
    return dict(tflops=1400.0, tflops_burst=1590.0, hbm=6650.0, src="fallback")


def flops_per_pass(cfg, hoisted=True):
    """Live MACs*2 of one network pass per sample; hoisted=True drops the step-invariant part."""
    D, FF, NL = cfg["latent_dim"], cfg["ff_size"], cfg["num_layers"]
    S, dn, Fm = cfg["seq_len"], cfg["dancers"], cfg["cond_feature_dim"]
    L, M = S * dn, S + 2
    front = L * 151 * D + S * (dn * D * 2 * D + 2 * D * 2 * D + 2 * D * dn * D)
    music = S * (2 * Fm * Fm + Fm * D) + 2 * (3 * S * D * D + 2 * S * S * D + S * D * D + 2 * S * D * FF)
    layer = (3 * L * D * D + 2 * L * L * D + L * D * D) + (L * D * D + 2 * M * D * D + 2 * L * M * D + L * D * D) \
        + 2 * L * D * FF + L * D * D
    head = L * D * 151
    hoist = music + NL * 2 * S * D * D
    total = front + music + NL * layer + head
    return 2 * (total - hoist if hoisted else total)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled while the timed region runs."""

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def __enter__(self):
        q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={q}", "--format=csv,noheader,nounits",
                                          "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if len(r) >= 7 and r[0].replace(".", "").isdigit())
        if not sm:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) >= 7 and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": float(self.rows[0][1]), "reasons": reasons,
                "samples": len(sm), "power_w_max": max(float(r[2]) for r in self.rows if len(r) >= 7)}


def build_ours(cfg_name, dtype, dev):
    import tcdiff_b200 as T
    from tcdiff_b200 import synth        # synthetic weights/inputs only (not the checker)
    cfg = synth.CONFIGS[cfg_name]
    m = T.DanceDecoder(nfeats=151, seq_len=cfg["seq_len"], latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"],
                       num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], dropout=0.1,
                       cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"], dtype=dtype)
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, cfg["seq_len"], 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000,
                            predict_epsilon=False, loss_type="l2", use_p2=False, cond_drop_prob=0.25,
                            guidance_weight=2).to(dev).eval()
    return cfg, m, d


def kernel_breakdown(d, m, B, cfg):
    """Instrumented eager denoise step: CUDA-event time and algorithmic FLOPs per kernel class."""
    from tcdiff_b200 import ops
    den, _ = m.denoiser()
    ent = next(e for e in d._graphs.values() if "bufs" in e)          # the DDIM entry of the timed configuration
    ws, bufs = ent["ws"], ent["bufs"]
    rec = {}
    orig = {}

    def wrap(name, flops_fn):
        f = getattr(ops, name)
        orig[name] = f

        def g(*a, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            r = f(*a, **k)
            e1.record()
            rec.setdefault(name, []).append((e0, e1, flops_fn(*a, **k)))
            return r
        setattr(ops, name, g)

    shapes = []

    def gemm_flops(a, w, bias, act, out, M=None, N=None, K=None, **k):
        M = a.shape[0] if M is None else M
        N = w.shape[0] if N is None else N
        K = a.shape[1] if K is None else K
        shapes.append((M, N, K, act, str(out.dtype).replace("torch.", "")))
        return 2.0 * M * N * min(K, w.shape[1])

    def attn_flops(q, ldq, qbs, k, ldk, kbs, v, ldv, vbs, o, ldo, obs, samples, heads, Lq, Lk, scale, **kw):
        return 4.0 * samples * heads * Lq * Lk * 64

    # HBM-bound kernels: ALGORITHMIC bytes per launch, returned negative to tell them from FLOPs
    def frn_bytes(op_dtype, x_in, x_out, y, ln_in, eps_in, film, film_ld, film_off, ln_next, eps_next, out_plain, out_rot,
                  rot_cos, rot_sin, rows, D, tps):
        per = 4 + (0 if x_out is None else 4) + y.element_size()     # x read + x written (unless dead) + y read
        for o in (out_plain, out_rot):
            per += 0 if o is None else o.element_size()
        return -float(rows * D * per)

    def ln_bytes(x, gamma, beta, eps, out_plain, out_rot, rot_cos, rot_sin, rows, D, tps):
        per = 4 + sum(0 if o is None else o.element_size() for o in (out_plain, out_rot))
        return -float(rows * D * per)

    def step_bytes(x, out_cond, out_uncond, noise, traj, x_out, x0_out, xpad, xpad_ld, n_tokens, *a, **k):
        # x, cond, uncond read + x written = 16 B/elt, + 4 when the draw is read from a noise tensor (the default
        # path generates it in the kernel), + 2 for the bf16 operand copy, + 8 B/token of trajectory
        per = 151 * (16 + (4 if noise is not None else 0) + (2 if xpad is not None else 0)) + (8 if traj is not None else 0)
        return -float(n_tokens * per)

    fused_bytes = []

    def fused_flops(a, w, bias, x_in, x_out, ln_in, eps_in, film, film_ld, film_off, ln_next, eps_next, out_plain, out_rot,
                    rot_cos_t, rot_sin_t, rows, tps):
        # the contraction's FLOPs; the kernel is HBM-bound by design, its algorithmic bytes are reported beside them
        per = a.shape[1] * 2 + 512 * ((0 if x_in is None else 4) + (0 if x_out is None else 4) +
                                      sum(0 if o is None else 2 for o in (out_plain, out_rot)))
        fused_bytes.append(float(rows * per))
        return 2.0 * rows * 512 * a.shape[1]

    for n, fn in (("gemm", gemm_flops), ("attention", attn_flops), ("film_residual_norm", frn_bytes),
                  ("gemm_film_residual_norm", fused_flops),
                  ("layernorm_rotary", ln_bytes), ("scatter_rows", lambda *a, **k: 0.0), ("cfg_ddim_step", step_bytes)):
        wrap(n, fn)
    try:
        sched = d._ddim_schedule(50, 1.0)
        tab = d._prologue(den, ws, bufs["cond"], B, [e[0] for e in sched])
        rec.clear()
        shapes.clear()
        L = cfg["seq_len"] * cfg["dancers"]
        s = 25
        t, tn, sr, srm1, sa, c, sigma = sched[s]
        tot0, tot1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        tot0.record()
        d._denoise_step(den, ws, tab, s, bufs["x"], bufs["xpad"], B, bufs["out"])
        ops.cfg_ddim_step(bufs["x"], bufs["out"][: B * L], bufs["out"][B * L:], None, bufs["traj"], bufs["x"], None,
                          bufs["xpad"], 0 if bufs["xpad"] is None else bufs["xpad"].shape[1], B * L, 2.0, sr, srm1, sa, c,
                          sigma, True, False, rng=ent["rng"], rng_stream=1)
        tot1.record()
        torch.cuda.synchronize()
    finally:
        for n, f in orig.items():
            setattr(ops, n, f)
    out = {"eager_step_ms": tot0.elapsed_time(tot1)}
    if os.environ.get("TCD_BENCH_DUMP_GEMMS"):
        for (e0, e1, fl), shp in zip(rec.get("gemm", []), shapes):
            ms = e0.elapsed_time(e1)
            sys.stderr.write("GEMM M=%d N=%d K=%d act=%d out=%s  %.4f ms  %.0f TFLOP/s\n" % (*shp, ms, fl / ms / 1e9))
    hbm = peaks()["hbm"]
    for n, lst in rec.items():
        ms = sum(a.elapsed_time(b) for a, b, _ in lst)
        fl = sum(f for _, _, f in lst)
        out[n] = {"launches": len(lst), "ms": ms, "tflops": (fl / (ms * 1e-3) / 1e12) if ms > 0 and fl > 0 else None,
                  "flops": max(fl, 0.0)}
        if fl < 0 and ms > 0:                                         # HBM-bound kernel class: achieved GB/s vs the measured peak
            gbs = -fl / (ms * 1e-3) / 1e9
            out[n].update(gbs=gbs, frac_of_hbm=gbs / hbm)
        if n == "gemm_film_residual_norm" and ms > 0:                 # fused GEMM + tail: HBM-bound, bytes recorded separately
            gbs = sum(fused_bytes) / (ms * 1e-3) / 1e9
            out[n].update(gbs=gbs, frac_of_hbm=gbs / hbm)
    return out


def train_leg(args, cfg, dev, world, rank):
    """BASELINE.json configs[2]: one optimisation step = zero_grad, p_losses (FK + foot-contact loss), backward,
    gradient all-reduce (N > 1), fused Adan + EMA; batch 128 per GPU, bf16 tape, replayed from CUDA graphs.
    Reported beside the headline metric (samples/s over all ranks, max-over-ranks device time)."""
    import torch.distributed as dist
    import tcdiff_b200 as T
    from tcdiff_b200 import _lib
    from tcdiff_b200.train import GraphedTrainStep
    from tcdiff_b200 import synth
    torch.cuda.empty_cache()
    B, dn, S = args.train_batch, cfg["dancers"], cfg["seq_len"]
    m = T.DanceDecoder(nfeats=151, seq_len=S, latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"],
                       num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], dropout=args.train_dropout,
                       cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=dn, dtype=args.dtype)
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).train()
    d = T.GaussianDiffusion(m, S, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", use_p2=False, cond_drop_prob=0.25, guidance_weight=2).to(dev)
    opt = T.Adan(m.parameters(), lr=4e-4, weight_decay=0.02, data_parallel=world > 1)
    opt.attach_ema(d.master_model, d.model, 0.9999)
    gen = torch.Generator(device=dev).manual_seed(4321 + rank)
    x = torch.randn(B, dn, S, 151, device=dev, generator=gen) * 0.5
    cond = torch.randn(B, 2 * S + 1, cfg["cond_feature_dim"], device=dev, generator=gen)
    step = GraphedTrainStep(d, opt, x, cond, warmup=3)
    for _ in range(2):
        step(x, cond)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES[0]
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = max(3, args.steps)
    e0.record()
    for _ in range(k):
        total, _ = step(x, cond)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / k
    out = {"metric": "training samples/sec (p_losses + backward + grad all-reduce + fused Adan/EMA)", "unit": "samples/s",
           "value": world * B / (ms * 1e-3), "ms_per_step": ms, "n_gpus": world, "steps": k, "batch_per_gpu": B,
           "dtype": args.dtype, "cuda_graph": True, "dropout": args.train_dropout, "loss": float(total),
           "gpu_launches": (_lib.LAUNCHES[0] - l0) // k, "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30,
           # forward + backward = 3x the forward contractions of one un-hoisted pass (SURVEY §8d F_pass), per GPU
           "useful_tflops_per_gpu": 3.0 * flops_per_pass(cfg, hoisted=False) * B / (ms * 1e-3) / 1e12,
           "workload": f"BASELINE configs[2]: p_losses with 6D-rot FK + foot-contact loss, batch {B}/GPU, {dn} dancers, "
                       f"{S} frames, {cfg['cond_feature_dim']}-dim music, data parallel x{world}"}
    del step, opt, d, m
    torch.cuda.empty_cache()
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        out["cpu_baseline"] = cpu_train_sample(args.config)
    return out


def c4_leg(args, dev, world, rank):
    """BASELINE.json configs[3]: Jukebox-style 4800-dim music, 10 dancers, 300 frames (L = 3000 tokens), DDPM-1000 with
    CFG, batch-sharded: `--c4-batch` clips per GPU (1 = the 8-clip job over 8 GPUs), one final gather.  Timed: the LAST
    `--c4-steps` steps of the 1000-step chain (start_point, the reference's own way of starting late,
    model/diffusion.py:263-270) replayed from the captured graphs, scaled to 1000 (every step costs the same)."""
    import torch.distributed as dist
    import tcdiff_b200 as T
    from tcdiff_b200 import synth
    torch.cuda.empty_cache()
    cfg = synth.CONFIGS["c4"]
    S, dn, Fm = cfg["seq_len"], cfg["dancers"], cfg["cond_feature_dim"]
    m = T.DanceDecoder(nfeats=151, seq_len=S, latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"], num_layers=cfg["num_layers"],
                       num_heads=cfg["num_heads"], cond_feature_dim=Fm, required_dancer_num=dn, dtype=args.dtype)
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, S, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", cond_drop_prob=0.25, guidance_weight=2, seq_len=S).to(dev).eval()
    B, steps = args.c4_batch, min(args.c4_steps, 1000)
    shape = (B, S * dn, 151)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    cond = torch.randn(B, 2 * S + 1, Fm, device=dev, generator=gen)
    sp = None if steps >= 1000 else steps
    gathered = torch.empty((world,) + shape, device=dev) if world > 1 else None

    def run():
        out = d.p_sample_loop(shape, cond, start_point=sp)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        return out
    for _ in range(2):
        out = run()                                                  # packs the weights, captures the step graphs
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k = 2
    e0.record()
    for _ in range(k):
        out = run()
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / k
    full_ms = ms * 1000.0 / steps
    tfl = 2 * flops_per_pass(cfg, hoisted=True) * B / (ms / steps * 1e-3) / 1e12       # cond + uncond pass per step
    pk = peaks()
    res = {"metric": "10 s 10-dancer clips/sec (DDPM-1000, Jukebox-style 4800-dim music)", "unit": "clips/s",
           "value": world * B / (full_ms * 1e-3), "n_gpus": world, "batch_per_gpu": B, "ms_per_denoise_step": ms / steps,
           "timed_steps": steps, "extrapolated_to_steps": 1000, "seconds_per_1000_step_call": full_ms * 1e-3,
           "dtype": args.dtype, "cuda_graph": True, "finite": bool(torch.isfinite(out).all()),
           "useful_tflops_per_gpu": tfl, "useful_tensor_frac": tfl / pk["tflops"],
           "workload": f"BASELINE configs[3]: {dn} dancers, {S} frames (L = {S * dn}), {Fm}-dim music, DDPM-1000 with CFG, "
                       f"{B} clip(s)/GPU, batch-sharded x{world}, one final gather",
           "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}
    del d, m
    torch.cuda.empty_cache()
    return res


def c5_leg(args, dev, world, rank):
    """BASELINE.json configs[4]: end-to-end test mode — TrajDecoder trajectory generation + Kalman smoothing feeding
    DDIM-50 with classifier-free guidance and the post-sampling stage (un-normalise, 6D -> axis-angle, SMPL FK);
    `--c5-batch` clips per GPU (32 = batch 256 over 8 GPUs), one final gather of the joint positions."""
    import torch.distributed as dist
    import tcdiff_b200 as T
    from tcdiff_b200 import synth
    torch.cuda.empty_cache()
    cfg = synth.CONFIGS["c2"]
    S, dn, Fm = cfg["seq_len"], cfg["dancers"], cfg["cond_feature_dim"]
    m = T.DanceDecoder(nfeats=151, seq_len=S, latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"], num_layers=cfg["num_layers"],
                       num_heads=cfg["num_heads"], cond_feature_dim=Fm, required_dancer_num=dn, dtype=args.dtype)
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, S, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", cond_drop_prob=0.25, guidance_weight=2).to(dev).eval()
    torch.manual_seed(42)                                            # option_traj.py:63
    traj = T.TrajDecoder(nfeats=2, trans_layer=6, window_size=100).to(dev).eval()      # option_traj.py:33-36
    B = args.c5_batch
    gen = torch.Generator(device=dev).manual_seed(500 + rank)
    x = torch.rand(B, dn, S, 151, device=dev, generator=gen) * 2 - 1
    cond = torch.randn(B, 2 * S + 1, Fm, device=dev, generator=gen)
    norm = (torch.zeros(151, device=dev), torch.ones(151, device=dev))                 # identity MinMax scaler
    shape = (B, S * dn, 151)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t = [0.0, 0.0, 0.0]
    gathered = torch.empty((world, B, dn, S, 24, 3), device=dev) if world > 1 else None

    def step(timed):
        e = [ev() for _ in range(4)]
        e[0].record()
        x_traj = T.generate_trajectory(traj, x, cond, 100, 25)                         # TCDiff.py:526-556
        x0 = x_traj.permute(0, 2, 1, 3).reshape(B, S * dn, 3).contiguous()
        e[1].record()
        samples = d.ddim_sample(shape, cond, x_0=x0)
        e[2].record()
        out = d.samples_to_poses(samples, norm, mode="normal", required_dancer_num=dn)["full_pose"]
        e[3].record()
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        if timed:
            torch.cuda.synchronize()
            for i in range(3):
                t[i] += e[i].elapsed_time(e[i + 1])
        return out
    for _ in range(2):
        step(False)
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()
    k = 3
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(k):
        poses = step(True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.barrier()
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / k
    flops_clip = 50 * 2 * flops_per_pass(cfg, hoisted=True) + 2 * (flops_per_pass(cfg, False) - flops_per_pass(cfg, True))
    tfl = flops_clip * B / (t[1] / k * 1e-3) / 1e12
    pk = peaks()
    res = {"metric": "end-to-end clips/sec (TrajDecoder + Kalman -> DDIM-50 CFG -> FK joint positions)", "unit": "clips/s",
           "value": world * B / (ms * 1e-3), "n_gpus": world, "batch_per_gpu": B, "ms_per_call": ms,
           "ms_front_end": t[0] / k, "ms_sampler": t[1] / k, "ms_post": t[2] / k, "dtype": args.dtype, "cuda_graph": True,
           "finite": bool(torch.isfinite(poses).all()),
           "sampler_useful_tflops_per_gpu": tfl, "sampler_useful_tensor_frac": tfl / pk["tflops"],
           "workload": f"BASELINE configs[4]: batch {B}/GPU x{world} (= {world * B}), {dn} dancers, {S} frames, TrajDecoder 6 layers "
                       f"window 100 step 25 + Kalman, DDIM-50 CFG w=2, samples_to_poses"}
    del d, m, traj
    torch.cuda.empty_cache()
    return res


def cpu_train_sample(cfg_name, batch=2, threads=None):
    """The same optimisation step on the host cores through the oracle (autograd through the fp32 restatement +
    oracle.adan_step/ema_update), bounded: `batch` samples, one warm-up and one timed step."""
    from oracle import synth, tcdiff_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg, 0)
    dead = ("traj_Modulation", "traj_embedding", "embeddings_table")
    names = [k for k, v in sd.items() if v.dtype.is_floating_point and "rotary" not in k and
             not any(d_ in k for d_ in dead)]
    params = {k: sd[k].clone() for k in names}
    ma = {k: sd[k].clone() for k in names}
    st = O.adan_init([params[k] for k in names])
    sched = O.make_schedule("cosine", 1000)
    dn = cfg["dancers"]
    x = synth.make_motion(batch, dn, seed=7)
    cond = synth.make_music(batch, cfg["cond_feature_dim"], seed=8)
    g = torch.Generator().manual_seed(9)
    times = []
    for it in range(2):
        t = torch.randint(0, 1000, (batch,), generator=g)
        noise = torch.randn(batch, cfg["seq_len"], dn, 151, generator=g)
        keep = torch.rand(batch, generator=g) < 0.75
        t0 = time.perf_counter()
        leaf = dict(sd)
        for k in names:
            leaf[k] = params[k].clone().requires_grad_(True)
        total, _ = O.p_losses(leaf, sched, x, cond, t, noise, keep)
        total.backward()
        with torch.no_grad():
            O.adan_step([params[k] for k in names], [leaf[k].grad for k in names], st, lr=4e-4, weight_decay=0.02)
            O.ema_update([ma[k] for k in names], [params[k] for k in names], 0.9999)
        times.append(time.perf_counter() - t0)
    return {"value": batch / times[-1], "unit": "samples/s", "cores": threads, "kind": "port",
            "sample": f"oracle p_losses + autograd backward + adan_step/ema_update, batch {batch}, one timed step "
                      f"({times[-1]:.1f} s) after one warm-up step"}


def cpu_reference_sample(cfg_name, budget_s=20.0, threads=None):
    """The oracle port (unmodified-reference-equivalent PyTorch fp32 CPU evaluation) on a bounded sample of the
    same workload: 1 clip of the c2 shape, n DDIM steps (n chosen to fit the budget), extrapolated to 50."""
    from oracle import synth, tcdiff_oracle as O
    threads = threads or os.cpu_count()
    torch.set_num_threads(threads)
    cfg = synth.CONFIGS[cfg_name]
    sd = synth.make_state_dict(cfg, 0)
    dn = cfg["dancers"]
    shape = (1, 150 * dn, 151)
    cond = synth.make_music(1, cfg["cond_feature_dim"])
    x0 = synth.make_traj(synth.make_motion(1, dn))
    sched = O.make_schedule("cosine", 1000)
    with torch.no_grad():
        t0 = time.time()
        O.guided_forward(sd, torch.randn(shape), cond, torch.tensor([500]), 2.0)       # warm-up + cost probe
        probe = time.time() - t0
        t0 = time.time()
        O.guided_forward(sd, torch.randn(shape), cond, torch.tensor([500]), 2.0)
        per = time.time() - t0
        n = int(max(2, min(50, budget_s / max(per, 1e-3))))
        bank = synth.make_noise_bank(shape, n)
        t0 = time.time()
        O.ddim_sample(sd, sched, shape, cond, x0, bank, sampling_timesteps=n)
        el = time.time() - t0
    clip_s = el * 50.0 / n
    return dict(value=1.0 / clip_s, unit=UNIT, cores=threads, kind="port",
                sample=f"oracle/tcdiff_oracle.py (PyTorch fp32 CPU restatement pinned to the reference), 1 clip of the {cfg_name} "
                       f"workload, {n} of 50 DDIM steps timed ({el:.1f} s) and scaled to 50; first-call {probe:.2f} s",
                ms_per_denoise_step=el / n * 1e3)


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    base = None
    for i in range(max(1, min(args.steps, 3))):
        base = cpu_reference_sample(args.config, budget_s=15.0)
        vals.append(base["value"])
    v = sum(vals) / len(vals)
    base["value"] = v
    line = {"metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": len(vals), "warmup": 1,
            "ms_per_step": 1e3 / v, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "impl": "reference", "config": {"workload": WORKLOAD, "reference_arm": base["sample"]},
            "cpu_baseline": base, "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--config", default="c2")
    ap.add_argument("--batch", type=int, default=64)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-breakdown", action="store_true")
    ap.add_argument("--no-train", action="store_true", help="skip the training-step leg (BASELINE.json configs[2])")
    ap.add_argument("--train-batch", type=int, default=128)
    ap.add_argument("--train-dropout", type=float, default=0.1, help="the reference's training value (TCDiff.py:82)")
    ap.add_argument("--no-c4", action="store_true", help="skip the configs[3] leg (Jukebox-style music, 10 dancers, DDPM-1000)")
    ap.add_argument("--c4-batch", type=int, default=1)
    ap.add_argument("--c4-steps", type=int, default=200, help="timed slice of the 1000-step chain (scaled to 1000)")
    ap.add_argument("--no-c5", action="store_true", help="skip the configs[4] leg (TrajDecoder -> DDIM-50 -> poses)")
    ap.add_argument("--c5-batch", type=int, default=32)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)
    args.warmup = max(args.warmup, 3)

    import torch.distributed as dist
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    from tcdiff_b200 import synth
    from tcdiff_b200 import _lib
    cfg, m, d = build_ours(args.config, args.dtype, dev)
    B, dn = args.batch, cfg["dancers"]
    L = cfg["seq_len"] * dn
    shape = (B, L, 151)
    cond_h = synth.make_music(B, cfg["cond_feature_dim"], seed=1235 + rank).pin_memory()
    x0_h = synth.make_traj(synth.make_motion(B, dn, seed=1234 + rank)).pin_memory()
    out_h = torch.empty(shape, dtype=torch.float32).pin_memory()
    cond_d, x0_d = cond_h.to(dev), x0_h.to(dev)
    gathered = torch.empty((world,) + shape, device=dev) if world > 1 else None

    def step_resident():
        out = d.ddim_sample(shape, cond_d, x_0=x0_d)
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)       # one NCCL all-gather into one tensor (no per-rank copies)
        return out

    def step_e2e():
        out = d.ddim_sample(shape, cond_h.to(dev, non_blocking=True), x_0=x0_h.to(dev, non_blocking=True))
        if world > 1:
            dist.all_gather_into_tensor(gathered, out)
        out_h.copy_(out, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        return float(out_h[0, 0, 0])

    def timed(fn, k):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0 = time.perf_counter()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        wall = (time.perf_counter() - w0) * 1e3
        ms = torch.tensor([e0.elapsed_time(e1), wall], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms[0]), float(ms[1])

    for _ in range(args.warmup):
        step_resident()
    torch.cuda.synchronize()
    l0 = _lib.LAUNCHES[0]
    with ClockSampler(local) as clk:
        ms_total, wall_total = timed(step_resident, args.steps)
    launches_python = _lib.LAUNCHES[0] - l0
    ent = next(e for e in d._graphs.values() if "bufs" in e)
    launches = ent.get("launches_per_call", 0) * args.steps + launches_python
    ms_step = ms_total / args.steps
    value = world * B / (ms_step * 1e-3)
    for _ in range(2):
        step_e2e()
    ms_e2e, _ = timed(step_e2e, args.steps)
    e2e_val = world * B / (ms_e2e / args.steps * 1e-3)

    pk = peaks()
    flops_clip = 50 * 2 * flops_per_pass(cfg, hoisted=True) + 2 * (flops_per_pass(cfg, False) - flops_per_pass(cfg, True))
    step_tflops = flops_clip * B / (ms_step * 1e-3) / 1e12
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": args.dtype, "data": "synthetic",
            "config": {"workload": WORKLOAD, "config": args.config, "batch_per_gpu": B, "dancers": dn, "frames": cfg["seq_len"],
                       "music_dim": cfg["cond_feature_dim"], "sampler": "ddim50_eta1_cfg2", "cuda_graph": True,
                       "l2": "per-step working set (>=0.5 GB of activations per layer pass) exceeds the 126 MB L2; no flush needed",
                       "parallelism": f"batch-sharded x{world}, one final all_gather"},
            "ms_per_denoise_step": ms_step / 50.0, "wall_ms_per_step": wall_total / args.steps,
            "useful_tflops_per_gpu": step_tflops, "useful_tensor_frac": step_tflops / pk["tflops"],
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": cond_h.numel() * 4 + x0_h.numel() * 4,
                    "d2h_bytes_per_step": out_h.numel() * 4},
            "gpu_launches": launches, "clocks": clk.summary()}
    if rank == 0 and not args.no_breakdown:
        bd = kernel_breakdown(d, m, B, cfg)
        g = bd.get("gemm", {})
        fz = bd.get("gemm_film_residual_norm", {})
        at = bd.get("attention", {})
        step_ms_eager = bd.get("eager_step_ms") or 0.0
        share = lambda c: (c.get("ms", 0.0) / step_ms_eager) if step_ms_eager else None
        classes = {
            # ncu --set full of the dominant launch shape (M 96000, N 512, K 512, bf16 out): dram__bytes_read 98.9 MB +
            # dram__bytes_write 52.4 MB per launch against 98.3 + 0.5 read and 98.3 written algorithmically (the rest of C is
            # still in L2 at exit); profiles/r01_ncu_full_summary_final.txt
            "gemm": {"bound": "tensor",
                     "kernel": "gemm_bf16_tc2_kernel (tcgen05/TMA: the %d plain nn.Linear launches of one denoise step)" % g.get("launches", 0),
                     "achieved": g.get("tflops"), "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": (g.get("tflops") or 0.0) / pk["tflops"], "traffic": 151.3e6,
                     "traffic_unit": "B/launch (M96000 N512 K512 GEMM, ncu r01)",
                     "peak_source": pk["src"] + " bf16_tflops_sustained", "share_of_step": share(g)},
            # ncu --set full of the self-attention tail launch (profiles/r02_ncu_gemm_frn2.txt): dram read 297 MB + write 236 MB
            # per launch against 590 MB algorithmic (A partly still in L2)
            "gemm_film_residual_norm": {"bound": "hbm",
                     "kernel": "gemm_frn_kernel (the %d fc / linear2 projections of one denoise step fused with their FiLM + residual "
                               "+ LayerNorm (+ rotary) tails; algorithmic bytes = A + x in + x out + bf16 operands out)" % fz.get("launches", 0),
                     "achieved": fz.get("gbs"), "peak": pk["hbm"], "unit": "GB/s",
                     "frac": (fz.get("gbs") or 0.0) / pk["hbm"], "traffic": 533.4e6,
                     "traffic_unit": "B/launch (self-attention tail, 96000 rows, ncu r02)",
                     "peak_source": pk["src"] + " hbm_gbs", "share_of_step": share(fz)},
            "attention": {"bound": "tensor",
                     "kernel": "attention_tc2q_kernel (tcgen05 flash attention, %d launches; MUFU-bound at head dim 64: 64 ex2 per "
                               "row and key tile = twice the tile's MMA time)" % at.get("launches", 0),
                     "achieved": at.get("tflops"), "peak": pk["tflops"], "unit": "TFLOP/s",
                     "frac": (at.get("tflops") or 0.0) / pk["tflops"], "traffic": None,
                     "peak_source": pk["src"] + " bf16_tflops_sustained", "share_of_step": share(at)},
        }
        # the headline roofline object is the kernel class with the largest share of the step, whichever that is in this run
        top = max(classes, key=lambda k: classes[k]["share_of_step"] or 0.0)
        line["roofline"] = classes[top]
        line["roofline_classes"] = {k: v for k, v in classes.items() if k != top}
        line["kernel_breakdown"] = {k: ({kk: vv for kk, vv in v.items() if kk != "flops"} if isinstance(v, dict) else v)
                                    for k, v in bd.items()}
    if not args.no_train:
        tr = train_leg(args, cfg, dev, world, rank)          # every rank takes part (data parallel)
        if tr is not None:
            line["train"] = tr
    del d, m
    torch.cuda.empty_cache()
    if not args.no_c4:
        line["c4"] = c4_leg(args, dev, world, rank)          # every rank samples its own clip(s)
    if not args.no_c5:
        line["c5"] = c5_leg(args, dev, world, rank)
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_reference_sample(args.config, budget_s=20.0)
    if rank == 0:
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
