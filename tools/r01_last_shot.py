"""One-process GPU check of the end-of-round-1 changes (written for a single short gpurun slot):
  1. tcd_film_residual_norm with x_out = NULL (dead feed-forward residual) is bit-identical on its LayerNorm output;
  2. c2 sampler A/B: TCD_FFN_SKIP_X 0 vs 1 (same seed: outputs must be identical; clips/s each);
  3. tcd_gemm_film_residual_norm (csrc/gemm_frn.cu, EXPERIMENTAL) against a torch fp32 reference and against the
     unfused tcd_gemm + tcd_film_residual_norm pair, then its timing, then the sampler with TCD_FUSE_TAILS.
Results go to stdout and gpurun_out/last_shot.json.  The fused-kernel part runs last: a trapped kernel poisons the
CUDA context, so everything before it is already recorded."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
OUT = os.path.join(ROOT, "gpurun_out", "last_shot.json")
res = {}


def save():
    with open(OUT, "w") as f:
        json.dump(res, f, indent=1)


def log(*a):
    print(*a, flush=True)


t_start = time.time()
import torch  # noqa: E402

import tcdiff_b200 as T  # noqa: E402
from tcdiff_b200 import engine, ops  # noqa: E402
from tcdiff_b200._lib import BF16, ACT_NONE  # noqa: E402
from tcdiff_b200 import synth  # noqa: E402  (synthetic weights / inputs only)

dev = torch.device("cuda:0")
log("import %.1f s" % (time.time() - t_start))


def timeit(fn, n=20, warm=3):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


D, L = 512, 750


def make_case(R, K, seed=0, bias=False, inner=True):
    g = torch.Generator(device="cpu").manual_seed(seed)
    rn = lambda *s: torch.randn(*s, generator=g)
    n = (R + L - 1) // L
    c = dict(
        a=(rn(R, K) * 0.5).to(dev, torch.bfloat16), w=(rn(D, K) / K ** 0.5).to(dev, torch.bfloat16),
        bias=(rn(D) * 0.1).to(dev) if bias else None, x=rn(R, D).to(dev),
        gi=(1 + 0.1 * rn(D)).to(dev) if inner else None, bi=(0.1 * rn(D)).to(dev) if inner else None,
        film=(0.3 * rn(n, 3 * 2 * D)).to(dev), gn=(1 + 0.1 * rn(D)).to(dev), bn=(0.1 * rn(D)).to(dev))
    ang = torch.arange(L, dtype=torch.float32)[:, None] * (10000.0 ** (-torch.arange(0, D, 2).float() / D))[None, :]
    c["cos"], c["sin"] = ang.cos().to(dev).contiguous(), ang.sin().to(dev).contiguous()
    return c


def torch_ref(c, R, foff):
    y = c["a"].float() @ c["w"].float().t()
    if c["bias"] is not None:
        y = y + c["bias"]
    if c["gi"] is not None:
        y = torch.nn.functional.layer_norm(y, (D,), c["gi"], c["bi"], 1e-6)
    samp = torch.arange(R, device=dev) // L
    sc, sh = c["film"][samp, foff:foff + D], c["film"][samp, foff + D:foff + 2 * D]
    v = c["x"] + (1 + sc) * y + sh
    nrm = torch.nn.functional.layer_norm(v, (D,), c["gn"], c["bn"], 1e-5)
    pos = torch.arange(R, device=dev) % L
    cs, sn = c["cos"][pos], c["sin"][pos]
    rot = torch.empty_like(nrm)
    rot[:, 0::2] = nrm[:, 0::2] * cs - nrm[:, 1::2] * sn
    rot[:, 1::2] = nrm[:, 1::2] * cs + nrm[:, 0::2] * sn
    return v, nrm, rot


def run_unfused(c, R, K, foff, want_x=True, want_plain=True, want_rot=True):
    y = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    ops.gemm(c["a"], c["w"], c["bias"], ACT_NONE, y, M=R)
    x = c["x"].clone()
    plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_plain else None
    rot = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_rot else None
    ops.film_residual_norm(BF16, x, x if want_x else None, y, (c["gi"], c["bi"]) if c["gi"] is not None else None, 1e-6,
                           c["film"], c["film"].stride(0), foff, (c["gn"], c["bn"]), 1e-5, plain, rot,
                           c["cos"] if want_rot else None, c["sin"] if want_rot else None, R, D, L)
    return x, plain, rot


def run_fused(c, R, K, foff, want_x=True, want_plain=True, want_rot=True):
    x = c["x"].clone()
    plain = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_plain else None
    rot = torch.zeros(R, D, device=dev, dtype=torch.bfloat16) if want_rot else None
    ops.gemm_film_residual_norm(c["a"], c["w"], c["bias"], x, x if want_x else None,
                                (c["gi"], c["bi"]) if c["gi"] is not None else None, 1e-6, c["film"], c["film"].stride(0),
                                foff, (c["gn"], c["bn"]), 1e-5, plain, rot, c["cos"] if want_rot else None,
                                c["sin"] if want_rot else None, R, L)
    return x, plain, rot


# ---------------------------------------------------------------- 1. x_out = NULL
try:
    c = make_case(6000, 1024, seed=1, bias=True, inner=False)
    x1, p1, _ = run_unfused(c, 6000, 1024, 2 * D, want_x=True, want_rot=False)
    x2, p2, _ = run_unfused(c, 6000, 1024, 2 * D, want_x=False, want_rot=False)
    torch.cuda.synchronize()
    res["frn_null_xout"] = {"plain_bit_identical": bool(torch.equal(p1, p2)), "x_untouched": bool(torch.equal(x2, c["x"]))}
    log("1. frn x_out=NULL:", res["frn_null_xout"])
except Exception as e:  # noqa: BLE001
    res["frn_null_xout"] = {"error": repr(e)}
    log("1. FAILED", repr(e))
save()


# ---------------------------------------------------------------- 2. sampler A/B
def build_sampler():
    cfg = synth.CONFIGS["c2"]
    m = T.DanceDecoder(nfeats=151, seq_len=cfg["seq_len"], latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"],
                       num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], dropout=0.1,
                       cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"], dtype="bf16")
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, cfg["seq_len"], 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000,
                            predict_epsilon=False, loss_type="l2", use_p2=False, cond_drop_prob=0.25,
                            guidance_weight=2).to(dev).eval()
    return cfg, d


B = 64
cfg = synth.CONFIGS["c2"]
shape = (B, cfg["seq_len"] * cfg["dancers"], 151)
cond = synth.make_music(B, cfg["cond_feature_dim"], seed=1235).to(dev)
x0 = synth.make_traj(synth.make_motion(B, cfg["dancers"], seed=1234)).to(dev)


def sampler_variant(name, skip_x, fuse, steps=4):
    engine.SKIP_DEAD_X, engine.FUSE_TAILS = skip_x, fuse
    _, d = build_sampler()
    torch.manual_seed(7)
    out = d.ddim_sample(shape, cond, x_0=x0).clone()
    for _ in range(2):
        d.ddim_sample(shape, cond, x_0=x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        d.ddim_sample(shape, cond, x_0=x0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    res.setdefault("sampler", {})[name] = {"ms_per_clip_batch": ms, "clips_per_s": B / (ms * 1e-3), "ms_per_denoise_step": ms / 50,
                                           "finite": bool(torch.isfinite(out).all())}
    log("2/5. sampler", name, res["sampler"][name])
    save()
    del d
    return out


base_out = None
try:
    base_out = sampler_variant("base", False, 0)
    skip_out = sampler_variant("ffn_skip_x", True, 0)
    res["sampler"]["ffn_skip_x"]["identical_to_base"] = bool(torch.equal(base_out, skip_out))
    res["sampler"]["ffn_skip_x"]["max_abs_diff"] = float((base_out - skip_out).abs().max())
    log("   skip_x vs base:", res["sampler"]["ffn_skip_x"])
    del skip_out
except Exception as e:  # noqa: BLE001
    res.setdefault("sampler", {})["error"] = repr(e)
    log("2. FAILED", repr(e))
save()

# ---------------------------------------------------------------- 3. fused kernel: correctness
ok_fused = True
try:
    cases = [("sa  R96000 K512 inner rot", 96000, 512, False, True, 0, True, False, True),
             ("ca  R3000  K512 inner plain", 3000, 512, False, True, 2 * D, True, True, False),
             ("ffn R1130  K1024 bias plain no-x (ragged rows)", 1130, 1024, True, False, 4 * D, False, True, False),
             ("all R750   K512 inner plain+rot", 750, 512, False, True, 0, True, True, True)]
    res["fused"] = {}
    for name, R, K, bias, inner, foff, wx, wp, wr in cases:
        c = make_case(R, K, seed=3, bias=bias, inner=inner)
        v_ref, n_ref, r_ref = torch_ref(c, R, foff)
        xu, pu, ru = run_unfused(c, R, K, foff, wx, wp, wr)
        xf, pf, rf = run_fused(c, R, K, foff, wx, wp, wr)
        torch.cuda.synchronize()
        e = {}
        if wx:
            e["x_fused"], e["x_unfused"] = float((xf - v_ref).abs().max()), float((xu - v_ref).abs().max())
        else:
            e["x_untouched"] = bool(torch.equal(xf, c["x"]))
        if wp:
            e["plain_fused"], e["plain_unfused"] = float((pf.float() - n_ref).abs().max()), float((pu.float() - n_ref).abs().max())
        if wr:
            e["rot_fused"], e["rot_unfused"] = float((rf.float() - r_ref).abs().max()), float((ru.float() - r_ref).abs().max())
        good = all(e.get(k + "_fused", 0.0) <= 1.5 * e.get(k + "_unfused", 1.0) + 1e-3 for k in ("x", "plain", "rot")) and \
            e.get("x_untouched", True)
        e["ok"] = bool(good)
        ok_fused = ok_fused and good
        res["fused"][name] = e
        log("3. fused", name, e)
        save()
except Exception as e:  # noqa: BLE001
    ok_fused = False
    res.setdefault("fused", {})["error"] = repr(e)
    log("3. FAILED", repr(e))
save()

# ---------------------------------------------------------------- 4. fused kernel: timing
if ok_fused:
    try:
        R = 96000
        res["fused_timing_ms"] = {}
        for name, K, bias, inner, wx, wp, wr in (("sa_tail K512 (x, rot)", 512, False, True, True, False, True),
                                                 ("ca_tail K512 (x, plain)", 512, False, True, True, True, False),
                                                 ("ffn_tail K1024 (plain, no x)", 1024, True, False, False, True, False)):
            c = make_case(R, K, seed=5, bias=bias, inner=inner)
            tu = timeit(lambda: run_unfused(c, R, K, 0, wx, wp, wr), n=10)
            tf = timeit(lambda: run_fused(c, R, K, 0, wx, wp, wr), n=10)
            tclone = timeit(lambda: (c["x"].clone(), torch.zeros(R, D, device=dev, dtype=torch.bfloat16)), n=10)
            res["fused_timing_ms"][name] = {"unfused_pair": tu - tclone, "fused": tf - tclone, "harness_clone_zero": tclone}
            # phase breakdown of ONE instrumented launch (cycle counters summed over CTAs; see tcd_gemm_frn_set_debug)
            from tcdiff_b200 import _lib
            dbg = torch.zeros(8, dtype=torch.int64, device=dev)
            _lib.lib().tcd_gemm_frn_set_debug(dbg.data_ptr())
            run_fused(c, R, K, 0, wx, wp, wr)
            torch.cuda.synchronize()
            _lib.lib().tcd_gemm_frn_set_debug(0)
            ctas = min(148, (R + 127) // 128)
            names = ["epi_wait_mma", "pass1", "pass2", "pass2b", "pass3", "mma_wait_epi", "mma_loop", "mma_wait_tma"]
            res["fused_timing_ms"][name]["kcycles_per_cta"] = {n: round(float(v) / ctas / 1e3, 1) for n, v in zip(names, dbg.tolist())}
            log("4. timing", name, res["fused_timing_ms"][name])
            save()
    except Exception as e:  # noqa: BLE001
        res["fused_timing_ms"] = {"error": repr(e)}
        log("4. FAILED", repr(e))
    # ------------------------------------------------------------ 5. sampler with fused tails
    try:
        for mask in (7, 6):
            o = sampler_variant("fuse_tails_%d_skip_x" % mask, True, mask)
            if base_out is not None:
                res["sampler"]["fuse_tails_%d_skip_x" % mask]["max_abs_diff_vs_base"] = float((o - base_out).abs().max())
                res["sampler"]["fuse_tails_%d_skip_x" % mask]["mean_abs_diff_vs_base"] = float((o - base_out).abs().mean())
            log("   ", res["sampler"]["fuse_tails_%d_skip_x" % mask])
            save()
    except Exception as e:  # noqa: BLE001
        res["sampler"]["fuse_error"] = repr(e)
        log("5. FAILED", repr(e))
res["total_s"] = time.time() - t_start
save()
log("done %.1f s" % res["total_s"])
