"""Micro-benchmarks of single kernels at BASELINE-c2 shapes (CUDA events, L2 flushed between launches).
Usage: python tools/kernel_bench.py [gemm|attn|frn|fused|step|loss|train|sampler|all] [--ncu] [--lib path/to/lib.so]
  --ncu: one launch each, no timing loop
  --lib: load an A/B build of the library (python -m tcdiff_b200.build --define ... --out ...) instead of the product one
  sampler: c2 DDIM-50 clips/s at batch 64 in this process + a checksum of the samples (same seed in every A/B process)
"""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tcdiff_b200 import ops, _lib as _tlib
from tcdiff_b200._lib import BF16

if "--lib" in sys.argv:
    _tlib.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
    sys.argv.pop(sys.argv.index("--lib") + 1)
print("library:", _tlib.LIB_PATH, {k: _tlib.lib().tcd_tuning(k.encode()) for k in ("fuse_tails", "attn_2q", "frn_rc")})

dev = torch.device("cuda:0")
NCU = "--ncu" in sys.argv
which = [a for a in sys.argv[1:] if not a.startswith("--")] or ["all"]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
PEAKS = json.load(open(os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json"))) \
    if os.path.exists("MEASURED_PEAKS.json") else {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}


def timeit(fn, iters=10):
    if NCU:
        fn()
        torch.cuda.synchronize()
        return float("nan")
    for _ in range(3):
        fn()
    ts = []
    for _ in range(iters):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


res = {}
if "gemm" in which or "all" in which:
    for (M, N, K, outbf, act, bias) in [(96000, 512, 512, True, 0, False), (96000, 1024, 512, True, 0, False),
                                        (96000, 1024, 512, True, 2, True), (96000, 512, 1024, True, 0, True),
                                        (96000, 512, 512, False, 0, True), (9600, 2560, 1024, False, 0, True),
                                        (8192, 8192, 8192, True, 0, False)]:
        a = torch.randn(M, K, device=dev).bfloat16()
        w = (torch.randn(N, K, device=dev) / K ** 0.5).bfloat16()
        b = torch.randn(N, device=dev) if bias else None
        out = torch.empty(M, N, device=dev, dtype=torch.bfloat16 if outbf else torch.float32)
        ms = timeit(lambda: ops.gemm(a, w, b, act, out))
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        res[f"gemm M{M} N{N} K{K} out{'bf16' if outbf else 'f32'} act{act}"] = dict(ms=ms, tflops=tf, frac_of_burst=tf / PEAKS["bf16_tflops"])
        if NCU:
            break
if "attn" in which or "all" in which:
    for (n, Lq, Lk) in [(128, 750, 750), (128, 750, 152)]:
        H, HD = 8, 512
        qk = torch.randn(n, max(Lq, Lk), 2 * HD, device=dev).bfloat16()
        v = torch.randn(n, Lk, HD, device=dev).bfloat16()
        o = torch.empty(n, Lq, HD, device=dev, dtype=torch.bfloat16)
        Lp = max(Lq, Lk)
        ms = timeit(lambda: ops.attention(qk, 2 * HD, Lp * 2 * HD, qk, 2 * HD, Lp * 2 * HD, v, HD, Lk * HD, o, HD, Lq * HD, n, H, Lq, Lk, 0.125, k_off=HD))
        res[f"attn n{n} Lq{Lq} Lk{Lk}"] = dict(ms=ms, tflops=4.0 * n * H * Lq * Lk * 64 / (ms * 1e-3) / 1e12)
if "frn" in which or "all" in which:
    R, D, L = 96000, 512, 750
    x = torch.randn(R, D, device=dev)
    y = torch.randn(R, D, device=dev).bfloat16()
    g = torch.randn(D, device=dev)
    film = torch.randn(128, 24576, device=dev)
    cs = torch.randn(L, D // 2, device=dev)
    op = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    orot = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    ms = timeit(lambda: ops.film_residual_norm(BF16, x, x, y, (g, g), 1e-6, film, 24576, 0, (g, g), 1e-5, None, orot, cs, cs, R, D, L))
    byt = R * D * (4 + 4 + 2 + 2)
    res["film_residual_norm R96000"] = dict(ms=ms, gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
    ms = timeit(lambda: ops.layernorm_rotary(x, g, g, 1e-5, op, orot, cs, cs, R, D, L))
    byt = R * D * (4 + 2 + 2)
    res["layernorm_rotary R96000"] = dict(ms=ms, gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
if "fused" in which:
    # the three decoder tails: unfused pair (tcd_gemm + tcd_film_residual_norm) against the fused kernel (csrc/gemm_frn.cu)
    R, D, L = 96000, 512, 750
    x = torch.randn(R, D, device=dev)
    g = torch.randn(D, device=dev)
    film = torch.randn(128, 24576, device=dev)
    cs = torch.randn(L, D // 2, device=dev)
    op = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    y = torch.empty(R, D, device=dev, dtype=torch.bfloat16)
    cst = cs.t().contiguous()                            # the fused entry reads the rotary tables angle-major
    for name, K, bias, inner, want_x, want_plain, want_rot in (("self-attention", 512, False, True, True, False, True),
                                                               ("cross-attention", 512, False, True, True, True, False),
                                                               ("feed-forward", 1024, True, False, False, True, False)):
        a = torch.randn(R, K, device=dev).bfloat16()
        w = (torch.randn(D, K, device=dev) / K ** 0.5).bfloat16()
        b = torch.randn(D, device=dev) if bias else None
        ln_in = (g, g) if inner else None
        pl, ro = (op if want_plain else None), (op if want_rot else None)

        def pair():
            ops.gemm(a, w, b, 0, y)
            ops.film_residual_norm(BF16, x, x if want_x else None, y, ln_in, 1e-6, film, 24576, 0, (g, g), 1e-5, pl, ro,
                                   cs if want_rot else None, cs if want_rot else None, R, D, L)

        def fused():
            ops.gemm_film_residual_norm(a, w, b, x, x if want_x else None, ln_in, 1e-6, film, 24576, 0, (g, g), 1e-5, pl, ro,
                                        cst if want_rot else None, cst if want_rot else None, R, L)

        byt = R * (K * 2 + D * (4 + (4 if want_x else 0) + 2))
        mp, mf = timeit(pair), timeit(fused)
        if "--phases" in sys.argv:                         # cycle counters of the fused kernel's roles, per tile
            cnt = torch.zeros(8, dtype=torch.int64, device=dev)
            _tlib.lib().tcd_gemm_frn_set_debug(cnt.data_ptr())
            fused()
            torch.cuda.synchronize()
            _tlib.lib().tcd_gemm_frn_set_debug(0)
            ntile = (R + 127) // 128 * 2                     # (tile, CTA) pairs: one epilogue warp and the MMA warp of each report
            c = [int(v) // ntile for v in cnt.tolist()]
            print(f"  phases {name}: epilogue waits MMAs {c[0]}, pass1 {c[1]}, pass2 {c[2]} (residual boxes {c[5]}), exchange2 {c[3]}, "
                  f"pass3 {c[4]}, in exchanges {c[6]} | MMA warp waits epilogue {c[7]}")
        res[f"tail {name} K{K}"] = dict(pair_ms=mp, fused_ms=mf, fused_gbs=byt / (mf * 1e-3) / 1e9, fused_frac_hbm=byt / (mf * 1e-3) / 1e9 / PEAKS["hbm_gbs"])

    # linear3 + LayerNorm + rotary of the next layer (no residual, no FiLM; x written out): unfused = GEMM (fp32 out) + tcd_layernorm_rotary
    a = torch.randn(R, 512, device=dev).bfloat16()
    w = (torch.randn(D, 512, device=dev) / 512 ** 0.5).bfloat16()
    b = torch.randn(D, device=dev)
    xo = torch.empty(R, D, device=dev)
    orot = torch.empty(R, D, device=dev, dtype=torch.bfloat16)

    def pair3():
        ops.gemm(a, w, b, 0, xo)
        ops.layernorm_rotary(xo, g, g, 1e-5, op, orot, cs, cs, R, D, L)

    def fused3():
        ops.gemm_film_residual_norm(a, w, b, None, xo, None, 0.0, None, 0, 0, (g, g), 1e-5, op, orot, cst, cst, R, L)

    mp, mf = timeit(pair3), timeit(fused3)
    res["tail linear3 + norm1/rotary K512"] = dict(pair_ms=mp, fused_ms=mf)
if "step" in which or "all" in which:
    for B in (64, 512):
        n = B * 750
        x, c, u, nz = (torch.randn(n, 151, device=dev) for _ in range(4))
        tr = torch.randn(n, 3, device=dev)
        xpad = torch.zeros(n, 160, device=dev, dtype=torch.bfloat16)
        ms = timeit(lambda: ops.cfg_ddim_step(x, c, u, nz, tr, x, None, xpad, 160, n, 2.0, 1.5, 1.1, 0.9, 0.3, 0.2, True, False))
        byt = n * 151 * 20 + n * 8          # 20 B/element (+8 B/token trajectory); bf16 copy adds 2 B/element
        res[f"cfg_ddim_step B{B}"] = dict(ms=ms, gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
if "loss" in which or "all" in which:
    for B, dn in ((128, 3), (128, 5), (1024, 5)):
        S = 150
        mo = torch.rand(B, S, dn, 151, device=dev) * 2 - 1
        tg = torch.rand(B, S, dn, 151, device=dev) * 2 - 1
        from tcdiff_b200 import _lib
        ws = torch.empty(_lib.lib().tcd_loss_workspace_floats(B, S, dn), device=dev)
        out5 = torch.empty(5, device=dev)
        st = torch.cuda.current_stream().cuda_stream
        ms = timeit(lambda: _lib.check(_lib.lib().tcd_loss_forward(mo.data_ptr(), tg.data_ptr(), 0, ws.data_ptr(), out5.data_ptr(), B, S, dn, 0, st)))
        byt = B * S * dn * 1208
        res[f"loss_forward B{B} dn{dn}"] = dict(ms=ms, gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
        # the same call back to back over rotating inputs larger than L2 together (no flush kernel, no event per launch):
        # what the loss kernel costs inside a step, where its launch latency is hidden behind the kernel before it
        if not NCU:
            nset = max(2, -(-(300 << 20) // byt))
            sets = [(torch.rand(B, S, dn, 151, device=dev) * 2 - 1, torch.rand(B, S, dn, 151, device=dev) * 2 - 1) for _ in range(nset)]
            def burst():
                for a_, b_ in sets:
                    _lib.check(_lib.lib().tcd_loss_forward(a_.data_ptr(), b_.data_ptr(), 0, ws.data_ptr(), out5.data_ptr(), B, S, dn, 0, st))
            for _ in range(2):
                burst()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); burst(); burst(); e1.record(); torch.cuda.synchronize()
            msb = e0.elapsed_time(e1) / (2 * nset)
            res[f"loss_forward B{B} dn{dn} back-to-back"] = dict(ms=msb, gbs=byt / (msb * 1e-3) / 1e9, frac=byt / (msb * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
            del sets
if "train" in which or "all" in which:
    # training-step kernels: attention forward(+LSE)/backward, wgrad GEMM, LayerNorm backward, optimizer
    from tcdiff_b200 import _lib
    lib = _lib.lib()
    st = torch.cuda.current_stream().cuda_stream
    for (n, Lq, Lk) in [(128, 750, 750), (128, 750, 152)]:
        H, HD = 8, 512
        qk = torch.randn(n, max(Lq, Lk), 2 * HD, device=dev).bfloat16()
        q, k = qk[:, :Lq, :HD], qk[:, :Lk, HD:]
        v = torch.randn(n, Lk, HD, device=dev).bfloat16()
        do = torch.randn(n, Lq, HD, device=dev).bfloat16()
        o, lse = ops.attention_train_forward(q, k, v, H, 0.125)
        unit = 2.0 * n * H * Lq * Lk * 64
        ms = timeit(lambda: ops.attention_train_forward(q, k, v, H, 0.125))
        res[f"attn_train_fwd n{n} Lq{Lq} Lk{Lk}"] = dict(ms=ms, tflops=2 * unit / (ms * 1e-3) / 1e12)
        dq, dk, dv = ops.attention_train_backward(q, k, v, o, do, lse, H, 0.125)
        ms = timeit(lambda: ops.attention_train_backward(q, k, v, o, do, lse, H, 0.125, dq=dq, dk=dk, dv=dv))
        # 7 GEMM units executed (S and dP recomputed in both kernels); 5 are algorithmically necessary
        res[f"attn_train_bwd n{n} Lq{Lq} Lk{Lk}"] = dict(ms=ms, tflops_executed=7 * unit / (ms * 1e-3) / 1e12,
                                                         tflops_algorithmic=5 * unit / (ms * 1e-3) / 1e12)
        rng = torch.tensor([123, 7], dtype=torch.int64, device=dev)
        ms = timeit(lambda: ops.attention_train_forward(q, k, v, H, 0.125, 0.1, rng, 3))
        res[f"attn_train_fwd+dropout n{n} Lq{Lq} Lk{Lk}"] = dict(ms=ms, tflops=2 * unit / (ms * 1e-3) / 1e12)
        ms = timeit(lambda: ops.attention_train_backward(q, k, v, o, do, lse, H, 0.125, dq=dq, dk=dk, dv=dv, dropout_p=0.1,
                                                         rng_state=rng, site=3))
        res[f"attn_train_bwd+dropout n{n} Lq{Lq} Lk{Lk}"] = dict(ms=ms, tflops_executed=7 * unit / (ms * 1e-3) / 1e12)
        if NCU:
            break
    xd = torch.randn(96000, 512, device=dev).bfloat16()
    yd = torch.empty_like(xd)
    rng = torch.tensor([123, 7], dtype=torch.int64, device=dev)
    ms = timeit(lambda: ops.dropout(xd, 0.1, rng, 5, out=yd))
    res["dropout bf16 96000x512"] = dict(ms=ms, gbs=xd.numel() * 4 / (ms * 1e-3) / 1e9, frac=xd.numel() * 4 / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
    for (K, M, N) in [(96000, 512, 512), (96000, 1024, 512), (96000, 512, 1024), (19200, 1024, 2560)]:
        a = torch.randn(K, M, device=dev).bfloat16()
        b = torch.randn(K, N, device=dev).bfloat16()
        out = torch.empty(M, N, device=dev)
        ws = torch.empty(lib.tcd_gemm_tn_workspace_floats(M, N, K), device=dev)
        ms = timeit(lambda: _lib.check(lib.tcd_gemm_tn(a.data_ptr(), M, b.data_ptr(), N, out.data_ptr(), N, M, N, K, ws.data_ptr(), st)))
        tf = 2.0 * M * N * K / (ms * 1e-3) / 1e12
        res[f"gemm_tn (wgrad) tokens{K} M{M} N{N}"] = dict(ms=ms, tflops=tf, frac_of_burst=tf / PEAKS["bf16_tflops"])
        if NCU:
            break
    R, D, L = 96000, 512, 750
    x = torch.randn(R, D, device=dev)
    g = torch.randn(D, device=dev)
    dy = torch.randn(R, D, device=dev).bfloat16()
    dyr = torch.randn(R, D, device=dev).bfloat16()
    dres = torch.randn(R, D, device=dev)
    dx = torch.empty(R, D, device=dev)
    cs = torch.randn(L, D // 2, device=dev)
    P_ = lib.tcd_layernorm_backward_mixed_partials(R)
    pg = torch.empty(2, P_, D, device=dev)
    ms = timeit(lambda: _lib.check(lib.tcd_layernorm_backward_mixed(0, 1, x.data_ptr(), g.data_ptr(), dy.data_ptr(), dyr.data_ptr(),
                                                                    cs.data_ptr(), cs.data_ptr(), L, 1e-5, dres.data_ptr(), dx.data_ptr(),
                                                                    pg[0].data_ptr(), pg[1].data_ptr(), R, D, 0.0, 0, 0, st)))
    byt = R * D * (4 + 2 + 2 + 4 + 4)
    res["layernorm_backward_mixed(+rotary,+residual) R96000"] = dict(ms=ms, gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
    nparam = 56_000_000
    bufs = [torch.randn(nparam, device=dev) * 0.01 for _ in range(7)]
    bufs[5].abs_()
    ms = timeit(lambda: _lib.check(lib.tcd_adan_ema_step(*[t.data_ptr() for t in bufs], nparam, 3, 1.0, 4e-4, 0.02, 0.08, 0.01, 1e-8, 0.02,
                                                        0.9999, st)))
    byt = nparam * 52
    res["adan_ema_step 56M params"] = dict(ms=ms, gbs=byt / (ms * 1e-3) / 1e9, frac=byt / (ms * 1e-3) / 1e9 / PEAKS["hbm_gbs"])
if "sampler" in which:
    import tcdiff_b200 as T
    from tcdiff_b200 import synth
    cfg = synth.CONFIGS["c2"]
    m = T.DanceDecoder(nfeats=151, seq_len=150, latent_dim=512, ff_size=cfg["ff_size"], num_layers=cfg["num_layers"], num_heads=8,
                       dropout=0.1, cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"], dtype="bf16")
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, 150, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", cond_drop_prob=0.25, guidance_weight=2).to(dev).eval()
    B = 64
    cond = synth.make_music(B, cfg["cond_feature_dim"], seed=1235).to(dev)
    x0 = synth.make_traj(synth.make_motion(B, cfg["dancers"], seed=1234)).to(dev)
    shape = (B, 750, 151)
    for _ in range(2):
        out = d.ddim_sample(shape, cond, x_0=x0, seed=777)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        out = d.ddim_sample(shape, cond, x_0=x0, seed=777)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    res["sampler c2 B64 DDIM-50"] = dict(ms_per_call=ms, clips_per_s=B / (ms * 1e-3), ms_per_denoise_step=ms / 50,
                                         checksum=float(out.double().abs().sum()), first=[float(v) for v in out[0, 0, :4]])
for k, v in res.items():
    print(k, json.dumps(v))
