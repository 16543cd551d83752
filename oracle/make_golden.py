"""TEST INFRASTRUCTURE — generate tests/golden/*.pt by running the UNMODIFIED reference.

Run in the build container only (needs /root/reference):

    python -m oracle.make_golden            # ~1-2 min on 8 cores

Every fixture stores the seeds / config needed to regenerate its inputs with oracle/synth.py,
a checksum of the synthetic weights, and the reference's outputs.  While generating, the
functional restatement (oracle/tcdiff_oracle.py) is checked against the reference and the
max abs difference is recorded in the fixture (``oracle_maxdiff``).
"""
import os
import sys
import time

import torch
import torch.nn.functional as F

from . import ref_shim, synth, tcdiff_oracle as O

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def build_reference(cfg, sd):
    ns = ref_shim.load()
    m = ns.DanceDecoder(nfeats=cfg["nfeats"], seq_len=cfg["seq_len"], latent_dim=cfg["latent_dim"],
                        ff_size=cfg["ff_size"], num_layers=cfg["num_layers"], num_heads=cfg["num_heads"],
                        dropout=0.1, cond_feature_dim=cfg["cond_feature_dim"], activation=F.gelu,
                        required_dancer_num=cfg["dancers"])
    m.load_state_dict(sd, strict=True)
    diff = ns.GaussianDiffusion(m, cfg["seq_len"], cfg["nfeats"], ns.SMPLSkeleton(None), schedule="cosine",
                                n_timestep=1000, predict_epsilon=False, loss_type="l2", use_p2=False,
                                cond_drop_prob=0.25, guidance_weight=2)      # TCDiff.py:90-102
    return m.eval(), diff.eval()


def save(name, obj):
    os.makedirs(OUT, exist_ok=True)
    path = os.path.join(OUT, name)
    torch.save(obj, path)
    print(f"  wrote {name}: {os.path.getsize(path) / 1024:.0f} KiB")


def gen_schedule():
    ns = ref_shim.load()
    diff = ns.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, None, schedule="cosine", n_timestep=1000,
                                predict_epsilon=False, loss_type="l2")
    ref = {k: v.clone() for k, v in diff.named_buffers()}
    mine = O.make_schedule("cosine", 1000)
    md = max(float((ref[k] - mine[k]).abs().max()) for k in mine)
    assert md == 0.0, md
    lin = ns.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, None, schedule="linear", n_timestep=1000)
    ref_lin = {k: v.clone() for k, v in lin.named_buffers()}
    mdl = max(float((ref_lin[k] - O.make_schedule("linear", 1000)[k]).abs().max()) for k in mine)
    assert mdl == 0.0, mdl
    save("schedule.pt", {"cosine": ref, "linear_alphas_cumprod": ref_lin["alphas_cumprod"], "oracle_maxdiff": md})


def gen_forward(name, B=2, times=(500, 17)):
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    m, _ = build_reference(cfg, sd)
    L = 150 * cfg["dancers"]
    x = torch.randn(B, L, 151, generator=torch.Generator().manual_seed(5))
    cond = synth.make_music(B, cfg["cond_feature_dim"])
    t = torch.tensor(times)
    with torch.no_grad():
        rc = m(x, cond, t, cond_drop_prob=0.0)
        ru = m(x, cond, t, cond_drop_prob=1.0)
        rg = m.guided_forward(x, cond, t, 2.0)
        md = max(float((rc - O.dance_decoder_forward(sd, x, cond, t, cond_drop_prob=0.0)).abs().max()),
                 float((ru - O.dance_decoder_forward(sd, x, cond, t, cond_drop_prob=1.0)).abs().max()),
                 float((rg - O.guided_forward(sd, x, cond, t, 2.0)).abs().max()))
    print(f"  {name} forward: oracle vs reference max|d| = {md:.3e}")
    assert md < 2e-5
    stride = 1 if name == "tiny" else 3
    save(f"{name}_forward.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd),
                                "x_seed": 5, "B": B, "times": list(times), "row_stride": stride,
                                "cond": rc[:, ::stride].clone(), "uncond": ru[:, ::stride].clone(),
                                "guided": rg[:, ::stride].clone(), "oracle_maxdiff": md})


def gen_ddim(name, B=2):
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    _, diff = build_reference(cfg, sd)
    dn = cfg["dancers"]
    shape = (B, 150 * dn, 151)
    cond = synth.make_music(B, cfg["cond_feature_dim"])
    x0 = synth.make_traj(synth.make_motion(B, dn))
    bank = synth.make_noise_bank(shape, 49)
    t0 = time.time()
    with ref_shim.NoiseBank(bank) as nb:
        ref = diff.ddim_sample(shape, cond, x_0=x0.clone())
        assert nb.i == 50, nb.i
    t_ref = time.time() - t0
    trace = []
    mine = O.ddim_sample(sd, O.make_schedule("cosine", 1000), shape, cond, x0, bank, trace=trace)
    md = float((ref - mine).abs().max())
    print(f"  {name} ddim-50: oracle vs reference max|d| = {md:.3e}  (reference took {t_ref:.1f} s)")
    assert md < 5e-3
    stride = 1 if name == "tiny" else 3
    save(f"{name}_ddim.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd),
                             "B": B, "noise_seed": 4321, "row_stride": stride,
                             "out": ref[:, ::stride].clone(), "oracle_maxdiff": md,
                             "reference_seconds": t_ref,
                             # teacher-forcing trace from the (validated) oracle: x_t entering steps 0, 25, 49
                             "trace_steps": [0, 25, 49],
                             "trace_row_stride": 4 * stride,
                             "trace_x": [trace[i][0][:, ::4 * stride].clone() for i in (0, 25, 49)],
                             "trace_x_start": [trace[i][1][:, ::4 * stride].clone() for i in (0, 25, 49)]})


def gen_ddpm(name="tiny", B=2, start_point=12):
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    _, diff = build_reference(cfg, sd)
    shape = (B, 150 * cfg["dancers"], 151)
    cond = synth.make_music(B, cfg["cond_feature_dim"])
    bank = synth.make_noise_bank(shape, start_point, seed=777)
    with ref_shim.NoiseBank(bank[1:]) as nb:
        ref = diff.p_sample_loop(shape, cond, noise=bank[0].clone(), start_point=start_point)
        assert nb.i == start_point
    mine = O.p_sample_loop(sd, O.make_schedule("cosine", 1000), shape, cond, bank, start_point=start_point)
    md = float((ref - mine).abs().max())
    print(f"  {name} ddpm last-{start_point}: oracle vs reference max|d| = {md:.3e}")
    assert md < 1e-4
    save(f"{name}_ddpm.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd),
                             "B": B, "noise_seed": 777, "start_point": start_point, "out": ref.clone(),
                             "oracle_maxdiff": md})


def gen_variants(name="tiny", B=3):
    """long_ddim_sample, ddim_sample_Footwork and long_inpaint_loop of the unmodified reference."""
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    _, diff = build_reference(cfg, sd)
    dn = cfg["dancers"]
    shape = (B, 150 * dn, 151)
    sched = O.make_schedule("cosine", 1000)
    cond = synth.make_music(B, cfg["cond_feature_dim"], seed=61)
    motion = synth.make_motion(B, dn, seed=62)
    traj4 = synth.make_traj(motion).reshape(B, 150, dn, 3)
    bank = synth.make_noise_bank(shape, 49, seed=63)
    with ref_shim.NoiseBank(bank):
        ref_long = diff.long_ddim_sample(shape, cond, traj4.clone())
    mine = O.ddim_sample(sd, sched, shape, cond, traj4.reshape(B, -1, 3), bank, long_mode=True)
    md_long = float((ref_long - mine).abs().max())
    full = synth.make_prediction(B, dn, seed=64)                    # (B, S*dn, 151) "ground-truth" motion for the foot constraint
    with ref_shim.NoiseBank(bank):
        ref_foot = diff.ddim_sample_Footwork(shape, cond, x_0=full.clone())
    mine = O.ddim_sample(sd, sched, shape, cond, full, bank, footwork=True)
    md_foot = float((ref_foot - mine).abs().max())
    sp = 8
    bank2 = synth.make_noise_bank(shape, sp, seed=65)
    with ref_shim.NoiseBank(bank2[1:]):
        ref_li = diff.long_inpaint_loop(shape, cond, noise=bank2[0].clone(), start_point=sp)
    mine = O.p_sample_loop(sd, sched, shape, cond, bank2, start_point=sp, long_mode=True)
    md_li = float((ref_li - mine).abs().max())
    print(f"  {name} variants: oracle vs reference max|d| long {md_long:.3e} foot {md_foot:.3e} long_inpaint {md_li:.3e}")
    assert max(md_long, md_foot, md_li) < 1e-3
    save(f"{name}_variants.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd), "B": B,
                                 "row_stride": 2, "long_ddim": ref_long[:, ::2].clone(), "footwork": ref_foot[:, ::2].clone(),
                                 "long_inpaint": ref_li[:, ::2].clone(), "start_point": sp,
                                 "oracle_maxdiff": max(md_long, md_foot, md_li)})


def gen_inpaint(name="tiny", B=2, start_point=6):
    """inpaint_loop of the unmodified reference (model/diffusion.py:518-557) with every randn_like pre-drawn: the bank holds
    x_T, then per step the p_sample draw followed (i > 0) by the q_sample draw of the constraint value."""
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    _, diff = build_reference(cfg, sd)
    shape = (B, 150 * cfg["dancers"], 151)
    cond = synth.make_music(B, cfg["cond_feature_dim"], seed=72)
    n_draws = 2 * start_point - 1
    bank = synth.make_noise_bank(shape, n_draws, seed=73)
    con = synth.make_inpaint_constraint(shape)
    with ref_shim.NoiseBank(bank[1:]) as nb:
        ref = diff.inpaint_loop(shape, cond, noise=bank[0].clone(), constraint={k: v.clone() for k, v in con.items()},
                                start_point=start_point)
        assert nb.i == n_draws, (nb.i, n_draws)
    mine = O.p_sample_loop(sd, O.make_schedule("cosine", 1000), shape, cond, bank, start_point=start_point, constraint=con)
    md = float((ref - mine).abs().max())
    # the reference's p_sample_loop ignores `constraint` (model/diffusion.py:255-286): same bank prefix, plain sampling
    with ref_shim.NoiseBank(bank[1:]) as nb:
        ref_plain = diff.p_sample_loop(shape, cond, noise=bank[0].clone(), constraint={k: v.clone() for k, v in con.items()},
                                       start_point=start_point)
        assert nb.i == start_point
    print(f"  {name} inpaint_loop last-{start_point}: oracle vs reference max|d| = {md:.3e}")
    assert md < 1e-4
    save(f"{name}_inpaint.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd), "B": B,
                                "noise_seed": 73, "cond_seed": 72, "constraint_seed": 71, "start_point": start_point,
                                "n_draws": n_draws, "out": ref.clone(), "out_p_sample_loop_ignores_constraint": ref_plain.clone(),
                                "oracle_maxdiff": md})


def gen_plosses(name="tiny", B=3):
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    _, diff = build_reference(cfg, sd)
    dn = cfg["dancers"]
    x = synth.make_motion(B, dn, seed=42)
    cond = synth.make_music(B, cfg["cond_feature_dim"], seed=43)
    t = torch.tensor([3, 500, 987])[:B]
    keep = torch.tensor([True, False, True])[:B]
    noise = torch.randn(B, 150, dn, 151, generator=torch.Generator().manual_seed(44))
    with torch.no_grad(), ref_shim.NoiseBank([noise], keep_mask=keep):
        tot, parts = diff.p_losses(x.clone(), cond, t)
    mtot, mparts = O.p_losses(sd, O.make_schedule("cosine", 1000), x, cond, t, noise, keep)
    ref = torch.stack([tot] + list(parts))
    mine = torch.stack([mtot] + list(mparts))
    md = float(((ref - mine).abs() / ref.abs().clamp_min(1e-12)).max())
    print(f"  {name} p_losses: ref {ref.tolist()}  rel diff {md:.3e}")
    assert md < 1e-4
    save(f"{name}_plosses.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd),
                                "B": B, "t": t, "keep_mask": keep, "losses": ref, "oracle_maxdiff_rel": md})


def gen_loss_terms(B=3, dn=3):
    """The four loss terms for a GIVEN prediction: the reference's network is replaced by a stub that
    returns a fixed tensor, so the foot-contact branch (contact > 0.95) is exercised."""
    ns = ref_shim.load()

    class Stub(torch.nn.Module):
        def forward(self, x, cond, t, cond_drop_prob=0.0, trj_dist=None):
            return self.pred

    stub = Stub()
    diff = ns.GaussianDiffusion(stub, 150, 151, ns.SMPLSkeleton(None), schedule="cosine", n_timestep=1000,
                                predict_epsilon=False, loss_type="l2", use_p2=False, cond_drop_prob=0.25,
                                guidance_weight=2).eval()
    target = synth.make_motion(B, dn, seed=50)                       # (B,dn,S,151)
    pred = synth.make_prediction(B, dn, seed=51)
    stub.pred = pred
    t = torch.tensor([0, 400, 999])[:B]
    with torch.no_grad():
        tot, parts = diff.p_losses(target.clone(), torch.zeros(B, 301, 4), t)
    ref = torch.stack([tot] + list(parts))
    xs = target.permute(0, 2, 1, 3)
    mtot, mparts = O.loss_terms(pred.reshape(B, 150, dn, 151), xs)
    mine = torch.stack([mtot] + list(mparts))
    md = float(((ref - mine).abs() / ref.abs()).max())
    print(f"  loss terms: ref {ref.tolist()} rel diff {md:.3e}")
    assert md < 1e-5 and float(ref[4]) > 0
    save("loss_terms.pt", {"B": B, "dn": dn, "target_seed": 50, "pred_seed": 51, "losses": ref,
                           "oracle_maxdiff_rel": md})


def gen_fk():
    ns = ref_shim.load()
    g = torch.Generator().manual_seed(9)
    d6 = torch.randn(2, 20, 24, 6, generator=g)
    d6[0, 0, 0] = torch.tensor([1.0, 0, 0, 0, 1.0, 0])          # identity
    d6[0, 0, 1] = torch.tensor([-1.0, 0, 0, 0, -1.0, 0])        # 180 deg about z (w == 0 branch)
    d6[0, 0, 2] = torch.tensor([1.0, 0, 0, 0, -1.0, 0])         # 180 deg about x
    root = torch.randn(2, 20, 3, generator=g)
    aa = ns.ax_from_6v(d6)
    pos = ns.SMPLSkeleton(None).forward(aa, root)
    md = max(float((aa - O.ax_from_6v(d6)).abs().max()), float((pos - O.smpl_forward(aa, root)).abs().max()))
    assert md == 0.0, md
    zero_pose = ns.SMPLSkeleton(None).forward(torch.zeros(1, 1, 24, 3), torch.zeros(1, 1, 3))
    save("fk.pt", {"seed": 9, "d6": d6, "root": root, "axis_angle": aa, "positions": pos,
                   "zero_pose": zero_pose, "oracle_maxdiff": md})


def gen_adan(steps=5):
    """5 steps of the reference's Adan (lr 4e-4, wd 0.02, TCDiff.py:45-46,110) + EMA(0.9999) on a small parameter
    set with seeded gradients; shapes include sizes that are not multiples of 4 and a parameter without gradient."""
    ref_shim.load()
    import importlib
    RefAdan = importlib.import_module("model.adan").Adan
    RefEMA = importlib.import_module("model.diffusion").EMA
    g = torch.Generator().manual_seed(77)
    shapes = [(151, 7), (513,), (64, 64), (3,), (5, 2)]
    p0 = [torch.randn(s, generator=g) for s in shapes]
    params = [torch.nn.Parameter(p.clone()) for p in p0]
    ma = [torch.nn.Parameter(p.clone()) for p in p0]

    class Holder:
        def __init__(self, ps):
            self.ps = ps

        def parameters(self):
            return iter(self.ps)

    opt = RefAdan(params, lr=4e-4, weight_decay=0.02)
    ema = RefEMA(0.9999)
    all_grads, trace = [], []
    for it in range(steps):
        grads = [torch.randn(s, generator=g) * (10.0 ** (it - 2)) for s in shapes]
        grads[4] = None
        for p, gr in zip(params, grads):
            p.grad = None if gr is None else gr.clone()
        opt.step()
        ema.update_model_average(Holder(ma), Holder(params))
        all_grads.append(grads)
        trace.append([p.detach().clone() for p in params])
    mine = [p.clone() for p in p0]
    mine_ma = [p.clone() for p in p0]
    st = O.adan_init(mine)
    for grads in all_grads:
        O.adan_step(mine, grads, st, lr=4e-4, weight_decay=0.02)
        O.ema_update(mine_ma, mine, 0.9999)
    md = max(float((a.detach() - b).abs().max()) for a, b in zip(params + ma, mine + mine_ma))
    assert md == 0.0, md
    save("adan.pt", {"seed": 77, "lr": 4e-4, "weight_decay": 0.02, "ema_beta": 0.9999, "p0": p0, "grads": all_grads,
                     "params_trace": trace, "ema_final": [p.detach().clone() for p in ma],
                     "state_final": [{k: (v.clone() if torch.is_tensor(v) else v) for k, v in opt.state[p].items()}
                                     for p in params], "oracle_maxdiff": md})


def gen_post(dn=3):
    """Post-sampling stage through the reference's OWN render_sample (model/diffusion.py:765-986) with
    skeleton_render intercepted: normal mode (b=2 clips) and long mode (b=3 windows of one song)."""
    import pickle
    import tempfile
    ns = ref_shim.load()
    Normalizer = __import__("importlib").import_module("dataset.preprocess").Normalizer
    g = torch.Generator().manual_seed(31)
    fit = torch.randn(4000, 151, generator=g) * (torch.rand(151, generator=g) * 3 + 0.2) + torch.randn(151, generator=g)
    normalizer = Normalizer(fit.clone())
    min_, scale_ = normalizer.scaler.min_.clone(), normalizer.scaler.scale_.clone()
    cfg = synth.CONFIGS["tiny"]
    diff = ns.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, ns.SMPLSkeleton(torch.device("cpu")), schedule="cosine",
                                n_timestep=1000, predict_epsilon=False, loss_type="l2")
    captured = []
    md = ns.model_diffusion
    old, old_map = md.skeleton_render, md.p_map
    md.skeleton_render = lambda poses, **kw: captured.append((poses, kw))
    md.p_map = lambda f, it: [f(a) for a in it]                       # in-process (the real p_map forks workers)
    out = {"min_": min_, "scale_": scale_, "dn": dn, "seed": 31}
    try:
        for mode, b in (("normal", 2), ("long", 3)):
            captured.clear()
            samples = (torch.rand(b, 150 * dn, 151, generator=g) * 2.4 - 1.2)
            with tempfile.TemporaryDirectory() as td:
                names = [f"data/test/features/clip{i}_x.npy" for i in range(b)]
                diff.render_sample(samples.clone(), torch.zeros(1), normalizer, 0, td, fk_out=td, name=names, sound=False,
                                   mode=mode, render=False, required_dancer_num=dn)
                pk = [pickle.load(open(os.path.join(td, f), "rb")) for f in sorted(os.listdir(td)) if f.endswith(".pkl")]
            ref = {"samples": samples}
            if mode == "normal":
                ref["full_pose"] = torch.stack([torch.as_tensor(c[0]) for c in captured])            # (b, dn, 150, 24, 3)
                ref["contact"] = torch.stack([torch.as_tensor(c[1]["contact"]) for c in captured])   # (b, dn, 150, 4)
                ref["smpl_poses"] = torch.stack([torch.as_tensor(p["smpl_poses"]) for p in pk])      # (b, 150*dn, 72)
                ref["smpl_trans"] = torch.stack([torch.as_tensor(p["smpl_trans"]) for p in pk])
            else:
                ref["full_pose"] = torch.as_tensor(pk[0]["full_pose"])                               # (dn, F, 24, 3)
                ref["smpl_poses"] = torch.as_tensor(pk[0]["smpl_poses"])                             # (F*dn, 72)
                ref["smpl_trans"] = torch.as_tensor(pk[0]["smpl_trans"])                             # (1?, F*dn, 3)
            mine = O.samples_to_poses(samples, min_, scale_, dn, mode)
            md_ = 0.0
            for k in ("full_pose", "smpl_poses", "smpl_trans") + (("contact",) if mode == "normal" else ()):
                md_ = max(md_, float((mine[k].reshape(-1) - ref[k].reshape(-1).float()).abs().max()))
            assert md_ < 1e-4, (mode, md_)
            ref["oracle_maxdiff"] = md_
            out[mode] = ref
    finally:
        md.skeleton_render, md.p_map = old, old_map
    save("post.pt", out)


def load_ref_trajdecoder():
    """The reference's TrajDecoder class (TrajDecoder/model/traj_model.py), imported as-is."""
    import importlib.util
    ref_shim.load()                                    # puts /root/reference on sys.path ('model.utils' = PositionalEncoding)
    path = os.path.join(ref_shim.REFERENCE_ROOT, "TrajDecoder", "model", "traj_model.py")
    spec = importlib.util.spec_from_file_location("ref_traj_model", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.TrajDecoder


def gen_traj(b=3, dn=2, window=20, step=5, Fm=12, layers=2):
    """Reduced TrajDecoder (same code paths: 3-layer LSTM over the batch axis, hidden 64, 4 heads of 32) through the
    reference module: one forward and the sliding-window generation loop of TCDiff.py:526-546."""
    from oracle import traj_oracle as TO
    TrajDecoder = load_ref_trajdecoder()
    torch.manual_seed(7)
    m = TrajDecoder(nfeats=2, trans_layer=layers, window_size=window, cond_feature_dim=Fm).eval()
    with torch.no_grad():                              # default init leaves biases at zero: make every term live
        for p in m.parameters():
            if p.dim() == 1:
                p.add_(torch.randn_like(p) * 0.05)
    sd = {k: v.clone() for k, v in m.state_dict().items()}
    g = torch.Generator().manual_seed(8)
    S = 2 * window                                     # total frames: first window + 4 predicted steps
    x = torch.randn(b, dn, S, 2, generator=g)
    cond = torch.randn(b, 2 * S + 1, Fm, generator=g)
    with torch.no_grad():
        one = m(x[:, :, :window], cond[:, :(window + step) * 2])
        cond_traj = x[:, :, :window]
        pre = [cond_traj]
        for start in range(0, cond.shape[1] + 1 - (window + step) * 2, step * 2):
            cond_traj = m(cond_traj, cond[:, start:start + (window + step) * 2])
            pre.append(cond_traj[:, :, -step:])
        full = torch.cat(pre, dim=2)
        mine_one = TO.traj_decoder_forward(sd, x[:, :, :window], cond[:, :(window + step) * 2])
        mine_full = TO.generate_trajectory(sd, x, cond, window, step)
    md = max(float((one - mine_one).abs().max()), float((full - mine_full).abs().max()))
    assert md < 1e-5, md
    sm = TO.kalman_smooth_batch(full.numpy())
    save("traj.pt", {"seed": 7, "cfg": dict(nfeats=2, trans_layer=layers, window_size=window, cond_feature_dim=Fm, step=step),
                     "state_dict": sd, "x": x, "cond": cond, "forward": one, "trajectory": full,
                     "kalman": torch.from_numpy(sm), "kalman_gains": torch.from_numpy(TO.kalman_gains(full.shape[2])),
                     "oracle_maxdiff": md})


def gen_defaults(name="tiny", B=3, start_point=12):
    """The reference CONSTRUCTOR DEFAULTS (model/diffusion.py:86-96: loss_type "l1", predict_epsilon True, guidance_weight 3,
    cond_drop_prob 0.2) — TCDiff.py never uses them, a bare GaussianDiffusion(model, horizon, repr_dim, smpl) does:
    p_losses (target = noise, L1 terms), the four L1 terms for a given prediction (foot branch live), and the last
    `start_point` ancestral steps with x_recon = predict_start_from_noise(x, t, out)."""
    ns = ref_shim.load()
    cfg = synth.CONFIGS[name]
    sd = synth.make_state_dict(cfg, 0)
    m, _ = build_reference(cfg, sd)
    diff = ns.GaussianDiffusion(m, cfg["seq_len"], cfg["nfeats"], ns.SMPLSkeleton(None), schedule="cosine", n_timestep=1000).eval()
    assert diff.predict_epsilon and diff.loss_fn is F.l1_loss and diff.guidance_weight == 3 and diff.cond_drop_prob == 0.2
    sched = O.make_schedule("cosine", 1000)
    dn = cfg["dancers"]
    # ---- p_losses
    x = synth.make_motion(B, dn, seed=42)
    cond = synth.make_music(B, cfg["cond_feature_dim"], seed=43)
    t = torch.tensor([3, 500, 987])[:B]
    keep = torch.tensor([True, False, True])[:B]
    noise = torch.randn(B, 150, dn, 151, generator=torch.Generator().manual_seed(44))
    with torch.no_grad(), ref_shim.NoiseBank([noise], keep_mask=keep):
        tot, parts = diff.p_losses(x.clone(), cond, t)
    mtot, mparts = O.p_losses(sd, sched, x, cond, t, noise, keep, loss_type="l1", predict_epsilon=True)
    ref = torch.stack([tot] + list(parts))
    mine = torch.stack([mtot] + list(mparts))
    md1 = float(((ref - mine).abs() / ref.abs().clamp_min(1e-12)).max())
    print(f"  defaults p_losses: ref {ref.tolist()}  rel diff {md1:.3e}")
    assert md1 < 1e-4
    # ---- loss terms for a given prediction
    class Stub(torch.nn.Module):
        def forward(self, x, cond, t, cond_drop_prob=0.0, trj_dist=None):
            return self.pred
    stub = Stub()
    dstub = ns.GaussianDiffusion(stub, 150, 151, ns.SMPLSkeleton(None), schedule="cosine", n_timestep=1000).eval()
    target = synth.make_motion(B, 3, seed=50)
    pred = synth.make_prediction(B, 3, seed=51)
    stub.pred = pred
    nz = torch.randn(B, 150, 3, 151, generator=torch.Generator().manual_seed(52)).clamp(-1, 1)   # the target of this variant
    with torch.no_grad(), ref_shim.NoiseBank([nz]):
        tot2, parts2 = dstub.p_losses(target.clone(), torch.zeros(B, 301, 4), torch.tensor([0, 400, 999])[:B])
    ref2 = torch.stack([tot2] + list(parts2))
    mt2, mp2 = O.loss_terms(pred.reshape(B, 150, 3, 151), nz, None, "l1")
    mine2 = torch.stack([mt2] + list(mp2))
    md2 = float(((ref2 - mine2).abs() / ref2.abs()).max())
    print(f"  defaults loss terms: ref {ref2.tolist()} rel diff {md2:.3e}")
    assert md2 < 1e-5 and float(ref2[4]) > 0
    # ---- ancestral sampling with predict_start_from_noise
    shape = (2, 150 * dn, 151)
    cond2 = synth.make_music(2, cfg["cond_feature_dim"])
    bank = synth.make_noise_bank(shape, start_point, seed=778)
    with ref_shim.NoiseBank(bank[1:]) as nb:
        refs = diff.p_sample_loop(shape, cond2, noise=bank[0].clone(), start_point=start_point)
        assert nb.i == start_point
    mines = O.p_sample_loop(sd, sched, shape, cond2, bank, guidance_weight=3, start_point=start_point, predict_epsilon=True)
    md3 = float((refs - mines).abs().max())
    print(f"  defaults ddpm last-{start_point}: oracle vs reference max|d| = {md3:.3e}")
    assert md3 < 1e-4
    save(f"{name}_defaults.pt", {"config": name, "weight_seed": 0, "weight_checksum": synth.weight_checksum(sd), "B": B, "t": t,
                                 "keep_mask": keep, "losses": ref, "terms_noise_seed": 52, "terms_losses": ref2,
                                 "ddpm_noise_seed": 778, "start_point": start_point, "ddpm_out": refs.clone(),
                                 "oracle_maxdiff": max(md1, md2, md3)})


def main():
    torch.set_num_threads(os.cpu_count())
    torch.manual_seed(0)
    assert ref_shim.available(), "needs /root/reference"
    print("schedule"); gen_schedule()
    print("fk"); gen_fk(); gen_loss_terms()
    print("adan"); gen_adan()
    print("post"); gen_post()
    print("traj"); gen_traj()
    print("forward"); gen_forward("tiny"); gen_forward("c1")
    print("p_losses"); gen_plosses()
    print("ddpm"); gen_ddpm()
    print("defaults"); gen_defaults()
    print("inpaint"); gen_inpaint()
    print("variants"); gen_variants()
    print("ddim"); gen_ddim("tiny"); gen_ddim("c1")


if __name__ == "__main__":
    sys.exit(main())
