"""Per-parameter gradient cosines of the bf16 training tape against autograd through the CPU oracle (tiny config, dropout 0),
for the loaded library (A/B aid): prints the lowest cosines and the whole-gradient cosine, for several input seeds."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tcdiff_b200 import _lib
if "--lib" in sys.argv:
    _lib.LIB_PATH = os.path.abspath(sys.argv[sys.argv.index("--lib") + 1])
import tcdiff_b200 as T
from oracle import synth, tcdiff_oracle as O
dev = torch.device("cuda:0")
print("library", _lib.LIB_PATH, "attn_2q", _lib.lib().tcd_tuning(b"attn_2q"))
cfg = synth.CONFIGS["tiny"]
sd = synth.make_state_dict(cfg, 0)
for seed in (42, 142, 242):
    m = T.DanceDecoder(nfeats=151, seq_len=150, latent_dim=512, ff_size=cfg["ff_size"], num_layers=cfg["num_layers"], num_heads=8,
                       dropout=0.0, cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"], dtype="bf16")
    m.load_state_dict(sd)
    m = m.to(dev).train()
    d = T.GaussianDiffusion(m, 150, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", use_p2=False, cond_drop_prob=0.25, guidance_weight=2).to(dev)
    B, dn = 2, cfg["dancers"]
    x = synth.make_motion(B, dn, seed=seed)
    cond = synth.make_music(B, cfg["cond_feature_dim"], seed=seed + 1)
    t = torch.tensor([3, 700])
    keep = torch.tensor([True, False])
    noise = torch.randn(B, 150, dn, 151, generator=torch.Generator().manual_seed(seed + 2))
    tot, _ = d.p_losses(x.to(dev), cond.to(dev), t.to(dev), noise=noise.to(dev), keep_mask=keep.to(dev))
    tot.backward()
    sdg = {k: v.clone().requires_grad_(v.dtype.is_floating_point) for k, v in sd.items()}
    otot, _ = O.p_losses(sdg, O.make_schedule("cosine", 1000), x, cond, t, noise, keep)
    otot.backward()
    cos, num, na, nb = {}, 0.0, 0.0, 0.0
    for name, p in m.named_parameters():
        gr = sdg[name].grad
        if gr is None or float(gr.abs().max()) == 0.0:
            continue
        g = p.grad.cpu().double(); r = gr.double()
        cos[name] = float((g * r).sum() / (g.norm() * r.norm()))
        num += float((g * r).sum()); na += float((g * g).sum()); nb += float((r * r).sum())
    # the foot-skate term masks velocities with the DISCRETE indicator model_contact > 0.95 (model/diffusion.py:723,731):
    # count the indicator entries on which the bf16 forward and the fp32 oracle disagree
    with torch.no_grad():
        x_noisy = O.q_sample(O.make_schedule("cosine", 1000), x.permute(0, 2, 1, 3), t, noise)       # (B, S, dn, C)
        x_noisy[..., 4:6] = x.permute(0, 2, 1, 3)[..., 4:6]
        xn = x_noisy.reshape(B, 150 * dn, 151)
        want = O.dance_decoder_forward(sd, xn, cond, t, keep_mask=keep)
        got = m.eval()(xn.to(dev), cond.to(dev), t.to(dev), keep_mask=keep.to(dev)).cpu()
        m.train()
    cw, cg = want[..., :4], got[..., :4]
    flips = int(((cw > 0.95) != (cg > 0.95)).sum())
    near = int(((cw - 0.95).abs() < 0.02).sum())
    print(f"   contact indicator: {int((cw > 0.95).sum())} of {cw.numel()} set, {near} within 0.02 of the threshold, {flips} flipped by the bf16 forward")
    # conditioning of the loss itself: dL/d(out) from the ORACLE's loss evaluated at the oracle's output and at the bf16 output
    tgt = x.permute(0, 2, 1, 3).reshape(B, 150, dn, 151)
    p2w = O.make_schedule("cosine", 1000)["p2_loss_weight"].gather(-1, t)
    gs = []
    for o_ in (want, got):
        oo = o_.clone().reshape(B, 150, dn, 151).requires_grad_(True)
        tt_, parts_ = O.loss_terms(oo, tgt, p2w)
        tt_.backward()
        gs.append(oo.grad.double().flatten())
    per = [float((gs[0][i::151] * gs[1][i::151]).sum() / (gs[0][i::151].norm() * gs[1][i::151].norm() + 1e-300)) for i in (0, 4, 7, 150)]
    print(f"   out rel-L2 {float((got - want).norm() / want.norm()):.3e}; dL/dout cosine (oracle loss at fp32 out vs at bf16 out): "
          f"{float((gs[0] * gs[1]).sum() / (gs[0].norm() * gs[1].norm())):.5f}; |dL/dout| {float(gs[0].norm()):.3e}; per-channel cos (0,4,7,150): {per}")
    low = sorted(cos.items(), key=lambda kv: kv[1])[:6]
    print(f"seed {seed}: loss {float(tot):.5f} vs {float(otot):.5f}; whole {num / (na * nb) ** 0.5:.5f}; median {sorted(cos.values())[len(cos) // 2]:.5f}; lowest:",
          ", ".join(f"{k.replace('seqTransDecoder.stack.', 'L')}={v:.4f}" for k, v in low))
