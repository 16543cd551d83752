"""A/B of engine.TAIL_CHUNK (row chunking of the fc / linear2 -> tail pairs so that y stays in L2) on the c2 sampler:
same seed, so the samples must be bit-identical; clips/s per variant.  usage: python tools/tail_chunk_ab.py 0 25 12"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402

import tcdiff_b200 as T  # noqa: E402
from tcdiff_b200 import engine  # noqa: E402
from tcdiff_b200 import synth  # noqa: E402  (synthetic weights / inputs only)

dev = torch.device("cuda:0")
cfg = synth.CONFIGS["c2"]
B = 64
shape = (B, cfg["seq_len"] * cfg["dancers"], 151)
cond = synth.make_music(B, cfg["cond_feature_dim"], seed=1235).to(dev)
x0 = synth.make_traj(synth.make_motion(B, cfg["dancers"], seed=1234)).to(dev)
sd = synth.make_state_dict(cfg, 0)
res, base = {}, None
for chunk in [int(a) for a in sys.argv[1:]] or [0, 25]:
    engine.TAIL_CHUNK = chunk
    m = T.DanceDecoder(nfeats=151, seq_len=cfg["seq_len"], latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"],
                       num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], dropout=0.1,
                       cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"], dtype="bf16")
    m.load_state_dict(sd)
    d = T.GaussianDiffusion(m.to(dev).eval(), cfg["seq_len"], 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000,
                            predict_epsilon=False, loss_type="l2", use_p2=False, cond_drop_prob=0.25,
                            guidance_weight=2).to(dev).eval()
    torch.manual_seed(7)
    out = d.ddim_sample(shape, cond, x_0=x0).clone()
    d.ddim_sample(shape, cond, x_0=x0)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(3):
        d.ddim_sample(shape, cond, x_0=x0)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 3
    base = out if base is None else base
    res[str(chunk)] = {"clips_per_s": B / (ms * 1e-3), "ms_per_denoise_step": ms / 50, "identical_to_first": bool(torch.equal(out, base))}
    print(chunk, res[str(chunk)], flush=True)
    json.dump(res, open(os.path.join(ROOT, "gpurun_out", "tail_chunk_ab.json"), "w"), indent=1)
    del d, m
