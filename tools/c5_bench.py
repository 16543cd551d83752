"""BASELINE.json configs[4] (c5): end-to-end test mode — TrajDecoder trajectory generation + Kalman smoothing feeding
TCDiff DDIM-50 with classifier-free guidance and the post-sampling stage (un-normalise, 6D -> axis-angle, SMPL FK);
batch 256 over 8 GPUs = 32 clips per GPU (one process per GPU, no data-path collective until the final gather).

    python tools/c5_bench.py [--batch 32] [--steps 3]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29571 tools/c5_bench.py
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=32)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=2)
    ap.add_argument("--lib", default=None, help="an A/B build of the library")
    a = ap.parse_args()
    if a.lib:
        from tcdiff_b200 import _lib as _tlib
        _tlib.LIB_PATH = os.path.abspath(a.lib)
    import torch.distributed as dist
    import tcdiff_b200 as T
    from tcdiff_b200 import synth                         # synthetic weights / inputs only
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS["c2"]
    S, dn, Fm = cfg["seq_len"], cfg["dancers"], cfg["cond_feature_dim"]
    m = T.DanceDecoder(nfeats=151, seq_len=S, latent_dim=512, ff_size=cfg["ff_size"], num_layers=cfg["num_layers"], num_heads=8,
                       cond_feature_dim=Fm, required_dancer_num=dn, dtype="bf16")
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, S, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", cond_drop_prob=0.25, guidance_weight=2).to(dev)
    torch.manual_seed(42)                            # option_traj.py:63
    traj = T.TrajDecoder(nfeats=2, trans_layer=6, window_size=100).to(dev).eval()      # option_traj.py:33-36
    B = a.batch
    gen = torch.Generator(device=dev).manual_seed(500 + rank)
    x = torch.rand(B, dn, S, 151, device=dev, generator=gen) * 2 - 1
    cond = torch.randn(B, 2 * S + 1, Fm, device=dev, generator=gen)
    norm = (torch.zeros(151, device=dev), torch.ones(151, device=dev))                 # identity MinMax scaler
    shape = (B, S * dn, 151)
    ev = lambda: torch.cuda.Event(enable_timing=True)
    t_front = t_sample = t_post = 0.0

    def step(timed):
        nonlocal t_front, t_sample, t_post
        e = [ev() for _ in range(4)]
        e[0].record()
        x_traj = T.generate_trajectory(traj, x, cond, 100, 25)                         # TCDiff.py:526-556
        x0 = x_traj.permute(0, 2, 1, 3).reshape(B, S * dn, 3).contiguous()
        e[1].record()
        samples = d.ddim_sample(shape, cond, x_0=x0)
        e[2].record()
        out = d.samples_to_poses(samples, norm, mode="normal", required_dancer_num=dn)
        e[3].record()
        if timed:
            torch.cuda.synchronize()
            t_front += e[0].elapsed_time(e[1]); t_sample += e[1].elapsed_time(e[2]); t_post += e[2].elapsed_time(e[3])
        return out["full_pose"]

    for _ in range(a.warmup):
        step(False)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = ev(), ev()
    e0.record()
    for _ in range(a.steps):
        poses = step(True)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms) / a.steps
    if rank == 0:
        print(json.dumps({"metric": "end-to-end clips/sec (TrajDecoder + Kalman -> DDIM-50 CFG -> FK joint positions)",
                          "unit": "clips/s", "value": world * B / (ms * 1e-3), "n_gpus": world, "batch_per_gpu": B,
                          "ms_per_call": ms, "ms_front_end": t_front / a.steps, "ms_sampler": t_sample / a.steps,
                          "ms_post": t_post / a.steps, "finite": bool(torch.isfinite(poses).all()),
                          "config": {"workload": f"c5: batch {B}/GPU x{world}, {dn} dancers, {S} frames, TrajDecoder 6 layers "
                                                 f"window 100 step 25"}}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
