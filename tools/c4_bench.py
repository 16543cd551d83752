"""BASELINE.json configs[3]: Jukebox-style 4800-dim music conditioning, 10 dancers, 300 frames, full 1000-step DDPM
sampling, batch-sharded (one process per GPU, `--batch` clips each, one final all_gather).

    python tools/c4_bench.py [--batch 1] [--steps 1000]
    torchrun --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29561 tools/c4_bench.py

`--steps N` times the LAST N steps of the 1000-step chain (start_point=N, the reference's own way of starting late,
model/diffusion.py:263-270) and scales to 1000; default = the full chain.  Prints one JSON line on rank 0.
"""
import argparse
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--batch", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000)
    ap.add_argument("--dtype", default="bf16")
    ap.add_argument("--lib", default=None, help="an A/B build of the library")
    a = ap.parse_args()
    if a.lib:
        from tcdiff_b200 import _lib as _tlib
        _tlib.LIB_PATH = os.path.abspath(a.lib)
    import torch.distributed as dist
    import tcdiff_b200 as T
    from tcdiff_b200 import synth
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    cfg = synth.CONFIGS["c4"]
    S, dn, Fm = cfg["seq_len"], cfg["dancers"], cfg["cond_feature_dim"]
    m = T.DanceDecoder(nfeats=151, seq_len=S, latent_dim=512, ff_size=cfg["ff_size"], num_layers=cfg["num_layers"], num_heads=8,
                       cond_feature_dim=Fm, required_dancer_num=dn, dtype=a.dtype)
    m.load_state_dict(synth.make_state_dict(cfg, 0))
    m = m.to(dev).eval()
    d = T.GaussianDiffusion(m, S, 151, T.SMPLSkeleton(dev), schedule="cosine", n_timestep=1000, predict_epsilon=False,
                            loss_type="l2", cond_drop_prob=0.25, guidance_weight=2, seq_len=S).to(dev)
    B = a.batch
    shape = (B, S * dn, 151)
    gen = torch.Generator(device=dev).manual_seed(99 + rank)
    cond = torch.randn(B, 2 * S + 1, Fm, device=dev, generator=gen)
    sp = None if a.steps >= 1000 else a.steps
    run = lambda: d.p_sample_loop(shape, cond, start_point=sp)
    out = run()                                                     # warm-up: packs weights, captures the step graphs
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    out = run()
    if world > 1:
        parts = [torch.empty_like(out) for _ in range(world)]
        dist.all_gather(parts, out)
    e1.record()
    torch.cuda.synchronize()
    ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    ms = float(ms)
    steps = min(a.steps, 1000)
    if rank == 0:
        full_ms = ms * 1000.0 / steps
        print(json.dumps({"metric": "10 s 10-dancer clips/sec (DDPM-1000, Jukebox-style 4800-dim music)", "unit": "clips/s",
                          "value": world * B / (full_ms * 1e-3), "n_gpus": world, "batch_per_gpu": B,
                          "ms_per_denoise_step": ms / steps, "timed_steps": steps, "seconds_per_1000_step_call": full_ms * 1e-3,
                          "dtype": a.dtype, "finite": bool(torch.isfinite(out).all()),
                          "config": {"workload": f"c4: {dn} dancers, {S} frames, music {Fm}-dim, DDPM-1000, cfg guidance, "
                                                 f"batch-sharded x{world}"},
                          "peak_mem_gb": torch.cuda.max_memory_allocated() / 2 ** 30}))
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
