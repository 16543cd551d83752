"""Deterministic synthetic weights, inputs and noise banks (pure torch CPU generators; no kernels, no checker code).

Used by bench.py / tools/ (the measured arm needs random-init weights of the reference's architecture and synthetic
inputs of BASELINE.json's shapes) and, through an alias in the CPU checker's package, by the tests and the golden-fixture
generator, so that both sides of every parity check see the same bits.

The reference ships no checkpoints, datasets or golden vectors, and a 60-117 M
parameter state_dict cannot be committed, so weights are regenerated from a seed on every
machine (torch CPU generators are bit-reproducible for a fixed torch version; every golden
fixture stores a weight checksum that the tests re-check).  Key names/shapes follow
model/model.py:417-540 and are verified against the real reference in
the `not gpu` tests that import the unmodified reference.
"""
import torch

CONFIGS = {
    # BASELINE.json configs: c1 CPU-runnable, c2 headline, c4 Jukebox-style
    "c1": dict(nfeats=151, seq_len=150, latent_dim=512, ff_size=1024, num_layers=8, num_heads=8,
               cond_feature_dim=35, dancers=3),
    "c2": dict(nfeats=151, seq_len=150, latent_dim=512, ff_size=1024, num_layers=8, num_heads=8,
               cond_feature_dim=438, dancers=5),
    "c4": dict(nfeats=151, seq_len=300, latent_dim=512, ff_size=1024, num_layers=8, num_heads=8,
               cond_feature_dim=4800, dancers=10),
    # reduced model for fast CPU/GPU parity loops (same code paths, every odd size kept:
    # 151 channels, L=300 not a multiple of 128, memory length 152, 2*Fm=26).  latent_dim must
    # stay 512 (the reference's dead traj_Modulation hard-codes context_dim=512, model.py:256,301)
    # and heads must stay 8 so that the music encoder's head dim is 64 like the decoder's.
    "tiny": dict(nfeats=151, seq_len=150, latent_dim=512, ff_size=256, num_layers=2, num_heads=8,
                 cond_feature_dim=13, dancers=2),
}


def state_dict_spec(cfg):
    """[(key, shape, kind)] in a fixed order; kind in {w, b, ln_w, ln_b, randn, freqs}."""
    D, FF, Fm, dn, S = cfg["latent_dim"], cfg["ff_size"], cfg["cond_feature_dim"], cfg["dancers"], cfg["seq_len"]
    H = cfg["num_heads"]
    nf = cfg["nfeats"]
    spec = []

    def lin(name, out, inp, bias=True):
        spec.append((name + ".weight", (out, inp), "w"))
        if bias:
            spec.append((name + ".bias", (out,), "b"))

    def ln(name, d=D):
        spec.append((name + ".weight", (d,), "ln_w"))
        spec.append((name + ".bias", (d,), "ln_b"))

    spec.append(("null_cond_embed", (1, S, D), "randn"))
    spec.append(("null_cond_hidden", (1, D), "randn"))
    spec.append(("rotary.freqs", (D // 2,), "freqs"))
    lin("time_mlp.1", 4 * D, D)
    lin("to_time_cond.0", D, 4 * D)
    lin("to_time_tokens.0", 2 * D, 4 * D)
    ln("norm_cond")
    lin("input_projection", D, nf)
    for i in range(2):
        p = f"cond_encoder.{i}"
        spec.append((p + ".self_attn.in_proj_weight", (3 * D, D), "w"))
        spec.append((p + ".self_attn.in_proj_bias", (3 * D,), "b"))
        lin(p + ".self_attn.out_proj", D, D)
        lin(p + ".linear1", FF, D)
        lin(p + ".linear2", D, FF)
        ln(p + ".norm1")
        ln(p + ".norm2")
        spec.append((p + ".rotary.freqs", (D // 2,), "freqs"))
    lin("cond_projection.0", Fm, 2 * Fm)
    lin("cond_projection.2", D, Fm)
    ln("non_attn_cond_projection.0")
    lin("non_attn_cond_projection.1", D, D)
    lin("non_attn_cond_projection.3", D, D)
    for i in range(cfg["num_layers"]):
        p = f"seqTransDecoder.stack.{i}"
        for att in ("self_attn", "multihead_attn"):
            for w in ("w_qs", "w_ks", "w_vs"):
                lin(f"{p}.{att}.{w}", H * 64, D, bias=False)
            lin(f"{p}.{att}.fc", D, H * 64, bias=False)
            ln(f"{p}.{att}.layer_norm")
        lin(p + ".linear1", FF, D)
        lin(p + ".linear2", D, FF)
        for n in ("norm1", "norm2", "norm3"):
            ln(f"{p}.{n}")
        for f in ("film1", "film2", "film3"):
            lin(f"{p}.{f}.block.1", 2 * D, D)
        spec.append((p + ".rotary.freqs", (D // 2,), "freqs"))
        lin(p + ".linear3", D, D)
        ln(p + ".norm4")
        # dead w.r.t. the output (model.py:300-304,344-355) but part of the state_dict contract
        for j, (o, k) in enumerate(((128, D), (128, 128), (D, 128))):
            lin(f"{p}.traj_Modulation.{j}._layer", o, k)
            lin(f"{p}.traj_Modulation.{j}._hyper_bias", o, 512, bias=False)
            lin(f"{p}.traj_Modulation.{j}._hyper_gate", o, 512)
    lin("final_layer", nf, D)
    lin("relative_projection_layer.0", 2 * D, D * dn)
    lin("relative_projection_layer.2", 2 * D, 2 * D)
    lin("relative_projection_layer.4", D * dn, 2 * D)
    spec.append(("embeddings_table.weight", (10, 64 * H), "randn"))
    lin("traj_embedding.0", 64, 2)
    lin("traj_embedding.2", D, 64)
    return spec


def make_state_dict(cfg, seed=0):
    """torch.nn.Linear-like magnitudes (U(+-1/sqrt(fan_in))), perturbed LayerNorm affines so that
    affine mistakes are visible, N(0,1) for the null embeddings."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    D = cfg["latent_dim"]
    for key, shape, kind in state_dict_spec(cfg):
        if kind == "w":
            bound = 1.0 / (shape[-1] ** 0.5)
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * bound
        elif kind == "b":
            sd[key] = (torch.rand(shape, generator=g) * 2 - 1) * 0.05
        elif kind == "ln_w":
            sd[key] = 1.0 + 0.1 * torch.randn(shape, generator=g)
        elif kind == "ln_b":
            sd[key] = 0.1 * torch.randn(shape, generator=g)
        elif kind == "randn":
            sd[key] = torch.randn(shape, generator=g)
        elif kind == "freqs":
            sd[key] = 1.0 / (10000 ** (torch.arange(0, D, 2)[: D // 2].float() / D))
    return sd


def weight_checksum(sd):
    """Order-independent fp64 checksum of a state_dict (guards generator reproducibility)."""
    tot = 0.0
    for k in sorted(sd):
        v = sd[k].double()
        tot += float((v * torch.linspace(0.5, 1.5, v.numel(), dtype=torch.float64).reshape(v.shape)).sum())
    return tot


def make_motion(B, dn, S=150, seed=1234):
    """(B, dn, S, 151) dataset-style motion: MinMax-scaled channels in [-1,1]
    (dataset/preprocess.py:31), contact channels 0-3 in {-1,+1} (dataset/group_dataset.py:204-213)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(B, dn, S, 151, generator=g) * 2 - 1
    x[..., :4] = (torch.rand(B, dn, S, 4, generator=g) > 0.5).float() * 2 - 1
    return x


def make_prediction(B, dn, S=150, seed=51):
    """(B, S*dn, 151) synthetic network output, frame-major: motion-like values plus noise; positive
    contact logits pushed above the 0.95 threshold of model/diffusion.py:722 so the foot term is live."""
    pred = make_motion(B, dn, S, seed=seed).permute(0, 2, 1, 3).reshape(B, S * dn, 151).contiguous()
    pred = pred + 0.05 * torch.randn(pred.shape, generator=torch.Generator().manual_seed(seed + 1))
    pred[..., :4] = torch.where(pred[..., :4] > 0, pred[..., :4].abs().clamp_min(0.96), pred[..., :4])
    return pred


def make_music(B, Fm, S=150, seed=1235):
    """(B, 2S+1, Fm) N(0,1) music features (dataset returns 301 frames at 60 fps, TCDiff.py:223)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(B, 2 * S + 1, Fm, generator=g)


def make_traj(x_start):
    """Dataset-style trajectory conditioning, TCDiff.py:283-302: channels [4,5] of a motion batch
    (B,dn,S,151) -> (B, S*dn, 3) frame-major, z zero-padded."""
    B, dn, S, _ = x_start.shape
    t = torch.zeros(B, dn, S, 3)
    t[..., :2] = x_start[..., 4:6]
    return t.permute(0, 2, 1, 3).reshape(B, S * dn, 3).contiguous()


def make_noise_bank(shape, n, seed=4321):
    """[x_T, step noises...] — n+1 tensors N(0,1) from one CPU generator."""
    g = torch.Generator().manual_seed(seed)
    return [torch.randn(shape, generator=g) for _ in range(n + 1)]


def make_inpaint_constraint(shape, seed=71):
    """Deterministic in-painting constraint of the inpaint_loop fixture: contact + root channels of every other dancer
    token block fully imposed, a soft 0.5 mask on one rotation block, everything else free."""
    g = torch.Generator().manual_seed(seed)
    value = (torch.rand(shape, generator=g) * 2 - 1) * 0.8
    mask = torch.zeros(shape)
    mask[:, ::2, :7] = 1.0
    mask[:, 40:200, 7:13] = 0.5
    return {"mask": mask, "value": value}
