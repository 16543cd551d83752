"""not-gpu, build container only: pin the oracle and the drop-in signatures against the live, unmodified
reference imported from /root/reference (skipped where that tree does not exist, e.g. on the GPU box)."""
import inspect

import pytest
import torch
import torch.nn.functional as F

from oracle import ref_shim, synth, tcdiff_oracle as O

pytestmark = pytest.mark.skipif(not ref_shim.available(), reason="reference tree not present")


def _ref_model(cfg):
    ns = ref_shim.load()
    return ns.DanceDecoder(nfeats=151, seq_len=cfg["seq_len"], latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"],
                           num_layers=cfg["num_layers"], num_heads=cfg["num_heads"], dropout=0.1,
                           cond_feature_dim=cfg["cond_feature_dim"], activation=F.gelu,
                           required_dancer_num=cfg["dancers"]).eval()


@pytest.mark.parametrize("name", ["tiny", "c1", "c2"])
def test_state_dict_contract(name):
    """Key names and shapes of synth spec == reference == drop-in (446 tensors at the TCDiff.py hyper-parameters)."""
    import tcdiff_b200 as T
    cfg = synth.CONFIGS[name]
    ref = {k: tuple(v.shape) for k, v in _ref_model(cfg).state_dict().items()}
    spec = {k: tuple(s) for k, s, _ in synth.state_dict_spec(cfg)}
    mine = T.DanceDecoder(nfeats=151, seq_len=cfg["seq_len"], latent_dim=cfg["latent_dim"], ff_size=cfg["ff_size"],
                          num_layers=cfg["num_layers"], num_heads=cfg["num_heads"],
                          cond_feature_dim=cfg["cond_feature_dim"], required_dancer_num=cfg["dancers"])
    got = {k: tuple(v.shape) for k, v in mine.state_dict().items()}
    assert ref == spec == got
    if name != "tiny":
        assert len(ref) == 446


def test_signatures_match_reference():
    import tcdiff_b200 as T
    ns = ref_shim.load()

    def params(f):
        return [(p.name, p.default) for p in inspect.signature(f).parameters.values() if p.kind != p.KEYWORD_ONLY
                and p.kind != p.VAR_KEYWORD]
    for a, b in ((ns.DanceDecoder.__init__, T.DanceDecoder.__init__), (ns.DanceDecoder.forward, T.DanceDecoder.forward),
                 (ns.DanceDecoder.guided_forward, T.DanceDecoder.guided_forward),
                 (ns.GaussianDiffusion.__init__, T.GaussianDiffusion.__init__),
                 (ns.GaussianDiffusion.ddim_sample, T.GaussianDiffusion.ddim_sample),
                 (ns.GaussianDiffusion.p_sample_loop, T.GaussianDiffusion.p_sample_loop),
                 (ns.GaussianDiffusion.p_losses, T.GaussianDiffusion.p_losses),
                 (ns.GaussianDiffusion.q_sample, T.GaussianDiffusion.q_sample),
                 (ns.GaussianDiffusion.loss, T.GaussianDiffusion.loss),
                 (ns.SMPLSkeleton.forward, T.SMPLSkeleton.forward),
                 (ns.RotaryEmbedding.rotate_queries_or_keys, T.RotaryEmbedding.rotate_queries_or_keys)):
        assert params(a) == params(b), (a.__qualname__, params(a), params(b))
    d_ref = ns.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, None, schedule="cosine")
    d_mine = T.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, None, schedule="cosine")
    rb, mb = dict(d_ref.named_buffers()), dict(d_mine.named_buffers())
    assert list(rb) == list(mb)
    for k in rb:
        assert torch.equal(rb[k], mb[k]), k
    for attr in ("model", "master_model", "ema", "smpl", "n_timestep", "guidance_weight", "cond_drop_prob", "seq_len", "horizon"):
        assert hasattr(d_mine, attr)


def test_oracle_forward_equals_live_reference():
    cfg = synth.CONFIGS["tiny"]
    sd = synth.make_state_dict(cfg, 3)
    m = _ref_model(cfg)
    m.load_state_dict(sd, strict=True)
    g = torch.Generator().manual_seed(11)
    x = torch.randn(3, 300, 151, generator=g)
    cond = synth.make_music(3, cfg["cond_feature_dim"], seed=12)
    t = torch.tensor([0, 512, 999])
    keep = torch.tensor([True, False, True])
    with torch.no_grad(), ref_shim.NoiseBank([], keep_mask=keep):
        ref = m(x, cond, t, cond_drop_prob=0.25)
        mine = O.dance_decoder_forward(sd, x, cond, t, keep_mask=keep)
    assert float((ref - mine).abs().max()) < 2e-5


def test_ddim_schedule_host_arithmetic_matches_reference_formula():
    import tcdiff_b200 as T
    d = T.GaussianDiffusion(torch.nn.Linear(1, 1), 150, 151, None, schedule="cosine", predict_epsilon=False, loss_type="l2")
    sched = O.make_schedule("cosine", 1000)
    for (t, tn, sr, srm1, sa, c, sigma), (t2, tn2) in zip(d._ddim_schedule(50, 1.0), O.ddim_times()):
        assert (t, tn) == (t2, tn2)
        if tn >= 0:
            a, b, s = O.ddim_coeffs(sched, t, tn)
            assert (sa, c, sigma) == (float(a), float(b), float(s))
        assert sr == float(sched["sqrt_recip_alphas_cumprod"][t]) and srm1 == float(sched["sqrt_recipm1_alphas_cumprod"][t])


def test_oracle_adan_and_ema_equal_reference_classes():
    """oracle.adan_step / ema_update vs the reference's Adan (model/adan.py) and EMA (model/diffusion.py:61-76):
    bit-identical over 4 steps, including a parameter that never gets a gradient."""
    ns = ref_shim.load()
    import importlib
    RefAdan = importlib.import_module("model.adan").Adan
    g = torch.Generator().manual_seed(5)
    shapes = [(7, 5), (33,), (4, 3, 2), (6,)]
    ref_p = [torch.nn.Parameter(torch.randn(s, generator=g)) for s in shapes]
    my_p = [p.detach().clone() for p in ref_p]
    ref_ma = [torch.nn.Parameter(p.detach().clone()) for p in ref_p]
    my_ma = [p.detach().clone() for p in ref_p]
    opt = RefAdan(ref_p, lr=4e-4, weight_decay=0.02)
    st = O.adan_init(my_p)

    class Holder:
        def __init__(self, ps):
            self.ps = ps

        def parameters(self):
            return iter(self.ps)

    ema = importlib.import_module("model.diffusion").EMA(0.9999)
    for it in range(4):
        grads = [torch.randn(s, generator=g) * (10.0 ** (it - 2)) for s in shapes]
        grads[3] = None                                        # dead parameter
        for p, gr in zip(ref_p, grads):
            p.grad = None if gr is None else gr.clone()
        opt.step()
        ema.update_model_average(Holder(ref_ma), Holder(ref_p))
        O.adan_step(my_p, grads, st, lr=4e-4, weight_decay=0.02)
        O.ema_update(my_ma, my_p, 0.9999)
        for a, b in zip(ref_p, my_p):
            assert torch.equal(a.detach(), b), it
        for a, b in zip(ref_ma, my_ma):
            assert torch.equal(a.detach(), b), it
    assert not torch.equal(my_p[0], my_ma[0])


def test_oracle_dropout_sites_match_reference_train_mode(monkeypatch):
    """Train mode: every nn.Dropout / attention-dropout site of the reference (model/model.py:98,103,240-245,383,396,
    400-401; nn.MultiheadAttention's dropout) is where the oracle's mask hook sits, in the same order.  torch's dropout
    is replaced by a deterministic mask keyed by call order in BOTH implementations; outputs must then agree."""
    import torch.nn.functional as TF
    cfg = synth.CONFIGS["tiny"]
    sd = synth.make_state_dict(cfg, 3)
    m = _ref_model(cfg)
    m.load_state_dict(sd, strict=True)
    m.train()
    assert any(isinstance(x, torch.nn.Dropout) and x.p == 0.1 for x in m.modules())
    p = 0.1
    calls = [0]

    def nth_mask(shape, dtype):
        n = calls[0]
        calls[0] += 1
        g = torch.Generator().manual_seed(1000 + n)
        return (torch.rand(tuple(shape), generator=g) >= p).to(dtype) / (1 - p)

    def fake_dropout(input, p=0.5, training=True, inplace=False):
        if not training or p == 0.0:
            return input
        return input * nth_mask(input.shape, input.dtype)

    def fake_sdpa(q, k, v, attn_mask=None, dropout_p=0.0, is_causal=False, scale=None, **kw):
        assert attn_mask is None and not is_causal
        s = torch.matmul(q, k.transpose(-2, -1)) * (scale if scale is not None else q.shape[-1] ** -0.5)
        att = torch.softmax(s, dim=-1)
        if dropout_p > 0:
            att = att * nth_mask(att.shape, att.dtype)
        return torch.matmul(att, v)

    monkeypatch.setattr(TF, "dropout", fake_dropout)
    monkeypatch.setattr(TF, "scaled_dot_product_attention", fake_sdpa)
    g = torch.Generator().manual_seed(21)
    x = torch.randn(2, 300, 151, generator=g)
    cond = synth.make_music(2, cfg["cond_feature_dim"], seed=22)
    t = torch.tensor([7, 640])
    keep = torch.tensor([True, False])
    with torch.no_grad(), ref_shim.NoiseBank([], keep_mask=keep):
        ref = m(x, cond, t, cond_drop_prob=0.25)
    n_ref = calls[0]
    assert n_ref == 2 * 4 + cfg["num_layers"] * 8                    # 4 sites per encoder layer, 8 per decoder layer
    calls[0] = 0
    seen = []

    def hook(kind, layer, k, tensor):
        seen.append((kind, layer, k))
        return tensor * nth_mask(tensor.shape, tensor.dtype)

    with torch.no_grad(), O.dropout_hook(hook):
        mine = O.dance_decoder_forward(sd, x, cond, t, keep_mask=keep)
    assert calls[0] == n_ref
    assert seen[:4] == [("enc", 0, 0), ("enc", 0, 1), ("enc", 0, 2), ("enc", 0, 3)]
    assert seen[8:16] == [("dec", 0, k) for k in range(8)]
    assert float((ref - mine).abs().max()) < 2e-5
    # and it is not a no-op
    with torch.no_grad():
        plain = O.dance_decoder_forward(sd, x, cond, t, keep_mask=keep)
    assert float((plain - mine).abs().max()) > 1e-3


def test_trajdecoder_contract_and_oracle_vs_live_reference():
    """Drop-in tcdiff_b200.TrajDecoder has the reference's constructor and state_dict keys/shapes (full-size defaults of
    TrajDecoder/options/option_traj.py: 6 layers, window 100), and the oracle equals the live reference module."""
    import tcdiff_b200 as T
    from oracle import make_golden as MG, traj_oracle as TO
    Ref = MG.load_ref_trajdecoder()
    assert list(inspect.signature(Ref.__init__).parameters) == list(inspect.signature(T.TrajDecoder.__init__).parameters)
    ref = Ref(nfeats=2, trans_layer=6, window_size=100)
    mine = T.TrajDecoder(nfeats=2, trans_layer=6, window_size=100)
    rs, ms = ref.state_dict(), mine.state_dict()
    assert list(rs) == list(ms)
    assert all(rs[k].shape == ms[k].shape and rs[k].dtype == ms[k].dtype for k in rs)
    mine.load_state_dict(rs, strict=True)
    torch.manual_seed(3)
    small = Ref(nfeats=2, trans_layer=2, window_size=10, cond_feature_dim=6).eval()
    x = torch.randn(4, 3, 10, 2)
    music = torch.randn(4, 25, 6)                      # odd length: last frame dropped (traj_model.py:178-179)
    with torch.no_grad():
        want = small(x, music)
        got = TO.traj_decoder_forward({k: v for k, v in small.state_dict().items()}, x, music)
    assert float((want - got).abs().max()) < 1e-5
