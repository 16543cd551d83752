"""gpu: the config-5 front end (TrajDecoder + sliding-window loop + Kalman smoother) against the reference-generated
golden tests/golden/traj.pt."""
import pytest
import torch

from conftest import load_golden

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    if not torch.cuda.is_available():
        pytest.skip("needs a GPU")
    return torch.device("cuda:0")


def test_trajdecoder_forward_and_generation_vs_reference_golden(dev):
    import tcdiff_b200 as T
    g = load_golden("traj.pt")
    cfg = g["cfg"]
    w, st = cfg["window_size"], cfg["step"]
    m = T.TrajDecoder(nfeats=cfg["nfeats"], trans_layer=cfg["trans_layer"], window_size=w, cond_feature_dim=cfg["cond_feature_dim"])
    m.load_state_dict(g["state_dict"], strict=True)
    m = m.to(dev).eval()
    x, cond = g["x"].to(dev), g["cond"].to(dev)
    one = m(x[:, :, :w].contiguous(), cond[:, :(w + st) * 2]).cpu()
    err = float((one - g["forward"]).abs().max() / g["forward"].abs().max())
    assert err < 1e-4, err
    # the generation loop takes channels 4, 5 of a 151-dim motion tensor (TCDiff.py:531)
    motion = torch.zeros(x.shape[0], x.shape[1], x.shape[2], 151, device=dev)
    motion[..., 4:6] = x
    raw = T.generate_trajectory(m, motion, cond, w, st, smooth=False).cpu()
    assert raw.shape == (3, 2, g["trajectory"].shape[2], 3) and float(raw[..., 2].abs().max()) == 0.0
    err = float((raw[..., :2] - g["trajectory"]).abs().max() / g["trajectory"].abs().max())
    assert err < 1e-4, err
    sm = T.kalman_smooth_batch(g["trajectory"].to(dev)).cpu()
    assert float((sm - g["kalman"]).abs().max()) < 1e-6
    full = T.generate_trajectory(m, motion, cond, w, st).cpu()
    assert float((full[..., :2] - g["kalman"]).abs().max()) < 1e-3
    with pytest.raises(T.TcdError):
        T.kalman_smooth_batch(g["trajectory"])
    with pytest.raises(NotImplementedError):
        m.train()(x[:, :, :w].contiguous(), cond[:, :(w + st) * 2])


def test_trajdecoder_full_size_runs(dev):
    """Reference defaults (6 layers, window 100, step 25, 438-dim music), config-5 per-GPU batch 32, 5 dancers."""
    import tcdiff_b200 as T
    torch.manual_seed(0)
    m = T.TrajDecoder(nfeats=2, trans_layer=6, window_size=100).to(dev).eval()
    x = torch.randn(32, 5, 150, 151, device=dev) * 0.3
    cond = torch.randn(32, 301, 438, device=dev)
    out = T.generate_trajectory(m, x, cond, 100, 25)
    assert out.shape == (32, 5, 150, 3) and bool(torch.isfinite(out).all())
    x0 = out.permute(0, 2, 1, 3).reshape(32, 750, 3)
    assert x0.shape == (32, 750, 3)
