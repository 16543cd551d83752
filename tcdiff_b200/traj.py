"""Trajectory front end of the end-to-end test mode: drop-in `TrajDecoder`
(TrajDecoder/model/traj_model.py:125-200, same constructor and state_dict keys), the sliding-window generation loop
(TCDiff.py:526-556) and `kalman_smooth_batch` (TrajDecoder/utils/utils_model.py:10-74) on the GPU.

The model is small (LSTM hidden 64, transformer width 128, 4 heads of 32) and latency bound, so it runs on the fp32
kernels: `tcd_lstm_layer` (the reference's LSTM recurs over the BATCH axis, traj_model.py:139,174 — one thread block per
(dancer, frame) sequence, sequential over the batch), `tcd_gemm(TCD_F32)` with ReLU/LeakyReLU/GELU epilogues,
`tcd_layernorm_rotary`, `tcd_attention_f32_hd` (head dim 32) and `tcd_film_residual_norm` for the residual adds.
The Kalman filter of the reference is a Python triple loop through filterpy on the host; its gain sequence does not
depend on the data, so it is computed once on the host in float64 (filterpy 1.4.5's predict/update algebra) and one
kernel runs every track's recurrence.  Inference only (eval mode); no CPU path.
"""
import math

import numpy as np
import torch
import torch.nn as nn

from . import _lib, ops
from ._lib import ACT_GELU, ACT_LEAKY_RELU, ACT_NONE, F32, check

H = 64          # LSTM hidden size == latent_dim default


def _stream():
    return torch.cuda.current_stream().cuda_stream


class PositionalEncoding(nn.Module):
    """TrajDecoder/model/utils.py:11-32 (the buffer is what matters: state_dict key `pe`)."""

    def __init__(self, d_model, dropout=0.1, max_len=500, batch_first=False):
        super().__init__()
        self.batch_first = batch_first
        self.dropout = nn.Dropout(p=dropout)
        pe = torch.zeros(max_len, d_model)
        position = torch.arange(0, max_len).unsqueeze(1)
        div_term = torch.exp(torch.arange(0, d_model, 2) * (-np.log(10000.0) / d_model))
        pe[:, 0::2] = torch.sin(position * div_term)
        pe[:, 1::2] = torch.cos(position * div_term)
        self.register_buffer("pe", pe.unsqueeze(0).transpose(0, 1))


class CausalCrossConditionalSelfAttention(nn.Module):
    def __init__(self, embed_dim=512, block_size=120, n_head=8, drop_out_rate=0.1):
        super().__init__()
        assert embed_dim % n_head == 0
        self.key = nn.Linear(embed_dim, embed_dim)
        self.query = nn.Linear(embed_dim, embed_dim)
        self.value = nn.Linear(embed_dim, embed_dim)
        self.attn_drop = nn.Dropout(drop_out_rate)
        self.resid_drop = nn.Dropout(drop_out_rate)
        self.proj = nn.Linear(embed_dim, embed_dim)
        # registered like the reference (state_dict key); the reference never applies it (traj_model.py:36-40)
        self.register_buffer("mask", torch.tril(torch.ones(block_size, block_size)).view(1, 1, block_size, block_size))
        self.n_head = n_head


class Block(nn.Module):
    def __init__(self, embed_dim=512, block_size=120, n_head=8, drop_out_rate=0.1, fc_rate=4):
        super().__init__()
        self.ln1 = nn.LayerNorm(embed_dim)
        self.ln2 = nn.LayerNorm(embed_dim)
        self.attn = CausalCrossConditionalSelfAttention(embed_dim, block_size, n_head, drop_out_rate)
        self.mlp = nn.Sequential(nn.Linear(embed_dim, fc_rate * embed_dim), nn.GELU(), nn.Linear(fc_rate * embed_dim, embed_dim),
                                 nn.Dropout(drop_out_rate))


class music2traj_Transformer(nn.Module):
    def __init__(self, embed_dim=64, music_dim=64, block_size=60, num_layers=2, n_head=8, drop_out_rate=0.1, fc_rate=4):
        super().__init__()
        self.cond_emb = nn.Linear(music_dim, embed_dim)
        self.traj_emb = nn.Linear(3, embed_dim)                       # unused by the reference's forward as well
        self.drop = nn.Dropout(drop_out_rate)
        self.blocks = nn.Sequential(*[Block(embed_dim + music_dim, block_size, n_head, drop_out_rate, fc_rate)
                                      for _ in range(num_layers)])
        self.pos_embed = PositionalEncoding(embed_dim, drop_out_rate, batch_first=True)
        self.block_size = block_size
        self.apply(self._init_weights)

    def _init_weights(self, module):
        if isinstance(module, (nn.Linear, nn.Embedding)):
            module.weight.data.normal_(mean=0.0, std=0.02)
            if isinstance(module, nn.Linear) and module.bias is not None:
                module.bias.data.zero_()
        elif isinstance(module, nn.LayerNorm):
            module.bias.data.zero_()
            module.weight.data.fill_(1.0)


class TrajDecoder(nn.Module):
    def __init__(self, nfeats, trans_layer=4, window_size=60, latent_dim: int = 64, dropout: float = 0.1, n_head: int = 4,
                 cond_feature_dim: int = 438):
        super().__init__()
        if latent_dim != H:
            raise NotImplementedError("the sm_100a LSTM kernel is built for latent_dim=64 (the reference default)")
        self.latent_dim, self.n_head, self.nfeats = latent_dim, n_head, nfeats
        self.lstm = torch.nn.LSTM(input_size=nfeats, hidden_size=latent_dim, num_layers=3)
        self.music_projection = nn.Sequential(nn.Linear(cond_feature_dim * 2, cond_feature_dim), nn.LeakyReLU(),
                                              nn.Linear(cond_feature_dim, cond_feature_dim), nn.LeakyReLU(),
                                              nn.Linear(cond_feature_dim, latent_dim))
        self.trans_extractor = music2traj_Transformer(embed_dim=latent_dim, drop_out_rate=dropout, block_size=window_size,
                                                      n_head=n_head, music_dim=latent_dim, num_layers=trans_layer)
        self.Decoder = nn.Sequential(nn.Linear(latent_dim * 3, latent_dim * 2), nn.LeakyReLU(),
                                     nn.Linear(latent_dim * 2, latent_dim * 2), nn.LeakyReLU(),
                                     nn.Linear(latent_dim * 2, latent_dim), nn.LeakyReLU(), nn.Linear(latent_dim, nfeats))

    # ------------------------------------------------------------------------------------------------ kernels
    @staticmethod
    def _lin(x, lin, act=ACT_NONE, out=None):
        """out (M, N) fp32 = act(x (M, K) W^T + b) on the fp32 GEMM."""
        M = x.shape[0]
        w = lin.weight.detach()
        out = torch.empty(M, w.shape[0], device=x.device) if out is None else out
        ops.gemm(x, w, lin.bias.detach(), act, out, M=M, N=w.shape[0], K=w.shape[1])
        return out

    @torch.no_grad()
    def forward(self, x, music_feat):
        if self.training:
            raise NotImplementedError("tcdiff_b200.TrajDecoder is inference only; call .eval()")
        if not x.is_cuda:
            raise _lib.TcdError("tcdiff_b200.TrajDecoder runs on CUDA only (no CPU fallback)")
        lib = _lib.lib()
        dev = x.device
        b, dn, seq, c = x.shape
        N, E = dn * seq, 2 * H
        pe = self.trans_extractor.pos_embed.pe
        if N > pe.shape[0]:
            raise ValueError(f"dancers*window = {N} exceeds the positional table ({pe.shape[0]}), as in the reference")
        xin = x.reshape(b, N, c).float().contiguous()
        z = torch.empty(b, N, E, device=dev)                          # [music (64) | trajectory (64)] per token
        # --- 3-layer LSTM over the batch axis; the last layer writes into z[..., 64:] with the positional table added
        src, src_ts, src_ld, I = xin, N * c, c, c
        table = pe[:N, 0, :].contiguous()
        for l in range(3):
            last = l == 2
            dst = z if last else torch.empty(b, N, H, device=dev)
            w = [getattr(self.lstm, f"{n}_l{l}").detach() for n in ("weight_ih", "weight_hh", "bias_ih", "bias_hh")]
            check(lib.tcd_lstm_layer(src.data_ptr(), src_ts, src_ld, w[0].data_ptr(), w[1].data_ptr(), w[2].data_ptr(),
                                     w[3].data_ptr(), dst.data_ptr() + (4 * H if last else 0), N * (E if last else H),
                                     E if last else H, table.data_ptr() if last else 0, b, N, I, _stream()))
            src, src_ts, src_ld, I = dst, N * H, H, H
        # --- music features: pair frames, 3-layer LeakyReLU MLP (traj_model.py:176-186)
        cb, cl, Fm = music_feat.shape
        mf = music_feat[:, : cl - (cl % 2), :].reshape(cb * (cl // 2), 2 * Fm).float().contiguous()
        cl2 = cl // 2
        mp = self.music_projection
        m = self._lin(self._lin(self._lin(mf, mp[0], ACT_LEAKY_RELU), mp[2], ACT_LEAKY_RELU), mp[4])       # (b*cl2, 64)
        ce = self._lin(m, self.trans_extractor.cond_emb)                                                   # (b*cl2, 64)
        for d in range(dn):            # music_feat[:, :seq].repeat(1, dn, 1) -> z[:, d*seq:(d+1)*seq, :64]
            ops.scatter_rows(ce, H, z, E, N * E, d * seq, seq, H, b, src_batch_stride=cl2 * H)
        # --- transformer blocks (no mask is applied by the reference)
        M = b * N
        z2 = z.view(M, E)
        scale = 1.0 / math.sqrt(E // self.n_head)
        for blk in self.trans_extractor.blocks:
            n1 = torch.empty(M, E, device=dev)
            ops.layernorm_rotary(z2, blk.ln1.weight.detach(), blk.ln1.bias.detach(), blk.ln1.eps, n1, None, None, None, M, E, 1)
            q, k, v = self._lin(n1, blk.attn.query), self._lin(n1, blk.attn.key), self._lin(n1, blk.attn.value)
            a = torch.empty(M, E, device=dev)
            check(lib.tcd_attention_f32_hd(E // self.n_head, q.data_ptr(), E, N * E, k.data_ptr(), E, N * E, v.data_ptr(), E,
                                           N * E, a.data_ptr(), E, N * E, b, self.n_head, N, N, scale, _stream()))
            y = self._lin(a, blk.attn.proj)
            ops.film_residual_norm(F32, z2, z2, y, None, 0.0, None, 0, 0, None, 0.0, None, None, None, None, M, E, N)
            ops.layernorm_rotary(z2, blk.ln2.weight.detach(), blk.ln2.bias.detach(), blk.ln2.eps, n1, None, None, None, M, E, 1)
            y = self._lin(self._lin(n1, blk.mlp[0], ACT_GELU), blk.mlp[2])
            ops.film_residual_norm(F32, z2, z2, y, None, 0.0, None, 0, 0, None, 0.0, None, None, None, None, M, E, N)
        # --- decoder input [features (128) | music of the predicted span (64)] and the 4-layer LeakyReLU MLP
        feat = torch.empty(b, N, 3 * H, device=dev)
        ops.scatter_rows(z2, E, feat, 3 * H, N * 3 * H, 0, N, E, b, src_batch_stride=N * E)
        for d in range(dn):            # music_feat[:, -seq:].repeat(1, dn, 1)
            ops.scatter_rows(m, H, feat, 3 * H, N * 3 * H, d * seq, seq, H, b, src_off=(cl2 - seq) * H, dst_off=E,
                             src_batch_stride=cl2 * H)
        dc = self.Decoder
        o = self._lin(self._lin(self._lin(self._lin(feat.view(M, 3 * H), dc[0], ACT_LEAKY_RELU), dc[2], ACT_LEAKY_RELU), dc[4],
                                ACT_LEAKY_RELU), dc[6])
        return o.reshape(b, dn, seq, c)


def kalman_gains(T, dt=1.0, process_noise_std=1e-2, measurement_noise_std=1e-1):
    """Gain sequence K_t (T, 4, 2) float64 of filterpy 1.4.5's KalmanFilter for the constant-velocity model of
    utils_model.py:31-55 (predict: P = FPF' + Q; update: S = HPH' + R, K = PH'S^-1, P = (I-KH)P(I-KH)' + KRK').  It does
    not depend on the measurements, so the host computes it once per call."""
    Fm = np.array([[1, 0, dt, 0], [0, 1, 0, dt], [0, 0, 1, 0], [0, 0, 0, 1]], dtype=np.float64)
    Hm = np.array([[1, 0, 0, 0], [0, 1, 0, 0]], dtype=np.float64)
    P = np.eye(4) * 10.0
    R = np.eye(2) * measurement_noise_std ** 2
    Q = np.eye(4) * process_noise_std
    out = np.zeros((T, 4, 2))
    for t in range(T):
        P = Fm @ P @ Fm.T + Q
        PHT = P @ Hm.T
        K = PHT @ np.linalg.inv(Hm @ PHT + R)
        IKH = np.eye(4) - K @ Hm
        P = IKH @ P @ IKH.T + K @ R @ K.T
        out[t] = K
    return out


@torch.no_grad()
def kalman_smooth_batch(xy_batch, dt=1.0, process_noise_std=1e-2, measurement_noise_std=1e-1):
    """utils_model.py:10-74 for a CUDA tensor (b, dn, T, 2): same shape/dtype/device out (the reference round-trips
    through numpy on the host, TCDiff.py:549-550)."""
    if not torch.is_tensor(xy_batch) or not xy_batch.is_cuda:
        raise _lib.TcdError("tcdiff_b200.kalman_smooth_batch needs a CUDA tensor (no CPU fallback)")
    b, dn, T, two = xy_batch.shape
    assert two == 2
    x = xy_batch.float().contiguous()
    out = torch.empty_like(x)
    gains = torch.from_numpy(kalman_gains(T, dt, process_noise_std, measurement_noise_std)).to(x.device)
    check(_lib.lib().tcd_kalman_smooth(x.data_ptr(), out.data_ptr(), gains.data_ptr(), b * dn, T, float(dt), _stream()))
    return out.to(xy_batch.dtype)


@torch.no_grad()
def generate_trajectory(traj_model, x, cond, window_size, step, smooth=True):
    """The trajectory generation of TCDiff.test_loop (TCDiff.py:526-556): x (b, dn, S, 151) dataset motion supplies the
    first window of root xy (channels 4, 5); the model extends it autoregressively `step` frames at a time over the music;
    Kalman smoothing; zero z.  Returns x_traj_padding (b, dn, S', 3) — reshape with
    `.permute(0, 2, 1, 3).reshape(b, S' * dn, 3)` for `ddim_sample(x_0=...)` as the reference does (:566)."""
    x_traj_xy = x[:, :, :, [4, 5]]
    cond_traj = x_traj_xy[:, :, :window_size, :].contiguous()
    pre = [cond_traj]
    for start in range(0, cond.shape[1] + 1 - (window_size + step) * 2, step * 2):
        cond_traj = traj_model(cond_traj, cond[:, start:start + (window_size + step) * 2])
        pre.append(cond_traj[:, :, -step:])
    x_traj = torch.cat(pre, dim=2)
    if smooth:
        x_traj = kalman_smooth_batch(x_traj)
    pad = torch.zeros(x_traj.shape[0], x_traj.shape[1], x_traj.shape[2], 3, device=x_traj.device, dtype=x_traj.dtype)
    pad[..., :2] = x_traj
    return pad
