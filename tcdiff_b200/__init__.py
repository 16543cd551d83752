"""tcdiff_b200 — B200-native (sm_100a) implementation of the TCDiff denoising hot path.

Drop-in names (same signatures as the reference):
    DanceDecoder, GaussianDiffusion, EMA, SMPLSkeleton, RotaryEmbedding, ax_from_6v, Adan,
    TrajDecoder, kalman_smooth_batch (config-5 front end)
"""
from ._lib import LIB_PATH, TcdError, lib  # noqa: F401
from .model import DanceDecoder  # noqa: F401
from .diffusion import GaussianDiffusion, EMA  # noqa: F401
from .skeleton import SMPLSkeleton, ax_from_6v  # noqa: F401
from .rotary import RotaryEmbedding  # noqa: F401
from .adan import Adan  # noqa: F401
from .traj import TrajDecoder, kalman_smooth_batch, generate_trajectory  # noqa: F401
