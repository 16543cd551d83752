"""Drop-in SMPLSkeleton (reference: vis.py:330-406) and ax_from_6v (dataset/quaternion.py:28-32)."""
import torch

from . import ops


class SMPLSkeleton:
    def __init__(self, device=None):
        self.device = device

    def forward(self, rotations, root_positions):
        """rotations (N, L, 24, 3) axis-angle, root_positions (N, L, 3) -> (N, L, 24, 3) joint positions."""
        assert len(rotations.shape) == 4, "Rotations should be a 4D tensor."
        assert len(root_positions.shape) == 3, "Root positions should be a 3D tensor."
        assert rotations.shape[2] == 24 and rotations.shape[3] == 3
        aa = rotations.float().contiguous()
        root = root_positions.float().contiguous()
        pos = torch.empty_like(aa)
        ops.smpl_fk(aa, root, pos, aa.shape[0] * aa.shape[1])
        return pos

    __call__ = forward

    def motion_forward(self, motion):
        """Fused path used by the loss / post-processing: motion (..., 151) rows -> (..., 24, 3)."""
        m = motion.float().contiguous()
        pos = torch.empty(m.shape[:-1] + (24, 3), dtype=torch.float32, device=m.device)
        ops.motion_fk(m, pos, m.numel() // 151)
        return pos


def ax_from_6v(q):
    assert q.shape[-1] == 6
    d6 = q.float().contiguous()
    aa = torch.empty(d6.shape[:-1] + (3,), dtype=torch.float32, device=d6.device)
    ops.ax_from_6v(d6, aa, d6.numel() // 6)
    return aa
