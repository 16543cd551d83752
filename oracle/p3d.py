"""TEST INFRASTRUCTURE — restatement of the ``pytorch3d.transforms`` functions on the hot path.

pytorch3d==0.7.1 (/root/reference/requirements.txt:70) is a third-party, un-vendored
dependency that is not installed offline; the reference has no test that pins results
at this boundary, so parity here is **unpinned**: these functions restate the published
0.7.x ``transforms/rotation_conversions.py`` algorithms.

Reference call sites:
  dataset/quaternion.py:30  rotation_6d_to_matrix     dataset/quaternion.py:31  matrix_to_axis_angle
  vis.py:369  axis_angle_to_quaternion   vis.py:390-391  quaternion_apply   vis.py:397-398  quaternion_multiply
"""
import torch
import torch.nn.functional as F


def rotation_6d_to_matrix(d6):
    # Gram-Schmidt on the two 3-vectors; F.normalize uses eps=1e-12; rows are b1,b2,b3.
    a1, a2 = d6[..., 0:3], d6[..., 3:6]
    b1 = F.normalize(a1, dim=-1)
    proj = (b1 * a2).sum(dim=-1, keepdim=True)
    b2 = F.normalize(a2 - proj * b1, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack([b1, b2, b3], dim=-2)


def _sqrt_pos(x):
    out = torch.zeros_like(x)
    pos = x > 0
    out[pos] = torch.sqrt(x[pos])
    return out


def matrix_to_quaternion(m):
    # 0.7.1: four candidate quaternions, pick the one with the largest |component|,
    # divide by 2*max(q_abs, 0.1); no sign standardisation in this release.
    lead = m.shape[:-2]
    e = m.reshape(lead + (9,))
    m00, m01, m02, m10, m11, m12, m20, m21, m22 = e.unbind(-1)
    q_abs = _sqrt_pos(torch.stack([1.0 + m00 + m11 + m22,
                                   1.0 + m00 - m11 - m22,
                                   1.0 - m00 + m11 - m22,
                                   1.0 - m00 - m11 + m22], dim=-1))
    cand = torch.stack([
        torch.stack([q_abs[..., 0] ** 2, m21 - m12, m02 - m20, m10 - m01], dim=-1),
        torch.stack([m21 - m12, q_abs[..., 1] ** 2, m10 + m01, m02 + m20], dim=-1),
        torch.stack([m02 - m20, m10 + m01, q_abs[..., 2] ** 2, m12 + m21], dim=-1),
        torch.stack([m10 - m01, m20 + m02, m21 + m12, q_abs[..., 3] ** 2], dim=-1),
    ], dim=-2)
    floor = torch.tensor(0.1, dtype=q_abs.dtype, device=q_abs.device)
    cand = cand / (2.0 * q_abs[..., None].max(floor))
    pick = F.one_hot(q_abs.argmax(dim=-1), num_classes=4) > 0.5
    return cand[pick, :].reshape(lead + (4,))


def quaternion_to_axis_angle(q):
    n = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(n, q[..., :1])
    ang = 2 * half
    small = ang.abs() < 1e-6
    k = torch.empty_like(ang)
    k[~small] = torch.sin(half[~small]) / ang[~small]
    k[small] = 0.5 - (ang[small] * ang[small]) / 48
    return q[..., 1:] / k


def axis_angle_to_quaternion(aa):
    ang = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = ang * 0.5
    small = ang.abs() < 1e-6
    k = torch.empty_like(ang)
    k[~small] = torch.sin(half[~small]) / ang[~small]
    k[small] = 0.5 - (ang[small] * ang[small]) / 48
    return torch.cat([torch.cos(half), aa * k], dim=-1)


def matrix_to_axis_angle(m):
    return quaternion_to_axis_angle(matrix_to_quaternion(m))


def quaternion_raw_multiply(a, b):
    aw, ax, ay, az = a.unbind(-1)
    bw, bx, by, bz = b.unbind(-1)
    return torch.stack([aw * bw - ax * bx - ay * by - az * bz,
                        aw * bx + ax * bw + ay * bz - az * by,
                        aw * by - ax * bz + ay * bw + az * bx,
                        aw * bz + ax * by - ay * bx + az * bw], dim=-1)


def standardize_quaternion(q):
    return torch.where(q[..., 0:1] < 0, -q, q)


def quaternion_multiply(a, b):
    return standardize_quaternion(quaternion_raw_multiply(a, b))


def quaternion_invert(q):
    return q * q.new_tensor([1, -1, -1, -1])


def quaternion_apply(q, p):
    pq = torch.cat([p.new_zeros(p.shape[:-1] + (1,)), p], dim=-1)
    return quaternion_raw_multiply(quaternion_raw_multiply(q, pq), quaternion_invert(q))[..., 1:]


# --- import-time-only names (dataset/quaternion.py:2-4, model/diffusion.py:13-14); not on the hot path
def quaternion_to_matrix(q):
    r, i, j, k = q.unbind(-1)
    s = 2.0 / (q * q).sum(-1)
    o = torch.stack([1 - s * (j * j + k * k), s * (i * j - k * r), s * (i * k + j * r),
                     s * (i * j + k * r), 1 - s * (i * i + k * k), s * (j * k - i * r),
                     s * (i * k - j * r), s * (j * k + i * r), 1 - s * (i * i + j * j)], dim=-1)
    return o.reshape(q.shape[:-1] + (3, 3))


def axis_angle_to_matrix(aa):
    return quaternion_to_matrix(axis_angle_to_quaternion(aa))


def matrix_to_rotation_6d(m):
    return m[..., :2, :].clone().reshape(m.shape[:-2] + (6,))
