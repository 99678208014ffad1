"""Oracle: the glue between the sampler output and the 330-d SMPL-X pose features.

Follows diffusion_rvqvae_trainer.py:199-219 (joint masks), :457-531 (split, x5, decode, de-norm,
trans cumsum, 6d -> R -> axis-angle -> scatter(165) -> R -> 6d) and utils/rotation_conversions.py
(:39-64 quaternion_to_matrix, :96-118 matrix_to_quaternion, :432-508 axis-angle <-> quaternion,
:511-550 6d <-> matrix). h3d: h3d_diffusion_new_trainer.py:194-221,573-607 (623-d scatter).
"""
import torch
import torch.nn.functional as F

UPPER_J = [3, 6, 9, 12, 13, 14, 15, 16, 17, 18, 19, 20, 21]
HANDS_J = list(range(25, 55))
LOWER_J = [0, 1, 2, 4, 5, 7, 8, 10, 11]


def six_d_mask(joints):
    return [6 * j + i for j in joints for i in range(6)]


def rotation_6d_to_matrix(d6):
    a1, a2 = d6[..., :3], d6[..., 3:]
    b1 = F.normalize(a1, dim=-1)
    b2 = a2 - (b1 * a2).sum(-1, keepdim=True) * b1
    b2 = F.normalize(b2, dim=-1)
    b3 = torch.cross(b1, b2, dim=-1)
    return torch.stack((b1, b2, b3), dim=-2)


def _sqrt_pos(x):
    return torch.where(x > 0, torch.sqrt(torch.clamp(x, min=0)), torch.zeros_like(x))


def _copysign(a, b):
    return torch.where((a < 0) != (b < 0), -a, a)


def matrix_to_quaternion(m):
    m00, m11, m22 = m[..., 0, 0], m[..., 1, 1], m[..., 2, 2]
    o0 = 0.5 * _sqrt_pos(1 + m00 + m11 + m22)
    x = 0.5 * _sqrt_pos(1 + m00 - m11 - m22)
    y = 0.5 * _sqrt_pos(1 - m00 + m11 - m22)
    z = 0.5 * _sqrt_pos(1 - m00 - m11 + m22)
    o1 = _copysign(x, m[..., 2, 1] - m[..., 1, 2])
    o2 = _copysign(y, m[..., 0, 2] - m[..., 2, 0])
    o3 = _copysign(z, m[..., 1, 0] - m[..., 0, 1])
    return torch.stack((o0, o1, o2, o3), -1)


def _half_sinc(angles, half):
    small = angles.abs() < 1e-6
    safe = torch.where(small, torch.ones_like(angles), angles)
    return torch.where(small, 0.5 - angles * angles / 48, torch.sin(half) / safe)


def quaternion_to_axis_angle(q):
    norms = torch.norm(q[..., 1:], p=2, dim=-1, keepdim=True)
    half = torch.atan2(norms, q[..., :1])
    return q[..., 1:] / _half_sinc(2 * half, half)


def axis_angle_to_quaternion(aa):
    angles = torch.norm(aa, p=2, dim=-1, keepdim=True)
    half = 0.5 * angles
    return torch.cat([torch.cos(half), aa * _half_sinc(angles, half)], dim=-1)


def quaternion_to_matrix(q):
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack((1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
                     two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
                     two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)), -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def matrix_to_axis_angle(m):
    return quaternion_to_axis_angle(matrix_to_quaternion(m))


def axis_angle_to_matrix(aa):
    return quaternion_to_matrix(axis_angle_to_quaternion(aa))


def assemble_330(rec_upper, rec_hands, rec_lower, ms, jaw_aa=None):
    """rec_* = decoder outputs [B,n,78|180|57]; ms = dict(mean[330], std[330], trans_mean[3], trans_std[3]).
    Returns rec_pose [B,n,330], rec_trans [B,n,3] (diffusion_rvqvae_trainer.py:484-531)."""
    B, n, _ = rec_upper.shape
    v = rec_lower[..., -3:] * ms["trans_std"] + ms["trans_mean"]
    trans = torch.cumsum(v, dim=-2)
    trans[..., 1] = v[..., 1]
    lower = rec_lower[..., :-3]
    um, hm, lm = six_d_mask(UPPER_J), six_d_mask(HANDS_J), six_d_mask(LOWER_J)
    upper = rec_upper * ms["std"][um] + ms["mean"][um]
    hands = rec_hands * ms["std"][hm] + ms["mean"][hm]
    lower = lower * ms["std"][lm] + ms["mean"][lm]
    aa = torch.zeros(B * n, 55, 3, dtype=rec_upper.dtype)
    for part, joints in ((upper, UPPER_J), (hands, HANDS_J), (lower, LOWER_J)):
        r = matrix_to_axis_angle(rotation_6d_to_matrix(part.reshape(B, n, len(joints), 6)))
        aa[:, joints] = r.reshape(B * n, len(joints), 3)
    if jaw_aa is not None:
        aa[:, 22] = jaw_aa.reshape(B * n, 3)
    m = axis_angle_to_matrix(aa)
    return m[..., :2, :].reshape(B, n, 330), trans


def sample_to_parts(sample, latent_scale=5.0):
    """sample [B,1536,1,T] -> three [B,T,512] latents x latent_scale (trainer:457-474; batched form h3d:735)."""
    lat = sample[:, :, 0, :].permute(0, 2, 1) * latent_scale
    return lat[..., :512].contiguous(), lat[..., 512:1024].contiguous(), lat[..., 1024:].contiguous()


# ---- h3d 623-d scatter (h3d_diffusion_new_trainer.py:194-221,604-607) --------------------------------

def h3d_masks():
    def body(joints, root=False):
        m = list(range(0, 4)) + list(range(619, 623)) if root else []
        for i in joints:
            if i > 0:
                m += [4 + (i - 1) * 3 + c for c in range(3)]
                m += [4 + 51 * 3 + (i - 1) * 6 + c for c in range(6)]
            m += [4 + 51 * 9 + i * 3 + c for c in range(3)]
        return m
    return body(UPPER_J), body(list(range(22, 52))), body(LOWER_J, root=True)


def assemble_623(rec_upper, rec_hands, rec_lower):
    B, n, _ = rec_upper.shape
    out = torch.zeros(B, n, 623, dtype=rec_upper.dtype)
    um, hm, lm = h3d_masks()
    out[..., um] = rec_upper
    out[..., hm] = rec_hands
    out[..., lm] = rec_lower
    return out
