"""TEST INFRASTRUCTURE (oracle) -- never imported by the product path.

Restatement of the Isaac Lab ``omni.isaac.lab.utils.math`` functions the racing
hot path calls.  Isaac Lab is a third-party dependency that is NOT vendored in
/root/reference (pinned only as ``omni-isaac-lab>=0.27.15``, pyproject.toml:32);
the formulas below are its published BSD-3 algorithms as frozen in SURVEY.md
Appendix B.  Call sites in the reference: QD/mdp/dynamics/droneDynamics.py:116-134,
163-172; QD/mdp/commands.py:212-218,294-306; QD/mdp/observation.py:28-31;
QD/mdp/events.py:153-166; QD/mdp/termination.py:29-31.  Quaternions are (w,x,y,z).
"""
from __future__ import annotations

import math

import torch


def quat_mul(q1: torch.Tensor, q2: torch.Tensor) -> torch.Tensor:
    shape = q1.shape
    q1 = q1.reshape(-1, 4)
    q2 = q2.reshape(-1, 4)
    w1, x1, y1, z1 = q1[:, 0], q1[:, 1], q1[:, 2], q1[:, 3]
    w2, x2, y2, z2 = q2[:, 0], q2[:, 1], q2[:, 2], q2[:, 3]
    ww = (z1 + x1) * (x2 + y2)
    yy = (w1 - y1) * (w2 + z2)
    zz = (w1 + y1) * (w2 - z2)
    xx = ww + yy + zz
    qq = 0.5 * (xx + (z1 - x1) * (x2 - y2))
    w = qq - ww + (z1 - y1) * (y2 - z2)
    x = qq - xx + (x1 + w1) * (x2 + w2)
    y = qq - yy + (w1 - x1) * (y2 + z2)
    z = qq - zz + (z1 + y1) * (w2 - x2)
    return torch.stack([w, x, y, z], dim=-1).view(shape)


def quat_conjugate(q: torch.Tensor) -> torch.Tensor:
    shape = q.shape
    q = q.reshape(-1, 4)
    return torch.cat((q[:, 0:1], -q[:, 1:]), dim=-1).view(shape)


def quat_inv(q: torch.Tensor) -> torch.Tensor:
    return torch.nn.functional.normalize(quat_conjugate(q), p=2.0, dim=-1)


def quat_rotate(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    q_w = q[:, 0]
    q_vec = q[:, 1:]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(q.shape[0], 1, 3), v.view(q.shape[0], 3, 1)).squeeze(-1) * 2.0
    return a + b + c


def quat_rotate_inverse(q: torch.Tensor, v: torch.Tensor) -> torch.Tensor:
    q_w = q[:, 0]
    q_vec = q[:, 1:]
    a = v * (2.0 * q_w ** 2 - 1.0).unsqueeze(-1)
    b = torch.cross(q_vec, v, dim=-1) * q_w.unsqueeze(-1) * 2.0
    c = q_vec * torch.bmm(q_vec.view(q.shape[0], 1, 3), v.view(q.shape[0], 3, 1)).squeeze(-1) * 2.0
    return a - b + c


def quat_from_euler_xyz(roll: torch.Tensor, pitch: torch.Tensor, yaw: torch.Tensor) -> torch.Tensor:
    cy = torch.cos(yaw * 0.5)
    sy = torch.sin(yaw * 0.5)
    cr = torch.cos(roll * 0.5)
    sr = torch.sin(roll * 0.5)
    cp = torch.cos(pitch * 0.5)
    sp = torch.sin(pitch * 0.5)
    qw = cy * cr * cp + sy * sr * sp
    qx = cy * sr * cp - sy * cr * sp
    qy = cy * cr * sp + sy * sr * cp
    qz = sy * cr * cp - cy * sr * sp
    return torch.stack([qw, qx, qy, qz], dim=-1)


def matrix_from_quat(q: torch.Tensor) -> torch.Tensor:
    r, i, j, k = torch.unbind(q, -1)
    two_s = 2.0 / (q * q).sum(-1)
    o = torch.stack(
        (
            1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
            two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
            two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j),
        ),
        -1,
    )
    return o.reshape(q.shape[:-1] + (3, 3))


def _copysign(mag: float, other: torch.Tensor) -> torch.Tensor:
    mag_t = torch.full_like(other, mag)
    return torch.abs(mag_t) * torch.sign(other)


def euler_xyz_from_quat(q: torch.Tensor):
    q_w, q_x, q_y, q_z = q[:, 0], q[:, 1], q[:, 2], q[:, 3]
    sin_roll = 2.0 * (q_w * q_x + q_y * q_z)
    cos_roll = 1 - 2 * (q_x * q_x + q_y * q_y)
    roll = torch.atan2(sin_roll, cos_roll)
    sin_pitch = 2.0 * (q_w * q_y - q_z * q_x)
    pitch = torch.where(torch.abs(sin_pitch) >= 1, _copysign(math.pi / 2.0, sin_pitch), torch.asin(sin_pitch))
    sin_yaw = 2.0 * (q_w * q_z + q_x * q_y)
    cos_yaw = 1 - 2 * (q_y * q_y + q_z * q_z)
    yaw = torch.atan2(sin_yaw, cos_yaw)
    return roll % (2 * math.pi), pitch % (2 * math.pi), yaw % (2 * math.pi)


def wrap_to_pi(angles: torch.Tensor) -> torch.Tensor:
    wrapped = (angles + math.pi) % (2 * math.pi)
    return torch.where((wrapped == 0) & (angles > 0), torch.full_like(wrapped, math.pi), wrapped - math.pi)


def sample_uniform_from(u: torch.Tensor, lower, upper) -> torch.Tensor:
    """``sample_uniform`` with the ``torch.rand`` draw passed in: rand * (upper - lower) + lower."""
    return u * (upper - lower) + lower
