"""CPU restatement of the reference's pose-delta helpers.

TEST INFRASTRUCTURE ONLY — imported by tests/ (never by styl3r_b200/).

PINNED: tests/test_camera_pose_cpu.py checks every function below against golden vectors produced by the reference's
own `SE3_exp` / `update_pose` (tests/golden/make_camera_pose_golden.py -> camera_pose_golden.npz).

Follows /root/reference/src/misc/cam_utils.py: skew_sym_mat :54-64, SO3_exp :67-82, V :85-100, SE3_exp :103-115,
update_pose :118-137.  fp32 throughout, one rounding per torch op like the reference.
"""
from __future__ import annotations

import numpy as np

f32 = np.float32


def skew(x):
    """cam_utils.py:54-64."""
    m = np.zeros((3, 3), f32)
    m[0, 1], m[0, 2], m[1, 0], m[1, 2], m[2, 0], m[2, 1] = -x[2], x[1], x[2], -x[0], -x[1], x[0]
    return m


def so3_exp(theta):
    """cam_utils.py:67-82 (first-order branch below 1e-5 rad)."""
    theta = np.asarray(theta, f32)
    W = skew(theta)
    W2 = (W @ W).astype(f32)
    angle = f32(np.sqrt(f32((theta * theta).sum())))
    I = np.eye(3, dtype=f32)
    if angle < 1e-5:
        return (I + W + f32(0.5) * W2).astype(f32)
    return (I + f32(np.sin(angle) / angle) * W + f32((f32(1) - f32(np.cos(angle))) / f32(angle * angle)) * W2).astype(f32)


def v_mat(theta):
    """cam_utils.py:85-100."""
    theta = np.asarray(theta, f32)
    W = skew(theta)
    W2 = (W @ W).astype(f32)
    angle = f32(np.sqrt(f32((theta * theta).sum())))
    I = np.eye(3, dtype=f32)
    if angle < 1e-5:
        return (I + f32(0.5) * W + f32(1.0 / 6.0) * W2).astype(f32)
    a2 = f32(angle * angle)
    a3 = f32(a2 * angle)
    return (I + W * f32((f32(1) - f32(np.cos(angle))) / a2) + W2 * f32((angle - f32(np.sin(angle))) / a3)).astype(f32)


def se3_exp(tau):
    """cam_utils.py:103-115: tau = (rho, theta) -> 4x4."""
    tau = np.asarray(tau, f32)
    T = np.eye(4, dtype=f32)
    T[:3, :3] = so3_exp(tau[3:])
    T[:3, 3] = (v_mat(tau[3:]) @ tau[:3]).astype(f32)
    return T


def update_pose(cam_trans_delta, cam_rot_delta, extrinsics):
    """cam_utils.py:118-137: c2w' = inverse(SE3_exp([trans, rot]) @ inverse(c2w)), batched."""
    out = []
    for t, r, e in zip(cam_trans_delta, cam_rot_delta, extrinsics):
        w2c = np.linalg.inv(np.asarray(e, f32).astype(np.float64)).astype(f32)
        new = (se3_exp(np.concatenate([t, r])) @ w2c).astype(f32)
        out.append(np.linalg.inv(new.astype(np.float64)).astype(f32))
    return np.stack(out)
