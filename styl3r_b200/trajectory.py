"""Camera trajectories for video rendering - the host-side mirror of
src/visualization/camera_trajectory/interpolation.py:8-258 with the same names, argument meaning and results:

    extrinsics = interpolate_extrinsics(initial[4,4], final[4,4], t[T])      # [T,4,4] fp32, camera-to-world
    intrinsics = interpolate_intrinsics(initial[3,3], final[3,3], t[T])      # [T,3,3]

`interpolate_extrinsics` rotates the camera around the "focus point" (least-squares intersection of the two look
rays; the origins' midpoint when the rays are parallel) in a 5-parameter pivot representation (3 translations in the
pivot frame + in-plane angle + twist).  The reference round-trips through scipy on the CPU (`Rotation.as_euler /
from_euler`, interpolation.py:90-110) - here the YXZ Euler conversions are closed-form torch ops, so the whole
trajectory is built on the tensors' own device (float64 inside, like the reference) with no host round trip.
"""
from __future__ import annotations

import math

import torch
from torch import Tensor


def interpolate_intrinsics(initial: Tensor, final: Tensor, t: Tensor) -> Tensor:
    """interpolation.py:8-16: linear blend, [*batch, T, 3, 3]."""
    initial, final = initial[..., None, :, :], final[..., None, :, :]
    return initial + (final - initial) * t[:, None, None]


def intersect_rays(a_origins: Tensor, a_directions: Tensor, b_origins: Tensor, b_directions: Tensor) -> Tensor:
    """Least-squares intersection of two rays (interpolation.py:19-49): solve (sum_i n_i n_i^T - I) p = sum_i (...) o_i."""
    a_origins, a_directions, b_origins, b_directions = torch.broadcast_tensors(a_origins, a_directions, b_origins,
                                                                               b_directions)
    origins = torch.stack((a_origins, b_origins), dim=-2)
    directions = torch.stack((a_directions, b_directions), dim=-2)
    n = directions[..., :, None] * directions[..., None, :] - torch.eye(3, dtype=origins.dtype, device=origins.device)
    lhs = n.sum(dim=-3)
    rhs = (n @ origins[..., None])[..., 0].sum(dim=-2)
    return torch.linalg.solve(lhs, rhs)


def _normalize(a: Tensor) -> Tensor:
    return a / a.norm(dim=-1, keepdim=True)


def generate_coordinate_frame(y: Tensor, z: Tensor) -> Tensor:
    """Columns (y x z, y, z) (interpolation.py:56-62)."""
    y, z = torch.broadcast_tensors(y, z)
    return torch.stack([torch.linalg.cross(y, z), y, z], dim=-1)


def generate_rotation_coordinate_frame(a: Tensor, b: Tensor, eps: float = 1e-4) -> Tensor:
    """Frame whose Y axis is normal to the plane of unit vectors a, b (interpolation.py:65-87): a b parallel to a is
    replaced by (0,0,1), then by (0,1,0) if still parallel."""
    b = b.detach().clone()
    for fallback in ((0.0, 0.0, 1.0), (0.0, 1.0, 0.0)):
        parallel = ((a * b).sum(-1).abs() - 1).abs() < eps
        b = torch.where(parallel[..., None], torch.tensor(fallback, dtype=b.dtype, device=b.device), b)
    return generate_coordinate_frame(_normalize(torch.linalg.cross(a, b)), a)


def matrix_to_euler_yxz(rot: Tensor) -> Tensor:
    """Intrinsic 'YXZ' Euler angles (y, x, z) of R = Ry(y) Rx(x) Rz(z) - what scipy's `as_euler("YXZ")` returns away
    from gimbal lock (interpolation.py:90-99):  R[0,2] = sy cx, R[2,2] = cy cx, R[1,2] = -sx, R[1,0] = cx sz,
    R[1,1] = cx cz."""
    y = torch.atan2(rot[..., 0, 2], rot[..., 2, 2])
    x = -torch.asin(rot[..., 1, 2].clamp(-1.0, 1.0))
    z = torch.atan2(rot[..., 1, 0], rot[..., 1, 1])
    return torch.stack((y, x, z), dim=-1)


def euler_yxz_to_matrix(angles: Tensor) -> Tensor:
    """R = Ry(y) Rx(x) Rz(z) (scipy `from_euler("YXZ")`, interpolation.py:102-110)."""
    y, x, z = angles.unbind(-1)
    cy, sy, cx, sx, cz, sz = y.cos(), y.sin(), x.cos(), x.sin(), z.cos(), z.sin()
    rows = [cy * cz + sy * sx * sz, -cy * sz + sy * sx * cz, sy * cx,
            cx * sz, cx * cz, -sx,
            -sy * cz + cy * sx * sz, sy * sz + cy * sx * cz, cy * cx]
    return torch.stack(rows, dim=-1).reshape(*angles.shape[:-1], 3, 3)


def extrinsics_to_pivot_parameters(extrinsics: Tensor, pivot_coordinate_frame: Tensor, pivot_point: Tensor) -> Tensor:
    """interpolation.py:113-140: (3 distances from the pivot in the (look x axis, axis, look) frame, in-plane angle,
    twist)."""
    pivot_axis = pivot_coordinate_frame[..., :, 1]
    translation_frame = generate_coordinate_frame(pivot_axis, extrinsics[..., :3, 2])
    delta = pivot_point - extrinsics[..., :3, 3]
    translation = (translation_frame.transpose(-1, -2) @ delta[..., None])[..., 0]
    inverted = torch.linalg.inv(pivot_coordinate_frame) @ extrinsics[..., :3, :3]
    ang = matrix_to_euler_yxz(inverted)
    return torch.cat([translation, ang[..., 0:1], ang[..., 2:3]], dim=-1)


def pivot_parameters_to_extrinsics(parameters: Tensor, pivot_coordinate_frame: Tensor, pivot_point: Tensor) -> Tensor:
    """interpolation.py:143-168."""
    translation, y, z = parameters.split((3, 1, 1), dim=-1)
    euler = torch.cat((y, torch.zeros_like(y), z), dim=-1)
    # the reference evaluates the Euler -> matrix step in scipy (float64) and rounds the matrix to the parameter dtype
    rotation = pivot_coordinate_frame @ euler_yxz_to_matrix(euler.double()).to(parameters.dtype)
    pivot_axis = pivot_coordinate_frame[..., :, 1]
    translation_frame = generate_coordinate_frame(pivot_axis, rotation[..., :3, 2])
    delta = (translation_frame @ translation[..., None])[..., 0]
    origin = pivot_point - delta
    extrinsics = torch.eye(4, dtype=parameters.dtype, device=parameters.device).expand(*origin.shape[:-1], 4, 4).clone()
    extrinsics[..., :3, :3] = rotation
    extrinsics[..., :3, 3] = origin
    return extrinsics


def interpolate_circular(a: Tensor, b: Tensor, t: Tensor) -> Tensor:
    """Shortest-arc angle interpolation (interpolation.py:171-196)."""
    a, b, t = torch.broadcast_tensors(a, b, t)
    tau = 2 * math.pi
    a, b = a % tau, b % tau
    d = (b - a).abs()
    a_left, a_right = a - tau, a + tau
    d_left, d_right = (b - a_left).abs(), (b - a_right).abs()
    use_d = (d < d_left) & (d < d_right)
    use_d_left = (d_left < d_right) & (~use_d)
    start = torch.where(use_d, a, torch.where(use_d_left, a_left, a_right))
    return start + (b - start) * t


def interpolate_pivot_parameters(initial: Tensor, final: Tensor, t: Tensor) -> Tensor:
    """interpolation.py:199-213."""
    initial, final, t = initial[..., None, :], final[..., None, :], t[:, None]
    ti, ri = initial.split((3, 2), dim=-1)
    tf, rf = final.split((3, 2), dim=-1)
    return torch.cat((ti + (tf - ti) * t, interpolate_circular(ri, rf, t)), dim=-1)


@torch.no_grad()
def interpolate_extrinsics(initial: Tensor, final: Tensor, t: Tensor, eps: float = 1e-4) -> Tensor:
    """interpolation.py:216-258: [*batch, T, 4, 4] fp32 camera-to-world matrices."""
    initial, final, t = initial.double(), final.double(), t.double()
    initial_look, final_look = initial[..., :3, 2], final[..., :3, 2]
    parallel_mask = ((initial_look * final_look).sum(-1).abs() - 1).abs() < eps
    initial_origin, final_origin = initial[..., :3, 3], final[..., :3, 3]
    midpoint = 0.5 * (initial_origin + final_origin)
    # branch-free form of the reference's masked assignment: for parallel rays the normal matrix is singular, so it is
    # replaced by the identity (its solution is discarded by the `where`)
    a_dir = torch.where(parallel_mask[..., None], torch.tensor((1.0, 0.0, 0.0), dtype=initial.dtype, device=initial.device),
                        initial_look)
    b_dir = torch.where(parallel_mask[..., None], torch.tensor((0.0, 1.0, 0.0), dtype=initial.dtype, device=initial.device),
                        final_look)
    focus = intersect_rays(initial_origin, a_dir, final_origin, b_dir)
    pivot_point = torch.where(parallel_mask[..., None], midpoint, focus)
    pivot_frame = generate_rotation_coordinate_frame(initial_look, final_look, eps=eps)
    initial_params = extrinsics_to_pivot_parameters(initial, pivot_frame, pivot_point)
    final_params = extrinsics_to_pivot_parameters(final, pivot_frame, pivot_point)
    params = interpolate_pivot_parameters(initial_params, final_params, t)
    return pivot_parameters_to_extrinsics(params.float(), pivot_frame[..., None, :, :].float(),
                                          pivot_point[..., None, :].float())


def smooth_time(num_frames: int, device=None, smooth: bool = True) -> Tensor:
    """Frame times of `render_video_generic` (infer_model_re10k.py:193-195): linspace(0,1) with cosine ease in/out."""
    t = torch.linspace(0, 1, num_frames, dtype=torch.float32, device=device)
    return (torch.cos(torch.pi * (t + 1)) + 1) / 2 if smooth else t


@torch.no_grad()
def generate_wobble_transformation(radius: Tensor, t: Tensor, num_rotations: int = 1,
                                   scale_radius_with_t: bool = True) -> Tensor:
    """src/visualization/camera_trajectory/wobble.py:7-22: circular translation in the image plane, [*batch, T, 4, 4]."""
    tf = torch.eye(4, dtype=torch.float32, device=t.device).broadcast_to((*radius.shape, t.shape[0], 4, 4)).clone()
    radius = radius[..., None]
    if scale_radius_with_t:
        radius = radius * t
    tf[..., 0, 3] = torch.sin(2 * torch.pi * num_rotations * t) * radius
    tf[..., 1, 3] = -torch.cos(2 * torch.pi * num_rotations * t) * radius
    return tf


@torch.no_grad()
def generate_wobble(extrinsics: Tensor, radius: Tensor, t: Tensor) -> Tensor:
    """wobble.py:25-32: camera-to-world poses wobbling around `extrinsics`, [*batch, T, 4, 4]."""
    return extrinsics[..., None, :, :] @ generate_wobble_transformation(radius, t)
