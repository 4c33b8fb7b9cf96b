"""`render_cuda` / `render_cuda_orthographic` with the reference's signature and semantics
(src/model/decoder/cuda_splatting.py:46-133, :136-227), re-designed for B200:

  * all `batch` views are rendered by ONE kernel chain (styl3r_b200.rasterizer) instead of a Python loop with
    two `.item()` host syncs, a `GaussianRasterizationSettings` and ~12 small torch ops per view;
  * the scale-invariant rescale (`mean*s`, `cov*s^2`, reference lines 65-72) and the 3x3 -> 6 covariance gather
    (lines 118,126) happen inside the preprocess kernel, so no scaled/gathered copies are materialised;
  * `view_set` (extension) lets several views share one Gaussian set instead of `repeat`-ing it
    (decoder_splatting_cuda.py:57-60 materialises v copies).

Camera conventions are the reference's: OpenCV camera-to-world extrinsics, normalised intrinsics, a symmetric
frustum from `get_fov` (principal point assumed centred), matrices handed over transposed.
"""
from __future__ import annotations

from math import isqrt
from typing import Literal, Optional

import torch
from torch import Tensor

import ctypes as _C

from .. import _lib
from .. import rasterizer as _rz

DepthRenderingMode = Literal["depth", "disparity", "relative_disparity", "log"]


def get_fov(intrinsics: Tensor) -> Tensor:
    """Field of view (x, y) in radians from normalised intrinsics [B,3,3] — angle between the rays through the
    mid-points of opposite image edges (src/geometry/projection.py:247-261)."""
    k_inv = intrinsics.inverse()
    edge = intrinsics.new_tensor([[0.0, 0.5, 1.0], [1.0, 0.5, 1.0], [0.5, 0.0, 1.0], [0.5, 1.0, 1.0]])
    rays = torch.einsum("bij,ej->bei", k_inv, edge)
    rays = rays / rays.norm(dim=-1, keepdim=True)
    fov_x = (rays[:, 0] * rays[:, 1]).sum(-1).acos()
    fov_y = (rays[:, 2] * rays[:, 3]).sum(-1).acos()
    return torch.stack((fov_x, fov_y), dim=-1)


def get_projection_matrix(near: Tensor, far: Tensor, fov_x: Tensor, fov_y: Tensor) -> Tensor:
    """[B,4,4] perspective matrix: x,y -> (-1,1), z -> (0,1), w = z (cuda_splatting.py:16-43)."""
    tan_x, tan_y = (0.5 * fov_x).tan(), (0.5 * fov_y).tan()
    right, top = tan_x * near, tan_y * near
    left, bottom = -right, -top
    m = near.new_zeros((near.shape[0], 4, 4), dtype=torch.float32)
    m[:, 0, 0] = 2 * near / (right - left)
    m[:, 1, 1] = 2 * near / (top - bottom)
    m[:, 0, 2] = (right + left) / (right - left)
    m[:, 1, 2] = (top + bottom) / (top - bottom)
    m[:, 3, 2] = 1
    m[:, 2, 2] = far / (far - near)
    m[:, 2, 3] = -(far * near) / (far - near)
    return m


def camera_setup(extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, scale_invariant: bool,
                 input_is_w2c: bool = False):
    """All per-view camera tensors of `render_cuda` in ONE kernel launch (s3r_camera_setup): returns
    (viewmatrix_T [B,4,4], full_projection_T [B,4,4], projection_T [B,4,4], campos [B,3], tan_fov [B,2], scale [B])
    in the transposed layout the rasterizer expects (reference lines 65-72, 81-88).  Camera tensors are treated
    as constants (no autograd), as everywhere in Styl3R — pose optimisation goes through the cam deltas."""
    if extrinsics.device.type != "cuda":
        raise _lib.S3RError("styl3r_b200 render_cuda needs CUDA tensors (no CPU fallback)")
    b = extrinsics.shape[0]
    dev = extrinsics.device
    f = lambda t: t.detach().to(torch.float32).contiguous()
    e, k, n_, f_ = f(extrinsics), f(intrinsics), f(near), f(far)
    buf = torch.empty(b * 54, dtype=torch.float32, device=dev)
    view_t, full, proj_t = (buf[16 * b * i:16 * b * (i + 1)].view(b, 4, 4) for i in range(3))
    campos = buf[48 * b:51 * b].view(b, 3)
    tan_fov = buf[51 * b:53 * b].view(b, 2)
    scale = buf[53 * b:54 * b]
    p = lambda t: _C.c_void_p(t.data_ptr())
    _lib.check(_lib.lib().s3r_camera_setup(p(e), p(k), p(n_), p(f_), int(bool(scale_invariant)), int(bool(input_is_w2c)), b,
                                           p(view_t), p(full),
                                           p(proj_t), p(campos), p(tan_fov), p(scale),
                                           _C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)), "s3r_camera_setup")
    return view_t, full, proj_t, campos, tan_fov, scale


def _sh_layout(sh: Tensor) -> Tensor:
    """[S,G,3,d_sh] -> [S,G,d_sh,3] (a free view when d_sh == 1)."""
    if sh.shape[-1] == 1:
        return sh.reshape(sh.shape[0], sh.shape[1], 1, 3)
    return sh.permute(0, 1, 3, 2).contiguous()


def _render(extrinsics, tan_fov, projection, near_scale, image_shape, background_color, means, covariances, sh,
            opacities, use_sh, cam_rot_delta, cam_trans_delta, view_set, want_aux=False, check="sync"):
    h, w = image_shape
    proj_t = projection.transpose(1, 2)                 # handed over transposed (reference :86)
    view_t = extrinsics.inverse().transpose(1, 2)       # (:87)
    full = view_t @ proj_t                              # (:88)
    degree = min(isqrt(sh.shape[-1]) - 1, 3)  # coefficients beyond degree 3 are stored but not evaluated (as upstream)
    shs = _sh_layout(sh)
    kw = dict(shs=shs) if use_sh else dict(colors_precomp=shs[:, :, 0, :])
    color, depth, opacity, radii, n_touched = _rz.rasterize(
        means, covariances, opacities, rho=cam_trans_delta, theta=cam_rot_delta, viewmatrix=view_t, projmatrix=full,
        projmatrix_raw=proj_t, campos=extrinsics[:, :3, 3], tanfov=tan_fov, background=background_color, W=w, H=h,
        sh_degree=degree, scales=near_scale, view_set=view_set, want_n_touched=want_aux, check=check, **kw)
    return color, depth, opacity, radii, n_touched


def render_cuda(
    extrinsics: Tensor,                 # [B,4,4] camera-to-world
    intrinsics: Tensor,                 # [B,3,3] normalised
    near: Tensor,                       # [B]
    far: Tensor,                        # [B]
    image_shape: tuple[int, int],
    background_color: Tensor,           # [B,3]
    gaussian_means: Tensor,             # [B,G,3]   ([S,G,3] with view_set)
    gaussian_covariances: Tensor,       # [B,G,3,3]
    gaussian_sh_coefficients: Tensor,   # [B,G,3,d_sh]
    gaussian_opacities: Tensor,         # [B,G]
    scale_invariant: bool = True,
    use_sh: bool = True,
    cam_rot_delta: Optional[Tensor] = None,    # [B,3]
    cam_trans_delta: Optional[Tensor] = None,  # [B,3]
    *,
    view_set: Optional[Tensor] = None,  # [B] int: Gaussian set rendered by each view (extension)
    check: str = "sync",
) -> tuple[Tensor, Tensor]:
    """Returns (color [B,3,h,w], depth [B,h,w]); differentiable w.r.t. means, covariances, SH, opacities and the
    camera deltas, like the reference."""
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    h, w = image_shape
    view_t, full, proj_t, campos, tan_fov, scale = camera_setup(extrinsics, intrinsics, near, far, scale_invariant)
    sh = gaussian_sh_coefficients
    # like upstream, coefficients beyond degree 3 are stored but not evaluated (config default sh_degree 4)
    degree = min(isqrt(sh.shape[-1]) - 1, 3)
    shs = _sh_layout(sh)
    kw = dict(shs=shs) if use_sh else dict(colors_precomp=shs[:, :, 0, :])
    color, depth, _, _, _ = _rz.rasterize(
        gaussian_means, gaussian_covariances, gaussian_opacities, rho=cam_trans_delta, theta=cam_rot_delta,
        viewmatrix=view_t, projmatrix=full, projmatrix_raw=proj_t, campos=campos, tanfov=tan_fov,
        background=background_color, W=w, H=h, sh_degree=degree, scales=scale if scale_invariant else None,
        view_set=view_set, check=check, **kw)
    return color, depth


def render_cuda_orthographic(
    extrinsics: Tensor, width: Tensor, height: Tensor, near: Tensor, far: Tensor, image_shape: tuple[int, int],
    background_color: Tensor, gaussian_means: Tensor, gaussian_covariances: Tensor,
    gaussian_sh_coefficients: Tensor, gaussian_opacities: Tensor, fov_degrees: float = 0.1, use_sh: bool = True,
    dump: Optional[dict] = None, *, view_set: Optional[Tensor] = None,
) -> Tensor:
    """Pseudo-orthographic render: the camera is moved back and given a tiny field of view
    (cuda_splatting.py:136-227). Returns color [B,3,h,w]."""
    b = extrinsics.shape[0]
    assert use_sh or gaussian_sh_coefficients.shape[-1] == 1
    fov_x = torch.tensor(fov_degrees, device=extrinsics.device).deg2rad()
    tan_x = (0.5 * fov_x).tan()
    dist = (0.5 * width) / tan_x
    tan_y = 0.5 * height / dist
    fov_y = (2 * tan_y).atan()
    near, far = near + dist, far + dist
    back = torch.eye(4, dtype=torch.float32, device=extrinsics.device).repeat(b, 1, 1)
    back[:, 2, 3] = -dist
    extrinsics = extrinsics @ back
    if dump is not None:
        dump.update(extrinsics=extrinsics, fov_x=fov_x, fov_y=fov_y, near=near, far=far)
    projection = get_projection_matrix(near, far, fov_x.expand(b), fov_y)
    tan_fov = torch.stack((tan_x.expand(b), tan_y.expand(b)), dim=-1)
    color, *_ = _render(extrinsics, tan_fov, projection, None, image_shape, background_color, gaussian_means,
                        gaussian_covariances, gaussian_sh_coefficients, gaussian_opacities, use_sh, None, None,
                        view_set)
    return color
