"""API-compatible replacement of the third-party `diff_gaussian_rasterization` (the "-w-pose" fork the
reference pins in requirements.txt:17) as used at src/model/decoder/cuda_splatting.py:5-8,101-129:

    settings   = GaussianRasterizationSettings(image_height, image_width, tanfovx, tanfovy, bg, scale_modifier,
                                               viewmatrix, projmatrix, projmatrix_raw, sh_degree, campos,
                                               prefiltered, debug)
    rasterizer = GaussianRasterizer(settings)
    color, radii, depth, opacity, n_touched = rasterizer(means3D, means2D, opacities, shs=..., colors_precomp=...,
                                                         scales=..., rotations=..., cov3D_precomp=...,
                                                         theta=..., rho=...)

One view per call (that is the upstream contract); the batched fast path is styl3r_b200.decoder.render_cuda.
"""
from __future__ import annotations

from typing import NamedTuple, Optional

import torch
from torch import nn

from ... import rasterizer as _rz


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


def _cov_from_scale_rot(scales, rotations, scale_modifier):
    """Upstream computeCov3D: Sigma = R S S^T R^T with quaternion (w, x, y, z), packed upper triangle."""
    q = rotations / rotations.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    r, x, y, z = q.unbind(-1)
    R = torch.stack([1 - 2 * (y * y + z * z), 2 * (x * y - r * z), 2 * (x * z + r * y),
                     2 * (x * y + r * z), 1 - 2 * (x * x + z * z), 2 * (y * z - r * x),
                     2 * (x * z - r * y), 2 * (y * z + r * x), 1 - 2 * (x * x + y * y)], -1).reshape(-1, 3, 3)
    M = R * (scales * scale_modifier)[:, None, :]
    S = M @ M.transpose(1, 2)
    i, j = torch.triu_indices(3, 3)
    return S[:, i, j]


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions: torch.Tensor) -> torch.Tensor:
        """Upstream mark_visible: camera-space z > 0.2 (the near cull of in_frustum)."""
        vm = self.raster_settings.viewmatrix
        z = positions @ vm[:3, 2] + vm[3, 2]
        return z > 0.2

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        s = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception("Please provide excatly one of either SHs or precomputed colors!")
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (
                (scales is not None or rotations is not None) and cov3D_precomp is not None):
            raise Exception("Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!")
        if cov3D_precomp is None:
            cov3D_precomp = _cov_from_scale_rot(scales, rotations, s.scale_modifier)
        dev = means3D.device
        f = lambda v: torch.as_tensor(v, dtype=torch.float32, device=dev)
        tanfov = torch.stack([f(s.tanfovx).reshape(()), f(s.tanfovy).reshape(())])[None]
        color, depth, opacity, radii, n_touched = _rz.rasterize(
            means3D[None], cov3D_precomp[None], opacities.reshape(1, -1),
            shs=None if shs is None else shs[None], colors_precomp=None if colors_precomp is None else colors_precomp[None],
            rho=None if rho is None else rho.reshape(1, 3), theta=None if theta is None else theta.reshape(1, 3),
            viewmatrix=s.viewmatrix[None], projmatrix=s.projmatrix[None], projmatrix_raw=s.projmatrix_raw[None],
            campos=s.campos.reshape(1, 3), tanfov=tanfov, background=s.bg.reshape(1, 3), W=int(s.image_width),
            H=int(s.image_height), sh_degree=int(s.sh_degree), want_n_touched=True, means2D=means2D)
        return color[0], radii[0], depth, opacity, n_touched[0]
