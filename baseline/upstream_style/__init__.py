"""COMPARATOR ONLY: upstream-style stand-in for the reference's rasterizer call path (see upstream_style.cu).

`render_cuda_upstream_style` restates the structure of the reference's `render_cuda`
(src/model/decoder/cuda_splatting.py:46-133): torch camera set-up, scaled copies of the Gaussians, then a Python
loop over views with two `.item()` host syncs, a gathered [G,6] covariance tensor and one launch chain per view.
Nothing under styl3r_b200/ imports this package; it exists so that bench / scripts can quote a measured GPU
comparator for the ">= 10x the reference rasterizer" target (parity unpinned, labelled as a stand-in)."""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import torch

HERE = Path(__file__).resolve().parent
LIB = HERE / "libupstream_style.so"
_lib = None


def build(force: bool = False) -> Path:
    src = HERE / "upstream_style.cu"
    if LIB.exists() and not force and LIB.stat().st_mtime >= src.stat().st_mtime:
        return LIB
    cmd = ["nvcc", "-ccbin", "/usr/bin/g++", "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
           "-Xcompiler", "-fPIC", "-shared", "-o", str(LIB), str(src)]
    subprocess.run(cmd, check=True, capture_output=True)
    return LIB


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(str(LIB))
        _lib.ups_forward.restype = C.c_longlong
        _lib.ups_forward.argtypes = [C.c_int] * 3 + [C.c_void_p] * 6 + [C.c_float, C.c_float] + [C.c_void_p] * 6 + \
            [C.c_size_t, C.c_void_p]
    return _lib


def _get_fov(K):
    Ki = K.inverse()
    pts = K.new_tensor([[0, 0.5, 1], [1, 0.5, 1], [0.5, 0, 1], [0.5, 1, 1]])
    rays = torch.einsum("bij,ej->bei", Ki, pts)
    rays = rays / rays.norm(dim=-1, keepdim=True)
    return torch.stack(((rays[:, 0] * rays[:, 1]).sum(-1).acos(), (rays[:, 2] * rays[:, 3]).sum(-1).acos()), -1)


def render_cuda_upstream_style(extrinsics, intrinsics, near, far, image_shape, background_color, means, covariances,
                               sh, opacities, scratch=None):
    """Same inputs as the reference's render_cuda with batch = views (Gaussians already repeated per view)."""
    L = lib()
    b = extrinsics.shape[0]
    h, w = image_shape
    scale = 1 / near
    extrinsics = extrinsics.clone()
    extrinsics[..., :3, 3] = extrinsics[..., :3, 3] * scale[:, None]
    covariances = covariances * (scale[:, None, None, None] ** 2)
    means = means * scale[:, None, None]
    near, far = near * scale, far * scale
    shs = sh.permute(0, 1, 3, 2).contiguous()
    fov = _get_fov(intrinsics)
    tan = (0.5 * fov).tan()
    proj = torch.zeros(b, 4, 4, device=means.device)
    proj[:, 0, 0] = 1 / tan[:, 0]
    proj[:, 1, 1] = 1 / tan[:, 1]
    proj[:, 3, 2] = 1
    proj[:, 2, 2] = far / (far - near)
    proj[:, 2, 3] = -(far * near) / (far - near)
    proj_t = proj.transpose(1, 2)
    view_t = extrinsics.inverse().transpose(1, 2)
    full = view_t @ proj_t
    P = means.shape[1]
    if scratch is None:
        scratch = torch.empty(P * 64 + 16 * P * 24 + (1 << 26), dtype=torch.uint8, device=means.device)
    row, col = torch.triu_indices(3, 3)
    images, depths = [], []
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    for i in range(b):
        tx, ty = tan[i, 0].item(), tan[i, 1].item()   # the two host syncs per view of the reference
        cov6 = covariances[i][:, row, col].contiguous()
        color = torch.empty(3, h, w, device=means.device)
        depth = torch.empty(1, h, w, device=means.device)
        opacity = torch.empty(1, h, w, device=means.device)
        radii = torch.empty(P, dtype=torch.int32, device=means.device)
        vm, pm = view_t[i].contiguous(), full[i].contiguous()
        n = L.ups_forward(P, w, h, p(means[i].contiguous()), p(cov6), p(shs[i, :, 0].contiguous()),
                          p(opacities[i].contiguous()), p(vm), p(pm), tx, ty, p(background_color[i].contiguous()),
                          p(color), p(depth), p(opacity), p(radii), p(scratch), scratch.numel(), stream)
        if n < 0:
            raise RuntimeError("upstream-style stand-in failed (scratch too small?)")
        images.append(color)
        depths.append(depth.squeeze(0))
    return torch.stack(images), torch.stack(depths)
