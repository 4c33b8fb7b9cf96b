"""Pose update from camera deltas, on device, batched — replaces the per-view Python loop of
src/misc/cam_utils.py:118-137 (`update_pose`) and its helpers SE3_exp / SO3_exp / V (:67-115)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def se3_update_w2c(w2c: torch.Tensor, rho: torch.Tensor, theta: torch.Tensor) -> torch.Tensor:
    """w2c' = SE3_exp([rho, theta]) @ w2c for a batch [B,4,4] (fp32, CUDA)."""
    if w2c.device.type != "cuda":
        raise _lib.S3RError("styl3r_b200.pose needs CUDA tensors (no CPU fallback)")
    w2c = w2c.float().contiguous()
    rho, theta = rho.float().contiguous(), theta.float().contiguous()
    out = torch.empty_like(w2c)
    stream = C.c_void_p(torch.cuda.current_stream(w2c.device).cuda_stream)
    _lib.check(_lib.lib().s3r_se3_update_w2c(C.c_void_p(w2c.data_ptr()), C.c_void_p(rho.data_ptr()),
                                             C.c_void_p(theta.data_ptr()), C.c_void_p(out.data_ptr()),
                                             w2c.shape[0], stream), "s3r_se3_update_w2c")
    return out


def update_pose(cam_trans_delta: torch.Tensor, cam_rot_delta: torch.Tensor, extrinsics: torch.Tensor) -> torch.Tensor:
    """Same signature/semantics as the reference: extrinsics are camera-to-world [B,4,4]; returns the updated
    camera-to-world matrices  inverse(SE3_exp(tau) @ inverse(c2w))."""
    w2c = extrinsics.inverse()
    return se3_update_w2c(w2c, cam_trans_delta, cam_rot_delta).inverse()
