"""`DecoderSplattingCUDA` with the reference's constructor/forward contract
(src/model/decoder/decoder_splatting_cuda.py:15-68, decoder.py:18-45, __init__.py:11-12)."""
from __future__ import annotations

from dataclasses import dataclass
from typing import Any, Literal, Optional

import torch
from torch import Tensor, nn

from .cuda_splatting import DepthRenderingMode, render_cuda


@dataclass
class DecoderSplattingCUDACfg:
    name: Literal["splatting_cuda"]
    background_color: list
    make_scale_invariant: bool


@dataclass
class DecoderOutput:
    color: Tensor            # [batch, view, 3, height, width]
    depth: Optional[Tensor]  # [batch, view, height, width]


class DecoderSplattingCUDA(nn.Module):
    """forward(gaussians, extrinsics[b,v,4,4], intrinsics[b,v,3,3], near[b,v], far[b,v], (h,w), depth_mode,
    cam_rot_delta[b,v,3], cam_trans_delta[b,v,3]) -> DecoderOutput. `gaussians` is any object with
    means[b,G,3], covariances[b,G,3,3], harmonics[b,G,3,d_sh], opacities[b,G] (src/model/types.py:7-12).
    As in the reference, `depth_mode` is accepted and ignored (raw blended depth is returned)."""

    def __init__(self, cfg: DecoderSplattingCUDACfg) -> None:
        super().__init__()
        self.cfg = cfg
        self.make_scale_invariant = cfg.make_scale_invariant
        self.register_buffer("background_color", torch.tensor(cfg.background_color, dtype=torch.float32),
                             persistent=False)

    def forward(self, gaussians: Any, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor,
                image_shape: tuple[int, int], depth_mode: DepthRenderingMode | None = None,
                cam_rot_delta: Tensor | None = None, cam_trans_delta: Tensor | None = None) -> DecoderOutput:
        b, v = extrinsics.shape[:2]
        dev = extrinsics.device
        # view (i, j) renders Gaussian set i: no v-fold copies of the scene (the reference repeats them)
        view_set = torch.arange(b, device=dev, dtype=torch.int32).repeat_interleave(v)
        flat = lambda t: None if t is None else t.reshape(b * v, *t.shape[2:])
        color, depth = render_cuda(
            flat(extrinsics), flat(intrinsics), flat(near), flat(far), image_shape,
            self.background_color.to(dev).expand(b * v, 3), gaussians.means, gaussians.covariances,
            gaussians.harmonics, gaussians.opacities, scale_invariant=self.make_scale_invariant,
            cam_rot_delta=flat(cam_rot_delta), cam_trans_delta=flat(cam_trans_delta), view_set=view_set)
        h, w = image_shape
        return DecoderOutput(color.reshape(b, v, 3, h, w), depth.reshape(b, v, h, w))


DECODERS = {"splatting_cuda": DecoderSplattingCUDA}


def get_decoder(decoder_cfg) -> DecoderSplattingCUDA:
    return DECODERS[decoder_cfg.name](decoder_cfg)
