"""DPT heads of Styl3R with the reference's parameter registry:

  <head>.dpt.act_postprocess.{0..3}   token -> feature-pyramid re-assembly          heads/dpt_block.py:365-412
  <head>.dpt.scratch.layer{1..4}_rn   (+ duplicate aliases scratch.layer_rn.{0..3})  heads/dpt_block.py:33-75
  <head>.dpt.scratch.refinenet{1..4}  fusion blocks (2 residual conv units + x2 + 1x1) heads/dpt_block.py:145-218
  <head>.dpt.head                     pts3d regression (dpt_head.py) / gs_params (dpt_gs_head.py, dpt_gs_sh_head.py)
  <head>.dpt.input_merger             7x7 image skip of the gs-parameter head        heads/dpt_gs_head.py:113-118

Hooks [0, 6, 9, 12] of the 13 decoder outputs, token dims (1024, 768, 768, 768), feature_dim 256.
"""
from __future__ import annotations

from typing import List, Sequence

import torch
import torch.nn.functional as F
from torch import Tensor, nn

HOOKS = (0, 6, 9, 12)
LAYER_DIMS = (96, 192, 384, 768)
FEATURE_DIM = 256


def _up2(x: Tensor) -> Tensor:
    return F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)


class ResidualConvUnit(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.conv1 = nn.Conv2d(c, c, 3, 1, 1, bias=True)
        self.conv2 = nn.Conv2d(c, c, 3, 1, 1, bias=True)

    def forward(self, x: Tensor) -> Tensor:
        return self.conv2(F.relu(self.conv1(F.relu(x)))) + x


class FusionBlock(nn.Module):
    def __init__(self, c: int):
        super().__init__()
        self.out_conv = nn.Conv2d(c, c, 1, bias=True)
        self.resConfUnit1 = ResidualConvUnit(c)
        self.resConfUnit2 = ResidualConvUnit(c)

    def forward(self, x: Tensor, skip: Tensor | None = None) -> Tensor:
        if skip is not None:
            x = x + self.resConfUnit1(skip)
        return self.out_conv(_up2(self.resConfUnit2(x)))


class _Up2(nn.Module):  # parameter-free slot 1 of the regression head (keeps the Sequential indices 0,2,4)
    def forward(self, x):
        return _up2(x)


class DPTAdapter(nn.Module):
    """`kind`: 'pts3d' (regression head, 128x128 -> x2 inside the head), 'gs_params' (image skip + x2 before the
    head) or 'gs_sh' (x2 before the head)."""

    def __init__(self, kind: str, out_channels: int, dim_tokens: Sequence[int] = (1024, 768, 768, 768)):
        super().__init__()
        assert kind in ("pts3d", "gs_params", "gs_sh")
        self.kind = kind
        L, Fd = LAYER_DIMS, FEATURE_DIM
        scratch = nn.Module()
        scratch.layer1_rn = nn.Conv2d(L[0], Fd, 3, 1, 1, bias=False)
        scratch.layer2_rn = nn.Conv2d(L[1], Fd, 3, 1, 1, bias=False)
        scratch.layer3_rn = nn.Conv2d(L[2], Fd, 3, 1, 1, bias=False)
        scratch.layer4_rn = nn.Conv2d(L[3], Fd, 3, 1, 1, bias=False)
        scratch.layer_rn = nn.ModuleList([scratch.layer1_rn, scratch.layer2_rn, scratch.layer3_rn, scratch.layer4_rn])
        scratch.refinenet1, scratch.refinenet2 = FusionBlock(Fd), FusionBlock(Fd)
        scratch.refinenet3, scratch.refinenet4 = FusionBlock(Fd), FusionBlock(Fd)
        self.scratch = scratch
        if kind == "pts3d":
            self.head = nn.Sequential(nn.Conv2d(Fd, Fd // 2, 3, 1, 1), _Up2(), nn.Conv2d(Fd // 2, Fd // 2, 3, 1, 1),
                                      nn.ReLU(True), nn.Conv2d(Fd // 2, out_channels, 1))
        else:
            self.head = nn.Sequential(nn.Conv2d(Fd, Fd, 3, padding=1, bias=False), nn.Identity(), nn.ReLU(True),
                                      nn.Dropout(0.1, False), nn.Conv2d(Fd, out_channels, 1))
        d = dim_tokens
        self.act_postprocess = nn.ModuleList([
            nn.Sequential(nn.Conv2d(d[0], L[0], 1), nn.ConvTranspose2d(L[0], L[0], 4, 4)),
            nn.Sequential(nn.Conv2d(d[1], L[1], 1), nn.ConvTranspose2d(L[1], L[1], 2, 2)),
            nn.Sequential(nn.Conv2d(d[2], L[2], 1)),
            nn.Sequential(nn.Conv2d(d[3], L[3], 1), nn.Conv2d(L[3], L[3], 3, 2, 1)),
        ])
        if kind == "gs_params":
            self.input_merger = nn.Sequential(nn.Conv2d(3, Fd, 7, 1, 3), nn.ReLU())

    def forward(self, tokens: List[Tensor], image_size, img: Tensor | None = None) -> Tensor:
        H, W = image_size
        nh, nw = H // 16, W // 16
        feats = []
        for i, hook in enumerate(HOOKS):
            t = tokens[hook]
            # [B, nh*nw, C] tokens viewed as NCHW are already a channels_last tensor: no copy
            x = t.contiguous().view(t.shape[0], nh, nw, t.shape[2]).permute(0, 3, 1, 2)
            feats.append(self.scratch.layer_rn[i](self.act_postprocess[i](x)))
        p = self.scratch.refinenet4(feats[3])[:, :, :feats[2].shape[2], :feats[2].shape[3]]
        p = self.scratch.refinenet3(p, feats[2])
        p = self.scratch.refinenet2(p, feats[1])
        p = self.scratch.refinenet1(p, feats[0])
        if self.kind == "gs_params":
            p = _up2(p) + self.input_merger(img)
        elif self.kind == "gs_sh":
            p = _up2(p)
        return self.head(p)


class PixelwiseDPT(nn.Module):
    """Holds the adapter under the attribute name `dpt` like the reference's PixelwiseTaskWithDPT."""

    def __init__(self, kind: str, out_channels: int):
        super().__init__()
        self.dpt = DPTAdapter(kind, out_channels)

    def forward(self, tokens: List[Tensor], image_size, img: Tensor | None = None) -> Tensor:
        return self.dpt(tokens, image_size, img)
