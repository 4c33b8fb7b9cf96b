"""Style / identity losses of Styl3R's stage-2 training - mirrors of
`LossStyle` (src/loss/loss_style.py:24-80) and `IdentityLoss` (src/loss/loss_identity.py:13-50) with the same
constructor arguments, `forward(prediction, batch, gaussians, global_step)` signature and arithmetic:

    style    = MSE(relu3_1) + MSE(relu4_1) between prediction and target  +  style_weight * sum over the four VGG levels
               of MSE(channel mean) + MSE(channel std) between prediction and style image
    identity = 70 * MSE(prediction, target) + sum over the four levels of MSE(features)

Images are [0,1]; ImageNet normalisation happens inside (transforms.Normalize in the reference).  The VGG encoder
(`vgg.VGGEncoder`) is frozen and kept out of the state dict (the reference converts it to non-persistent buffers)."""
from __future__ import annotations

from dataclasses import dataclass

import torch
import torch.nn.functional as F
from torch import Tensor, nn

from .vgg import VGGEncoder, calc_mean_std

IMAGENET_MEAN = (0.485, 0.456, 0.406)
IMAGENET_STD = (0.229, 0.224, 0.225)


@dataclass
class LossStyleCfg:
    style_weight: float


@dataclass
class LossStyleCfgWrapper:
    style: LossStyleCfg


def _preprocess(img: Tensor) -> Tensor:
    mean = torch.as_tensor(IMAGENET_MEAN, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    std = torch.as_tensor(IMAGENET_STD, dtype=img.dtype, device=img.device).view(1, 3, 1, 1)
    return (img - mean) / std


class _FrozenVGG(nn.Module):
    """Holds the VGG encoder outside `parameters()` / `state_dict()` (reference: convert_to_buffer(persistent=False))."""

    def __init__(self, fast: bool):
        super().__init__()
        object.__setattr__(self, "_vgg", VGGEncoder(fast=fast))

    @property
    def vgg(self) -> VGGEncoder:
        return self._vgg

    def load_vgg19_features(self, source) -> "_FrozenVGG":
        """The VGG is deliberately outside state_dict() (like the reference's non-persistent buffers), so checkpoints do
        not carry it: load torchvision VGG-19 weights explicitly (state dict or .pth path)."""
        self._vgg.load_vgg19_features(source)
        return self

    def _apply(self, fn, recurse=True):
        self._vgg._apply(fn)
        self._vgg._prep = None
        return super()._apply(fn, recurse)


class LossStyle(_FrozenVGG):
    name = "style"

    def __init__(self, cfg: LossStyleCfgWrapper, fast: bool = True) -> None:
        super().__init__(fast)
        self.cfg = cfg.style

    def forward(self, prediction, batch: dict, gaussians=None, global_step: int = 0) -> Tensor:
        b, v = batch["target"]["image"].shape[:2]
        target_img = _preprocess(batch["target"]["image"].flatten(0, 1))
        pred_img = _preprocess(prediction.color.flatten(0, 1))
        style_img = _preprocess(batch["style"]["image"])
        style_img = style_img[:, None].expand(-1, v, -1, -1, -1).flatten(0, 1)
        pred_f = self.vgg(pred_img)
        with torch.no_grad():
            target_f = self.vgg(target_img)
            style_f = self.vgg(style_img)
        content_loss = F.mse_loss(pred_f[-2], target_f[-2]) + F.mse_loss(pred_f[-1], target_f[-1])
        style_loss = 0
        for pf, sf in zip(pred_f, style_f):
            pm, ps = calc_mean_std(pf)
            sm, ss = calc_mean_std(sf)
            style_loss = style_loss + F.mse_loss(pm, sm) + F.mse_loss(ps, ss)
        return content_loss + self.cfg.style_weight * style_loss


class IdentityLoss(_FrozenVGG):
    name = "identity"

    def __init__(self, weight_1: float = 70, weight_2: float = 1, fast: bool = True):
        super().__init__(fast)
        self.weight_1, self.weight_2 = weight_1, weight_2

    def forward(self, prediction, batch: dict, gaussians=None, global_step: int = 0) -> Tensor:
        target_img = batch["target"]["image"].flatten(0, 1)
        pred_img = prediction.color.flatten(0, 1)
        loss_identity1 = F.mse_loss(pred_img, target_img)
        pred_f = self.vgg(_preprocess(pred_img))
        with torch.no_grad():
            target_f = self.vgg(_preprocess(target_img))
        loss_identity2 = 0
        for pf, tf in zip(pred_f, target_f):
            loss_identity2 = loss_identity2 + F.mse_loss(pf, tf)
        return loss_identity1 * self.weight_1 + loss_identity2 * self.weight_2
