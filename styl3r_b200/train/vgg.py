"""VGG-19 feature encoder of the style / identity losses - mirror of `VGGEncoder` and
`calc_mean_std` (src/test/vgg_model.py:19-28,79-98): features relu1_1, relu2_1, relu3_1, relu4_1 of ImageNet-normalised
images.  The four slices keep torchvision's `vgg19().features` indices as sub-module names (`slice1.0`, `slice2.2`,
`slice2.5`, `slice3.7`, `slice3.10`, `slice4.12/14/16/19`).  The reference builds them from
`torchvision.models.vgg19(pretrained=True)` (a download) and registers them as NON-persistent buffers, so its
checkpoints do NOT contain them; here they are loaded explicitly with `load_vgg19_features()` (a torchvision `vgg19`
state dict / `.pth` file, keys `features.N.weight|bias` or `N.weight|bias`).  Running the encoder with weights that
were never loaded raises a warning once: the losses would be computed on random features.

B200 path (`fast=True`, CUDA): every 3x3 convolution (+ bias + ReLU) is ONE launch of the tcgen05 implicit-GEMM kernel
(`conv.conv2d_nhwc`) in bf16 NHWC with fp32 accumulation, and so is its backward: the weights are frozen, so only the
data gradient is needed, and dgrad of a stride-1 "same" convolution is the same convolution with the spatially flipped,
channel-transposed filter applied to dL/dy masked by the ReLU.  conv1_1 (3 input channels) runs as im2col + GEMM.
The reference runs cuDNN fp32/TF32 NCHW convolutions through autograd."""
from __future__ import annotations

from typing import List

import torch
import torch.nn.functional as F
from torch import Tensor, nn

VGG19_CFG = [(0, 3, 64), (2, 64, 64), "M", (5, 64, 128), (7, 128, 128), "M", (10, 128, 256), (12, 256, 256),
             (14, 256, 256), (16, 256, 256), "M", (19, 256, 512)]
SLICES = ((0, 2), (2, 7), (7, 12), (12, 21))


def calc_mean_std(x: Tensor, eps: float = 1e-8):
    """Channel-wise instance mean / std over the flattened spatial dims (vgg_model.py:19-28): x [N, C, *]."""
    mean = torch.mean(x.flatten(2), dim=-1, keepdim=True)
    std = torch.std(x.flatten(2), dim=-1, keepdim=True) + eps
    return mean, std


def _vgg19_features_to_relu4_1() -> nn.Sequential:
    layers: List[nn.Module] = []
    for item in VGG19_CFG:
        if item == "M":
            layers.append(nn.MaxPool2d(kernel_size=2, stride=2))
        else:
            _, cin, cout = item
            layers += [nn.Conv2d(cin, cout, kernel_size=3, padding=1), nn.ReLU(inplace=True)]
    return nn.Sequential(*layers)  # indices 0..20 equal torchvision's vgg19().features[:21]


def _fast_supported(H: int, W: int) -> bool:
    """All four pyramid levels must satisfy the implicit-GEMM tiling rule (128-pixel tile = a box of the tensor)."""
    if H % 8 or W % 8 or min(H, W) < 128:   # small images: fp32 torch path
        return False
    for d in (1, 2, 4, 8):
        h, w = H // d, W // d
        if w >= 128:
            if w % 128:
                return False
        else:
            rows = 128 // w
            if 128 % w or not ((rows <= h and h % rows == 0) or (rows > h and rows % h == 0)):
                return False
    return True


class _ConvReLU(torch.autograd.Function):
    """y = relu(conv3x3_same(x) + b) on NHWC bf16 through the tcgen05 kernel; backward = dgrad only (frozen weights)."""

    @staticmethod
    def forward(ctx, x, w_fwd, w_bwd, bias):
        from ..conv import conv2d_nhwc
        y = conv2d_nhwc(x, w_fwd, (3, 3), bias=bias, relu=True)
        ctx.save_for_backward(y, w_bwd)
        return y

    @staticmethod
    def backward(ctx, gy):
        from ..conv import conv2d_nhwc
        y, w_bwd = ctx.saved_tensors
        g = torch.where(y > 0, gy.to(torch.bfloat16), torch.zeros((), dtype=torch.bfloat16, device=gy.device)).contiguous()
        return conv2d_nhwc(g, w_bwd, (3, 3)), None, None, None


class VGGEncoder(nn.Module):
    def __init__(self, fast: bool = True):
        super().__init__()
        vgg = _vgg19_features_to_relu4_1()
        self.slice1, self.slice2, self.slice3, self.slice4 = (vgg[a:b] for a, b in SLICES)
        self.requires_grad_(False)
        self.fast = fast
        self._prep = None
        self._loaded = False
        self._warned = False
        # any load_state_dict() (ours or a parent module's) invalidates the cached bf16 operands of the fast path
        self.register_load_state_dict_post_hook(lambda module, incompatible: module._invalidate(loaded=True))

    def _invalidate(self, loaded: bool = False):
        self._prep = None
        if loaded:
            self._loaded = True

    def _param_versions(self):
        return tuple(p._version for p in self.parameters())

    def load_vgg19_features(self, source) -> "VGGEncoder":
        """Load the 9 convolutions up to relu4_1 from a torchvision VGG-19 state dict (or a path to one):
        keys `features.{0,2,5,7,10,12,14,16,19}.{weight,bias}` (full model) or `{0,...}.{weight,bias}` (`.features`)."""
        sd = torch.load(source, map_location="cpu") if isinstance(source, (str, bytes)) or hasattr(source, "__fspath__") else source
        sd = {k[len("features."):] if k.startswith("features.") else k: v for k, v in sd.items()}
        mods = {}
        for sl, (a, b) in zip((self.slice1, self.slice2, self.slice3, self.slice4), SLICES):
            for idx, m in zip(range(a, b), sl):
                if isinstance(m, nn.Conv2d):
                    mods[idx] = m
        missing = [f"{i}.{n}" for i in mods for n in ("weight", "bias") if f"{i}.{n}" not in sd]
        if missing:
            raise KeyError(f"VGG-19 state dict lacks {missing}")
        with torch.no_grad():
            for i, m in mods.items():
                m.weight.copy_(sd[f"{i}.weight"])
                m.bias.copy_(sd[f"{i}.bias"])
        self._invalidate(loaded=True)
        return self

    def mark_weights_loaded(self) -> None:
        """For callers that fill the parameters by other means (tests: name-derived weights)."""
        self._invalidate(loaded=True)

    # ---- reference semantics (fp32 torch ops; also the numerics reference of the fast path)
    def forward_reference(self, images: Tensor, output_last_feature: bool = False):
        h1 = self.slice1(images)
        h2 = self.slice2(h1)
        h3 = self.slice3(h2)
        h4 = self.slice4(h3)
        return h4 if output_last_feature else (h1, h2, h3, h4)

    # ---- B200 path
    def _prepare(self, device):
        from ..conv import prep_conv_weight
        convs = [m for s in (self.slice1, self.slice2, self.slice3, self.slice4) for m in s if isinstance(m, nn.Conv2d)]
        prep = []
        for i, c in enumerate(convs):
            w = c.weight.detach().to(device)
            b = c.bias.detach().to(device, torch.bfloat16).contiguous()
            if i == 0:  # im2col GEMM operands: columns (ci, kh, kw) like F.unfold, K padded 27 -> 32
                wm = torch.zeros(w.shape[0], 32, dtype=torch.bfloat16, device=device)
                wm[:, :27] = w.flatten(1).to(torch.bfloat16)
                prep.append((wm, b))
            else:
                w_bwd = w.flip(2, 3).transpose(0, 1).contiguous()  # dgrad filter: [Cin, Cout, kh, kw], flipped
                prep.append((prep_conv_weight(w), prep_conv_weight(w_bwd), b))
        self._prep = (device, prep, self._param_versions())

    def _forward_fast(self, images: Tensor):
        from ..gemm import linear
        if self._prep is None or self._prep[0] != images.device or self._prep[2] != self._param_versions():
            self._prepare(images.device)
        prep = self._prep[1]
        B, _, H, W = images.shape
        # conv1_1: im2col (differentiable: its backward is F.fold = col2im) + tcgen05 GEMM with bias + ReLU
        cols = F.pad(F.unfold(images.to(torch.bfloat16), 3, padding=1).transpose(1, 2), (0, 5)).reshape(B * H * W, 32)
        x = _Linear1.apply(cols, prep[0][0], prep[0][1]).view(B, H, W, -1)
        feats = [x]
        # conv indices 1..8 = conv1_2, 2_1, 2_2, 3_1, 3_2, 3_3, 3_4, 4_1;  slice2: conv1_2, pool, conv2_1 | slice3: conv2_2, pool, conv3_1 | slice4: conv3_2..4, pool, conv4_1
        plan = ((1, "M", 2), (3, "M", 4), (5, 6, 7, "M", 8))
        for sl in plan:
            for op in sl:
                if op == "M":
                    x = F.max_pool2d(x.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1).contiguous()
                else:
                    x = _ConvReLU.apply(x, prep[op][0], prep[op][1], prep[op][2])
            feats.append(x)
        return feats  # NHWC bf16

    def forward(self, images: Tensor, output_last_feature: bool = False):
        """Returns NCHW feature maps like the reference (views of the NHWC buffers on the fast path, fp32)."""
        if not self._loaded and not self._warned:
            import warnings
            warnings.warn("VGGEncoder is running with weights that were never loaded (random initialisation): style / "
                          "identity losses are meaningless until load_vgg19_features() or load_state_dict() is called",
                          RuntimeWarning, stacklevel=2)
            self._warned = True
        if self.fast and images.is_cuda and _fast_supported(*images.shape[-2:]):
            feats = [f.permute(0, 3, 1, 2).float() for f in self._forward_fast(images)]
            return feats[-1] if output_last_feature else tuple(feats)
        return self.forward_reference(images, output_last_feature)


class _Linear1(torch.autograd.Function):
    """conv1_1 as relu(cols @ W^T + b) on the tcgen05 GEMM; backward: dcols = (dy * mask) @ W."""

    @staticmethod
    def forward(ctx, cols, wm, bias):
        from ..gemm import linear
        y = linear(cols, wm, bias, relu=True)
        ctx.save_for_backward(y, wm)
        return y

    @staticmethod
    def backward(ctx, gy):
        from ..gemm import linear
        y, wm = ctx.saved_tensors
        g = torch.where(y > 0, gy.to(torch.bfloat16), torch.zeros((), dtype=torch.bfloat16, device=gy.device)).contiguous()
        return linear(g, wm.t().contiguous()), None, None
