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


# ----------------------------------------------------------------------------------------------------------------------
# B200 inference path: the whole pyramid in bf16 NHWC on the tcgen05 kernels (tokens are already NHWC, so no layout
# change anywhere).  3x3 convolutions -> implicit GEMM (`conv.conv2d_nhwc`); 1x1 convolutions and the kernel==stride
# transposed convolutions -> plain GEMM (`gemm.linear`) on the pixel-major view; the 7x7 image skip and the one
# strided 3x3 (2 % of the head FLOPs) -> im2col + GEMM; bias / ReLU / residual adds fused into the epilogues; bilinear
# x2 (+ add) -> `conv.upsample2x_nhwc`.  Returns the head output as fp32 pixel-major rows [B*H*W, 8].
class _PreparedHead:
    """bf16 GEMM operands derived from a DPTAdapter's parameters (not registered: the state dict is untouched)."""

    def __init__(self, m: "DPTAdapter"):
        from ..conv import prep_conv_weight
        bf = lambda t: None if t is None else t.detach().to(torch.bfloat16).contiguous()
        self.device = m.scratch.layer1_rn.weight.device
        self.pp = []
        for i, seq in enumerate(m.act_postprocess):
            c0 = seq[0]
            e = {"w0": bf(c0.weight.flatten(1)), "b0": bf(c0.bias)}
            if i in (0, 1):  # ConvTranspose2d(k = stride): out[(i, j, co)] = sum_ci x[ci] * W[ci, co, i, j]
                ct = seq[1]
                k = ct.kernel_size[0]
                e["k"] = k
                e["wt"] = bf(ct.weight.permute(2, 3, 1, 0).reshape(k * k * ct.out_channels, ct.in_channels))
                e["bt"] = bf(ct.bias.repeat(k * k))
            elif i == 3:     # Conv2d(3, stride 2, pad 1) through im2col: columns ordered (ci, kh, kw) like F.unfold
                e["ws"] = bf(seq[1].weight.flatten(1))
                e["bs"] = bf(seq[1].bias)
            self.pp.append(e)
        self.rn = [prep_conv_weight(c.weight) for c in m.scratch.layer_rn]
        self.fusion = []
        for blk in (m.scratch.refinenet1, m.scratch.refinenet2, m.scratch.refinenet3, m.scratch.refinenet4):
            f = {"out_w": bf(blk.out_conv.weight.flatten(1)), "out_b": bf(blk.out_conv.bias)}
            for name in ("resConfUnit1", "resConfUnit2"):
                u = getattr(blk, name)
                f[name] = (prep_conv_weight(u.conv1.weight), bf(u.conv1.bias), prep_conv_weight(u.conv2.weight), bf(u.conv2.bias))
            self.fusion.append(f)
        h = m.head
        self.h0_w, self.h0_b = prep_conv_weight(h[0].weight), bf(h[0].bias)
        if m.kind == "pts3d":
            self.h2_w, self.h2_b = prep_conv_weight(h[2].weight), bf(h[2].bias)
        last = h[4]
        ld = (last.out_channels + 7) // 8 * 8  # pad Cout (3, 8, 3*d_sh) to a multiple of 8 rows (16-byte output rows)
        w = torch.zeros(ld, last.in_channels, dtype=torch.bfloat16, device=self.device)
        b = torch.zeros(ld, dtype=torch.bfloat16, device=self.device)
        w[:last.out_channels] = last.weight.detach().flatten(1).to(torch.bfloat16)
        b[:last.out_channels] = last.bias.detach().to(torch.bfloat16)
        self.last_w, self.last_b = w, b
        if m.kind == "gs_params":
            mw = m.input_merger[0].weight.detach().flatten(1)  # [256, 147], columns (ci, kh, kw)
            self.mg_w = torch.zeros(mw.shape[0], 152, dtype=torch.bfloat16, device=self.device)
            self.mg_w[:, :mw.shape[1]] = mw.to(torch.bfloat16)
            self.mg_b = bf(m.input_merger[0].bias)


def _rcu(x: Tensor, u) -> Tensor:
    from ..conv import conv2d_nhwc
    w1, b1, w2, b2 = u
    t = conv2d_nhwc(torch.relu(x), w1, (3, 3), bias=b1, relu=True)
    return conv2d_nhwc(t, w2, (3, 3), bias=b2, residual=x)


def _fusion(f, x: Tensor, skip: Tensor | None = None) -> Tensor:
    from ..conv import conv2d_nhwc, upsample2x_nhwc
    from ..gemm import linear
    if skip is not None:  # x + rcu1(skip) = conv2(...) + (skip + x)
        w1, b1, w2, b2 = f["resConfUnit1"]
        t = conv2d_nhwc(torch.relu(skip), w1, (3, 3), bias=b1, relu=True)
        x = conv2d_nhwc(t, w2, (3, 3), bias=b2, residual=skip + x)
    # out_conv (1x1) BEFORE the bilinear x2: a per-pixel channel mix commutes with an interpolation whose weights sum to
    # one (bias included), so the GEMM runs on a quarter of the pixels (reference order: heads/dpt_block.py:189-218)
    t = _rcu(x, f["resConfUnit2"])
    return upsample2x_nhwc(linear(t, f["out_w"], f["out_b"]))


def dpt_forward_nhwc(m: "DPTAdapter", tokens: List[Tensor], image_size, img: Tensor | None = None) -> Tensor:
    from ..conv import conv2d_nhwc, upsample2x_nhwc
    from ..gemm import linear
    prep = getattr(m, "_prep", None)
    if prep is None or prep.device != tokens[0].device:
        prep = m._prep = _PreparedHead(m)
    H, W = image_size
    nh, nw = H // 16, W // 16
    feats = []
    for i, hook in enumerate(HOOKS):
        t = tokens[hook]
        B = t.shape[0]
        e = prep.pp[i]
        x = linear(t.reshape(B * nh * nw, t.shape[-1]), e["w0"], e["b0"])          # 1x1 conv
        if i in (0, 1):
            k, co = e["k"], LAYER_DIMS[i]
            x = linear(x, e["wt"], e["bt"])                                         # [B*nh*nw, k*k*co]
            x = x.view(B, nh, nw, k, k, co).permute(0, 1, 3, 2, 4, 5).reshape(B, nh * k, nw * k, co)
        elif i == 2:
            x = x.view(B, nh, nw, -1)
        else:
            cols = F.unfold(x.view(B, nh, nw, -1).permute(0, 3, 1, 2), 3, padding=1, stride=2)   # [B, C*9, L]
            oh, ow = (nh + 1) // 2, (nw + 1) // 2
            x = linear(cols.transpose(1, 2).reshape(B * oh * ow, -1), e["ws"], e["bs"]).view(B, oh, ow, -1)
        feats.append(conv2d_nhwc(x.contiguous(), prep.rn[i], (3, 3)))
    f1, f2, f3, f4 = prep.fusion
    p = _fusion(f4, feats[3])[:, :feats[2].shape[1], :feats[2].shape[2]].contiguous()
    p = _fusion(f3, p, feats[2])
    p = _fusion(f2, p, feats[1])
    p = _fusion(f1, p, feats[0])
    if m.kind == "pts3d":
        p = conv2d_nhwc(p, prep.h0_w, (3, 3), bias=prep.h0_b)
        p = conv2d_nhwc(upsample2x_nhwc(p), prep.h2_w, (3, 3), bias=prep.h2_b, relu=True)
    else:
        add = None
        if m.kind == "gs_params":
            # patch matrix [B*H*W, 152] (columns (ci, kh, kw) like F.unfold, zero-padded to a multiple of 8) in one pass -
            # F.unfold + transpose copy + F.pad were three passes over it (1.4 ms for the 12 images of cfg3)
            import ctypes as C
            from .. import _lib
            B = img.shape[0]
            imb = img.to(torch.bfloat16).contiguous()
            cols = torch.empty(B * H * W, 152, dtype=torch.bfloat16, device=img.device)
            _lib.check(_lib.lib().s3r_im2col7x7_bf16(C.c_void_p(imb.data_ptr()), C.c_void_p(cols.data_ptr()), B, H, W,
                                                      C.c_void_p(torch.cuda.current_stream(img.device).cuda_stream)),
                       "s3r_im2col7x7_bf16")
            add = linear(cols, prep.mg_w, prep.mg_b, relu=True).view(B, H, W, -1)
        p = conv2d_nhwc(upsample2x_nhwc(p, add), prep.h0_w, (3, 3), relu=True)
    return linear(p.reshape(-1, p.shape[-1]), prep.last_w, prep.last_b, out_dtype=torch.float32)   # [B*H*W, ld] fp32


def nhwc_supported(image_size) -> bool:
    """The implicit-GEMM kernel tiles 128 consecutive pixels as a box: every pyramid width must divide or be a multiple
    of 128 (true for the 256x256 production resolution and any power-of-two size >= 128)."""
    H, W = image_size
    if H % 16 or W % 16 or min(H, W) < 256:   # below the production resolution the fp32 module path is used
        return False
    ok = True
    for div in (32, 16, 8, 4, 2, 1):
        w, h = W // div, H // div
        if w == 0 or h == 0:
            return False
        if w >= 128:
            ok &= w % 128 == 0
        elif 128 % w:
            return False
        else:
            rows = 128 // w
            ok &= (h % rows == 0) if rows <= h else (rows % h == 0)
    return ok


class PixelwiseDPT(nn.Module):
    """Holds the adapter under the attribute name `dpt` like the reference's PixelwiseTaskWithDPT."""

    def __init__(self, kind: str, out_channels: int):
        super().__init__()
        self.dpt = DPTAdapter(kind, out_channels)

    def forward(self, tokens: List[Tensor], image_size, img: Tensor | None = None) -> Tensor:
        return self.dpt(tokens, image_size, img)

    def forward_nhwc(self, tokens: List[Tensor], image_size, img: Tensor | None = None) -> Tensor:
        return dpt_forward_nhwc(self.dpt, tokens, image_size, img)
