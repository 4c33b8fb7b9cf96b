"""bf16 NHWC convolutions of the DPT heads on the tcgen05 implicit-GEMM kernel (styl3r_b200/csrc/gemm_tcgen05.cu,
`s3r_conv2d_bf16`) plus the fused bilinear x2 upsampling (`s3r_upsample2x_nhwc_bf16`).

    wp = prep_conv_weight(conv.weight)                       # [Cout, KH*KW, ceil(Cin/64)*64] bf16, once per layer
    y  = conv2d_nhwc(x, wp, (KH, KW), bias=None, residual=None, relu=False)      # x, y: [N, H, W, C] bf16

replace `nn.Conv2d(k, stride 1, padding k//2)` (+ ReLU, + residual add) of heads/dpt_block.py:33-75,121-142,189-218,
dpt_head.py:35-70, dpt_gs_head.py:113-157 and dpt_gs_sh_head.py:37-74 (cuDNN in the reference).  CUDA only."""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _lib
from .gemm import EPI_BIAS, EPI_OUT_F32, EPI_RESIDUAL

EPI_RELU = 32


def _stream(device) -> C.c_void_p:
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def prep_conv_weight(weight: torch.Tensor) -> torch.Tensor:
    """nn.Conv2d weight [Cout, Cin, KH, KW] -> K-major implicit-GEMM operand [Cout, KH*KW, Cin_pad] bf16 (Cin zero
    padded to a multiple of 64: one k-block of the kernel = one tap x 64 channels)."""
    co, ci, kh, kw = weight.shape
    cp = (ci + 63) // 64 * 64
    w = torch.zeros(co, kh * kw, cp, dtype=torch.bfloat16, device=weight.device)
    w[:, :, :ci] = weight.detach().permute(0, 2, 3, 1).reshape(co, kh * kw, ci).to(torch.bfloat16)
    return w.contiguous()


def conv2d_nhwc(x: torch.Tensor, wprep: torch.Tensor, ksize: Tuple[int, int], bias: Optional[torch.Tensor] = None,
                residual: Optional[torch.Tensor] = None, relu: bool = False,
                out_dtype: torch.dtype = torch.bfloat16) -> torch.Tensor:
    """y = relu?(conv_same(x, w) + bias) + residual, NHWC.  ReLU is applied before the residual add."""
    if x.device.type != "cuda":
        raise _lib.S3RError("styl3r_b200.conv.conv2d_nhwc needs CUDA tensors (no CPU fallback)")
    if x.dtype != torch.bfloat16 or wprep.dtype != torch.bfloat16 or not x.is_contiguous() or not wprep.is_contiguous():
        raise _lib.S3RError("conv2d_nhwc expects contiguous bf16 NHWC activations and prepared bf16 weights")
    n, h, w, cin = x.shape
    kh, kw = ksize
    cout = wprep.shape[0]
    if wprep.shape[1] != kh * kw or wprep.shape[2] != (cin + 63) // 64 * 64:
        raise _lib.S3RError(f"prepared weight {tuple(wprep.shape)} does not match Cin={cin}, k={ksize}")
    y = torch.empty((n, h, w, cout), dtype=out_dtype, device=x.device)
    flags, bp, rp = 0, None, None
    if bias is not None:
        flags |= EPI_BIAS
        bias = bias if bias.dtype == torch.bfloat16 else bias.to(torch.bfloat16)
        bp = C.c_void_p(bias.data_ptr())
    if relu:
        flags |= EPI_RELU
    if residual is not None:
        if residual.shape != y.shape or residual.dtype != torch.bfloat16 or not residual.is_contiguous():
            raise _lib.S3RError("residual must be a contiguous bf16 NHWC tensor of the output shape")
        flags |= EPI_RESIDUAL
        rp = C.c_void_p(residual.data_ptr())
    if out_dtype == torch.float32:
        flags |= EPI_OUT_F32
    elif out_dtype != torch.bfloat16:
        raise _lib.S3RError("out_dtype must be bf16 or fp32")
    _lib.check(_lib.lib().s3r_conv2d_bf16(C.c_void_p(x.data_ptr()), C.c_void_p(wprep.data_ptr()), bp, rp,
                                          C.c_void_p(y.data_ptr()), n, h, w, cin, cout, kh, kw, kh // 2, flags,
                                          _stream(x.device)), "s3r_conv2d_bf16")
    return y


def upsample2x_nhwc(x: torch.Tensor, add: Optional[torch.Tensor] = None) -> torch.Tensor:
    """F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) on NHWC bf16 (+ `add`, fused)."""
    if x.device.type != "cuda" or x.dtype != torch.bfloat16 or not x.is_contiguous():
        raise _lib.S3RError("upsample2x_nhwc expects a contiguous bf16 NHWC CUDA tensor (no CPU fallback)")
    n, h, w, c = x.shape
    y = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.bfloat16, device=x.device)
    ap = None
    if add is not None:
        if add.shape != y.shape or add.dtype != torch.bfloat16 or not add.is_contiguous():
            raise _lib.S3RError("add must be a contiguous bf16 NHWC tensor of the output shape")
        ap = C.c_void_p(add.data_ptr())
    _lib.check(_lib.lib().s3r_upsample2x_nhwc_bf16(C.c_void_p(x.data_ptr()), ap, C.c_void_p(y.data_ptr()), n, h, w, c,
                                                   _stream(x.device)), "s3r_upsample2x_nhwc_bf16")
    return y
