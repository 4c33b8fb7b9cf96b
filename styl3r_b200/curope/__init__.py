"""Drop-in for the reference's `curope` extension module and its Python wrapper
(src/model/encoder/backbone/croco/curope/curope.cpp:49-69 `rope_2d`; curope2d.py:12-40 `cuRoPE2D_func`,
`cuRoPE2D`).  Same names, argument meaning and error behaviour; the work is done by the sm_100a kernel behind
`s3r_rope2d` (styl3r_b200/csrc/rope2d.cu).  CUDA tensors only — there is no CPU path here (the reference's
`rope_2d_cpu` lives on as the oracle, oracle/rope_oracle.c)."""
from __future__ import annotations

import ctypes as C

import torch

from .. import _lib

_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2}


def rope_2d(tokens: torch.Tensor, positions: torch.Tensor, base: float, fwd: float) -> None:
    """In-place RoPE-2D on tokens[B,N,H,D] (any strides with a unit innermost stride) with positions[B,N,2]
    (int64, (y, x)).  `fwd = +F0` applies the rotation, `-F0` its inverse (the backward pass)."""
    if tokens.dim() != 4:
        raise RuntimeError("tokens must have 4 dimensions")
    if positions.dim() != 3:
        raise RuntimeError("positions must have 3 dimensions")
    if tokens.size(0) != positions.size(0):
        raise RuntimeError("batch size differs between tokens & positions")
    if tokens.size(1) != positions.size(1):
        raise RuntimeError("seq_length differs between tokens & positions")
    if positions.size(2) != 2:
        raise RuntimeError("positions.shape[2] must be equal to 2")
    if tokens.is_cuda != positions.is_cuda:
        raise RuntimeError("tokens and positions are not on the same device")
    if not tokens.is_cuda:
        raise RuntimeError("styl3r_b200.curope.rope_2d: CUDA tensors only (no CPU fallback in the product path)")
    if tokens.stride(3) != 1:
        raise RuntimeError("tokens are not contiguous")
    if not positions.is_contiguous():
        raise RuntimeError("positions are not contiguous")
    if tokens.size(3) % 4 != 0:
        raise RuntimeError("token dim must be multiple of 4")
    if tokens.dtype not in _DTYPES:
        raise RuntimeError(f"unsupported dtype {tokens.dtype}")
    if positions.dtype != torch.int64:
        raise RuntimeError("positions must be int64")
    B, N, H, D = tokens.shape
    if tokens.numel() == 0:
        return
    stream = C.c_void_p(torch.cuda.current_stream(tokens.device).cuda_stream)
    _lib.check(_lib.lib().s3r_rope2d(C.c_void_p(tokens.data_ptr()), C.c_void_p(positions.data_ptr()), B, N, H, D,
                                     tokens.stride(0), tokens.stride(1), tokens.stride(2), float(base), float(fwd),
                                     _DTYPES[tokens.dtype], stream), "s3r_rope2d")


class cuRoPE2D_func(torch.autograd.Function):
    @staticmethod
    def forward(ctx, tokens, positions, base, F0=1):
        ctx.save_for_backward(positions)
        ctx.saved_base, ctx.saved_F0 = base, F0
        rope_2d(tokens, positions, base, F0)
        ctx.mark_dirty(tokens)
        return tokens

    @staticmethod
    def backward(ctx, grad_res):
        positions, base, F0 = ctx.saved_tensors[0], ctx.saved_base, ctx.saved_F0
        grad_res = grad_res.contiguous() if grad_res.stride(3) != 1 else grad_res
        rope_2d(grad_res, positions, base, -F0)
        ctx.mark_dirty(grad_res)
        return grad_res, None, None, None


class cuRoPE2D(torch.nn.Module):
    """tokens: [B, heads, N, D]; positions: [B, N, 2] -> tokens (rotated in place), like the reference module."""

    def __init__(self, freq: float = 100.0, F0: float = 1.0):
        super().__init__()
        self.base = freq
        self.F0 = F0

    def forward(self, tokens, positions):
        cuRoPE2D_func.apply(tokens.transpose(1, 2), positions, self.base, self.F0)
        return tokens
