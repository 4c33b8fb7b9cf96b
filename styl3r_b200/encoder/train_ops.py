"""Autograd functions that keep the encoder's TRAINING path on our kernels (the reference runs
`ModelWrapperStyle.training_step`, src/model/model_wrapper_style.py:118-315, through torch autograd over cuBLAS / SDPA):

  LinearFn      y = x W^T + b (+ residual) (+ RoPE-2D on the q / k columns)   fwd: tcgen05 GEMM (fused epilogue)
                                                                               bwd: dgrad (B MN-major) + wgrad (A, B
                                                                               MN-major) on the same kernel, bias gradient
                                                                               as a ones-GEMM, inverse RoPE in place
  MlpFn         y = fc2(gelu(fc1 x)) + residual                                fwd: 2 GEMMs (GELU fused, pre-activation
                                                                               kept); bwd: gelu' fused into fc2's dgrad
  LayerNormFn   nn.LayerNorm(eps=1e-6)                                         s3r_layernorm_bf16 / s3r_layernorm_bwd_bf16
  AttentionFn   softmax(q k^T * scale) v                                       fwd: tcgen05 attention kernel; bwd:
                                                                               styl3r_b200.attention_bwd

Mixed precision like the inference layout: parameters stay fp32 (the optimiser's master copy), their bf16 operand
copies are cached per parameter version, activations are bf16, accumulation is fp32, weight / bias gradients are fp32.
Only tensors that require a gradient get one (`ctx.needs_input_grad`): frozen layers in front of a trainable one cost
a dgrad and no wgrad.
"""
from __future__ import annotations

import ctypes as C
import weakref
from typing import Optional

import torch
from torch import Tensor

from .. import _lib
from .. import gemm as _gemm
from ..curope import rope_2d

LN_EPS = 1e-6
_cast_cache: dict = {}
_ones_cache: dict = {}


def bf16_operand(p: Optional[Tensor]) -> Optional[Tensor]:
    """bf16 copy of a (fp32 master) parameter, re-made only when the optimiser has stepped (parameter ._version).
    Entries are tied to the parameter OBJECT through a weak reference: a freed parameter whose storage address is reused
    by another tensor can never produce a stale hit."""
    if p is None:
        return None
    if p.dtype == torch.bfloat16:
        return p.detach()
    key = id(p)
    hit = _cast_cache.get(key)
    if hit is None or hit[0]() is not p or hit[1] != p._version:
        if len(_cast_cache) > 8192:
            for k in [k for k, v in _cast_cache.items() if v[0]() is None]:
                del _cast_cache[k]
        hit = (weakref.ref(p), p._version, p.detach().to(torch.bfloat16).contiguous())
        _cast_cache[key] = hit
    return hit[2]


def _ones(M: int, device) -> Tensor:
    key = (str(device), M)
    t = _ones_cache.get(key)
    if t is None:
        if len(_ones_cache) > 64:
            _ones_cache.clear()
        t = torch.ones(M, 8, dtype=torch.bfloat16, device=device)
        _ones_cache[key] = t
    return t


def _bias_grad(dy2: Tensor) -> Tensor:
    """db[N] = sum over rows of dy [M, N], as a wgrad-shaped GEMM against a column of ones (fp32 accumulation in TMEM)."""
    M, N = dy2.shape
    return _gemm.gemm_majors(dy2, _ones(M, dy2.device), N, 8, M, True, True, out_dtype=torch.float32)[:, 0].contiguous()


def _as2d(t: Tensor) -> Tensor:
    t2 = t.reshape(-1, t.shape[-1])
    if t2.stride(1) != 1 or t2.stride(0) % 8 or t2.dtype != torch.bfloat16:
        t2 = t2.to(torch.bfloat16).contiguous()
    return t2


class LinearFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, residual, rope_pos, rope_cols, rope_base):
        w16, b16 = bf16_operand(weight), bf16_operand(bias)
        y = _gemm.linear(x, w16, b16, residual=residual, rope_pos=rope_pos, rope_cols=rope_cols, rope_base=rope_base)
        ctx.save_for_backward(x, weight, rope_pos)
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.rope = (rope_cols, rope_base)
        return y

    @staticmethod
    def backward(ctx, dy):
        x, weight, rope_pos = ctx.saved_tensors
        need_x, need_w, need_b, need_r = ctx.needs_input_grad[:4]
        N = weight.shape[0]
        dy2 = _as2d(dy)
        d_res = dy if (ctx.has_res and need_r) else None
        if rope_pos is not None:
            # d(rotated q, k) -> d(q, k): the rotation is orthogonal, its transpose is the rotation by -angle
            cols, base = ctx.rope
            dy2 = dy2.clone()
            M = dy2.shape[0]
            heads = cols // 64
            rope_2d(dy2[:, :cols].view(1, M, heads, 64), rope_pos.reshape(1, M, 2), base, -1.0)
        dx = dw = db = None
        if need_x:
            dx = _gemm.linear_dgrad(dy2, bf16_operand(weight)).view(x.shape)
        if need_w:
            dw = _gemm.linear_wgrad(dy2, _as2d(x)).to(weight.dtype)
        if ctx.has_bias and need_b:
            db = _bias_grad(dy2)
        return dx, dw, db, d_res, None, None, None


def linear(x: Tensor, layer: torch.nn.Linear, residual: Optional[Tensor] = None, rope_pos: Optional[Tensor] = None,
           rope_cols: int = 0, rope_base: float = 100.0) -> Tensor:
    return LinearFn.apply(x, layer.weight, layer.bias, residual, rope_pos, rope_cols, rope_base)


class MlpFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, w1, b1, w2, b2, residual):
        x2 = _as2d(x)
        M, hidden = x2.shape[0], w1.shape[0]
        h = torch.empty(M, hidden, dtype=torch.bfloat16, device=x.device)
        a = _gemm.gemm_majors(x2, bf16_operand(w1), M, hidden, x2.shape[1], False, False, bias=bf16_operand(b1), gelu=True,
                              pre_out=h)
        y = _gemm.linear(a, bf16_operand(w2), bf16_operand(b2), residual=None if residual is None else _as2d(residual))
        ctx.save_for_backward(x2, h, a, w1, w2)
        ctx.shape, ctx.has_res = x.shape, residual is not None
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, h, a, w1, w2 = ctx.saved_tensors
        nx, nw1, nb1, nw2, nb2, nr = ctx.needs_input_grad
        dy2 = _as2d(dy)
        dw2 = _gemm.linear_wgrad(dy2, a).to(w2.dtype) if nw2 else None
        db2 = _bias_grad(dy2) if nb2 else None
        dh = _gemm.linear_dgrad(dy2, bf16_operand(w2), pre_gelu=h)      # d(gelu input), gelu' fused in the epilogue
        dw1 = _gemm.linear_wgrad(dh, x2).to(w1.dtype) if nw1 else None
        db1 = _bias_grad(dh) if nb1 else None
        dx = _gemm.linear_dgrad(dh, bf16_operand(w1)).view(ctx.shape) if nx else None
        return dx, dw1, db1, dw2, db2, (dy if (ctx.has_res and nr) else None)


def mlp(x: Tensor, fc1: torch.nn.Linear, fc2: torch.nn.Linear, residual: Optional[Tensor] = None) -> Tensor:
    return MlpFn.apply(x, fc1.weight, fc1.bias, fc2.weight, fc2.bias, residual)


class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        C_ = x.shape[-1]
        x2 = _as2d(x)
        w16, b16 = bf16_operand(weight), bf16_operand(bias)
        y = torch.empty((x2.shape[0], C_), dtype=torch.bfloat16, device=x.device)
        st = C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)
        _lib.check(_lib.lib().s3r_layernorm_bf16(C.c_void_p(x2.data_ptr()), C.c_void_p(w16.data_ptr()), C.c_void_p(b16.data_ptr()),
                                                 C.c_void_p(y.data_ptr()), x2.shape[0], C_, x2.stride(0), float(eps), st),
                   "s3r_layernorm_bf16")
        ctx.save_for_backward(x2, weight)
        ctx.eps, ctx.shape = eps, x.shape
        return y.view(x.shape)

    @staticmethod
    def backward(ctx, dy):
        x2, weight = ctx.saved_tensors
        nx, nw, nb, _ = ctx.needs_input_grad
        C_ = x2.shape[1]
        dy2 = dy.reshape(-1, C_)
        if dy2.dtype != torch.bfloat16 or not dy2.is_contiguous():
            dy2 = dy2.to(torch.bfloat16).contiguous()
        dx = torch.empty_like(dy2)
        dw = torch.zeros(C_, dtype=torch.float32, device=dy.device) if nw else None
        db = torch.zeros(C_, dtype=torch.float32, device=dy.device) if nb else None
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
        st = C.c_void_p(torch.cuda.current_stream(dy.device).cuda_stream)
        _lib.check(_lib.lib().s3r_layernorm_bwd_bf16(p(x2), p(bf16_operand(weight)), p(dy2), p(dx), p(dw), p(db), x2.shape[0], C_,
                                                     x2.stride(0), float(ctx.eps), st), "s3r_layernorm_bwd_bf16")
        return (dx.view(ctx.shape) if nx else None, None if dw is None else dw.to(weight.dtype),
                None if db is None else db.to(weight.dtype), None)


def layer_norm(x: Tensor, norm: torch.nn.LayerNorm) -> Tensor:
    return LayerNormFn.apply(x, norm.weight, norm.bias, norm.eps)


class AttentionFn(torch.autograd.Function):
    """q [B,Nq,H,64], k / v [B,Nk,H,64] (strided views of the packed projections) -> [B,Nq,H,64]."""

    @staticmethod
    def forward(ctx, q, k, v, scale):
        from ..ops import attention_bf16
        o = attention_bf16(q, k, v, scale)
        ctx.save_for_backward(q, k, v, o)
        ctx.scale = scale
        return o

    @staticmethod
    def backward(ctx, do):
        from ..attention_bwd import attention_backward
        q, k, v, o = ctx.saved_tensors
        dq, dk, dv = attention_backward(q, k, v, o, do, ctx.scale)
        return dq, dk, dv, None


def attention(q: Tensor, k: Tensor, v: Tensor, scale: float) -> Tensor:
    return AttentionFn.apply(q, k, v, scale)


def supported(x: Tensor) -> bool:
    """The bf16 training path needs CUDA bf16 activations (mode selected by `encoder.to_training()`)."""
    return x.is_cuda and x.dtype == torch.bfloat16
