"""`xformers.ops` surface used by the reference: memory_efficient_attention(q, k, v, attn_bias=None, p=0.0, scale=None)
with q [B,Nq,H,D], k / v [B,Nk,H,D] -> [B,Nq,H,D] (blocks.py:126-130,192-196: no bias, p = 0, scale = head_dim^-0.5).

The reference hands over fp32 (TF32-matmul) tensors; the B200 kernels take bf16 operands with fp32 accumulation, so
inputs are rounded to bf16 on the way in and the result is returned in the caller's dtype.  Differentiable: autograd
runs `styl3r_b200.attention_bwd.attention_backward`.  CUDA only, head_dim 64 only - anything else raises (there is no
library fallback behind this module)."""
from __future__ import annotations

import torch

from ... import _lib
from ...encoder.train_ops import AttentionFn
from ...ops import attention_bf16


def memory_efficient_attention(query, key, value, attn_bias=None, p: float = 0.0, scale=None, **_):
    if attn_bias is not None or p != 0.0:
        raise NotImplementedError("styl3r_b200 xformers shim: attn_bias / dropout are not used by Styl3R")
    if not query.is_cuda or query.shape[-1] != 64:
        raise _lib.S3RError("styl3r_b200 xformers shim needs CUDA tensors with head_dim 64 (no CPU / library fallback)")
    sc = float(scale) if scale is not None else query.shape[-1] ** -0.5
    dt = query.dtype
    ok = lambda t: t.dtype == torch.bfloat16 and t.stride(-1) == 1 and all(s % 8 == 0 for s in t.stride()[:3])
    q, k, v = (t if ok(t) else t.to(torch.bfloat16).contiguous() for t in (query, key, value))
    if torch.is_grad_enabled() and any(t.requires_grad for t in (query, key, value)):
        out = AttentionFn.apply(q, k, v, sc)
    else:
        out = attention_bf16(q, k, v, sc)
    return out if dt == torch.bfloat16 else out.to(dt)
