"""Tensor-level ops of the encoder with the reference's call signatures.

`memory_efficient_attention(q, k, v, scale=, p=)` mirrors `xformers.ops.memory_efficient_attention` as called at
src/model/encoder/backbone/croco/blocks.py:126-130,192-196: q [B,Nq,H,D], k/v [B,Nk,H,D] -> [B,Nq,H,D], no mask.
bf16 inference layout (no autograd): our tcgen05/TMEM kernel `s3r_attention_bf16` (head_dim 64, strided q/k/v views,
no copies).  fp32 / training: PyTorch SDPA (library) — the reference's numerics path used by the fp32 parity tests.
"""
from __future__ import annotations

import ctypes as C

import torch
import torch.nn.functional as F

from . import _lib

# Benchmark comparator switch (baseline/library_encoder.py): True routes the bf16 inference layout through the torch
# library ops (cuBLAS / SDPA / ATen LayerNorm) instead of the tcgen05 kernels, so that bench.py can time "the same module
# tree on the libraries" next to ours on the same box.  Never set by the product.
FORCE_LIBRARY = False


def memory_efficient_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_bias=None, p: float = 0.0,
                               scale: float | None = None) -> torch.Tensor:
    if attn_bias is not None:
        raise NotImplementedError("attn_bias is not used by Styl3R")
    if (not FORCE_LIBRARY and q.dtype == torch.bfloat16 and q.shape[-1] == 64 and p == 0.0 and q.is_cuda and not torch.is_grad_enabled()
            and all(t.stride(-1) == 1 and all(s % 8 == 0 for s in t.stride()[:3]) for t in (q, k, v))):
        return attention_bf16(q, k, v, scale if scale is not None else q.shape[-1] ** -0.5)
    out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), dropout_p=p,
                                         scale=scale)
    return out.transpose(1, 2)


def attention_bf16(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, scale: float) -> torch.Tensor:
    """q [B,Nq,H,64], k/v [B,Nk,H,64] bf16 (any batch/token/head strides) -> [B,Nq,H,64] bf16, on tcgen05."""
    B, Nq, H, D = q.shape
    Nk = k.shape[1]
    out = torch.empty((B, Nq, H, D), dtype=torch.bfloat16, device=q.device)
    arr = lambda t: (C.c_int64 * 3)(*t.stride()[:3])
    _lib.check(_lib.lib().s3r_attention_bf16(
        C.c_void_p(q.data_ptr()), C.c_void_p(k.data_ptr()), C.c_void_p(v.data_ptr()), C.c_void_p(out.data_ptr()), B, H, Nq,
        Nk, D, arr(q), arr(k), arr(v), arr(out), float(scale),
        C.c_void_p(torch.cuda.current_stream(q.device).cuda_stream)), "s3r_attention_bf16")
    return out
