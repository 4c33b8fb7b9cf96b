"""Tensor-level ops of the encoder with the reference's call signatures.

`memory_efficient_attention(q, k, v, scale=, p=)` mirrors `xformers.ops.memory_efficient_attention` as called at
src/model/encoder/backbone/croco/blocks.py:126-130,192-196: q [B,Nq,H,D], k/v [B,Nk,H,D] -> [B,Nq,H,D], no mask.
Round 1: the contraction itself is a library call (PyTorch SDPA -> cuDNN / flash kernels on sm_100), the tcgen05
kernel is the next row; RoPE, which the reference applies just before, is our own kernel.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F


def memory_efficient_attention(q: torch.Tensor, k: torch.Tensor, v: torch.Tensor, attn_bias=None, p: float = 0.0,
                               scale: float | None = None) -> torch.Tensor:
    if attn_bias is not None:
        raise NotImplementedError("attn_bias is not used by Styl3R")
    out = F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), dropout_p=p,
                                         scale=scale)
    return out.transpose(1, 2)
