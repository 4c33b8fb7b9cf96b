"""Backward of `memory_efficient_attention` (croco/blocks.py:126-130,192-196: softmax(q k^T * scale) v, no mask) on our
kernels - `s3r_attention_bwd` of SURVEY.md §8b (B-attn).  Unfused formulation on the tensor cores:

    S  = Q K^T                      batched tcgen05 GEMM (fp32 out)        [BH, Nq, Nk]
    P  = softmax(scale * S)         s3r_softmax_rows_bf16
    dP = dO V^T                     batched GEMM (fp32 out)
    dS = scale * P o (dP - rowsum(dO o O))    s3r_attention_ds_bf16
    dV = P^T dO,  dK = dS^T Q       batched GEMMs, both operands MN-major (no transposes)
    dQ = dS K                       batched GEMM, B MN-major

Five launches + two row kernels per attention module; the scores are materialised (fp32 S / dP, bf16 P / dS - at most
4 x 120 x 514 x 520 elements for the stylizer decoder at batch 10), which a fused flash-style backward would avoid - the
contraction FLOPs of the backward are ~4 % of the encoder's, so the unfused form costs bandwidth, not tensor time.
"""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib

EPI_OUT_F32 = 8


def _bgemm(a, b, c, M, N, K, lda, ldb, ldc, sa, sb, sc, batch, a_mn, b_mn):
    flags = EPI_OUT_F32 if c.dtype == torch.float32 else 0
    _lib.check(_lib.lib().s3r_gemm_bf16_batched(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), C.c_void_p(c.data_ptr()),
                                                M, N, K, lda, ldb, ldc, sa, sb, sc, batch, flags, int(a_mn), int(b_mn),
                                                C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)),
               "s3r_gemm_bf16_batched")


def attention_backward(q, k, v, o, do, scale: float):
    """q / o / do [B,Nq,H,64], k / v [B,Nk,H,64] bf16 (any strides) -> (dq, dk, dv) with the shapes of q, k, v."""
    if q.device.type != "cuda" or q.dtype != torch.bfloat16 or q.shape[-1] != 64:
        raise _lib.S3RError("attention_backward expects bf16 CUDA tensors with head_dim 64")
    B, Nq, H, D = q.shape
    Nk = k.shape[1]
    BH = B * H
    heads = lambda t: t.to(torch.bfloat16).permute(0, 2, 1, 3).contiguous().view(BH, t.shape[1], D)
    Q, K_, V, O, dO = heads(q), heads(k), heads(v), heads(o), heads(do)
    dev = q.device
    ldp = (Nk + 7) // 8 * 8
    L = _lib.lib()
    st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
    p = lambda t: C.c_void_p(t.data_ptr())
    S = torch.empty(BH, Nq, ldp, dtype=torch.float32, device=dev)
    _bgemm(Q, K_, S, Nq, Nk, D, D, D, ldp, Nq * D, Nk * D, Nq * ldp, BH, False, False)
    P = torch.empty(BH, Nq, ldp, dtype=torch.bfloat16, device=dev)
    _lib.check(L.s3r_softmax_rows_bf16(p(S), p(P), BH * Nq, Nk, ldp, ldp, float(scale), st), "s3r_softmax_rows_bf16")
    _bgemm(dO, V, S, Nq, Nk, D, D, D, ldp, Nq * D, Nk * D, Nq * ldp, BH, False, False)      # S now holds dP
    dS = torch.empty(BH, Nq, ldp, dtype=torch.bfloat16, device=dev)
    _lib.check(L.s3r_attention_ds_bf16(p(P), p(S), p(O), p(dO), p(dS), BH * Nq, Nk, ldp, ldp, float(scale), st),
               "s3r_attention_ds_bf16")
    dV = torch.empty(BH, Nk, D, dtype=torch.bfloat16, device=dev)
    dK = torch.empty(BH, Nk, D, dtype=torch.bfloat16, device=dev)
    dQ = torch.empty(BH, Nq, D, dtype=torch.bfloat16, device=dev)
    _bgemm(P, dO, dV, Nk, D, Nq, ldp, D, D, Nq * ldp, Nq * D, Nk * D, BH, True, True)
    _bgemm(dS, Q, dK, Nk, D, Nq, ldp, D, D, Nq * ldp, Nq * D, Nk * D, BH, True, True)
    _bgemm(dS, K_, dQ, Nq, D, Nk, ldp, D, D, Nq * ldp, Nk * D, Nq * D, BH, False, True)
    back = lambda t, n: t.view(B, H, n, D).permute(0, 2, 1, 3)
    return back(dQ, Nq), back(dK, Nk), back(dV, Nk)
