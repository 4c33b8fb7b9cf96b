"""bf16 linear layers on the tcgen05 GEMM (styl3r_b200/csrc/gemm_tcgen05.cu) — the dense-contraction path of the
ViT trunks in the B200 inference layout (`EncoderNoPoSplatMultiTokenStyle.to_inference`).

    y = linear(x, weight, bias=None, residual=None, gelu=False)      # y = act(x W^T + b) + residual

replaces `nn.Linear` (+ `nn.GELU`, + the residual add of the transformer block) of
src/model/encoder/backbone/croco/blocks.py:61-82,97-134,149-152 with one kernel."""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from . import _lib

EPI_BIAS, EPI_GELU, EPI_RESIDUAL, EPI_OUT_F32, EPI_ROPE, EPI_RELU = 1, 2, 4, 8, 16, 32
EPI_RES_F32 = 512
ROPE_MAX_POS = 255  # positions are patch-grid coordinates (16 for 256 px, 64 for 1024 px)
_rope_tables: dict = {}
_workspaces: dict = {}
WORKSPACE_BYTES = 16384 + 32 * 1024 * 1024


def _workspace(device) -> torch.Tensor:
    """Split-K scratch, private to (device, current stream): 16 KiB of zeroed tile counters + fp32 partial tiles."""
    key = (str(device), torch.cuda.current_stream(device).cuda_stream)
    w = _workspaces.get(key)
    if w is None:
        w = torch.zeros(WORKSPACE_BYTES, dtype=torch.uint8, device=device)
        _workspaces[key] = w
    return w


def rope_table(device, base: float) -> torch.Tensor:
    """(cos, sin)[pos, d] of pos * base^(-d/16), built once per (device, base) by s3r_rope_table."""
    key = (str(device), float(base))
    t = _rope_tables.get(key)
    if t is None:
        t = torch.empty((ROPE_MAX_POS + 1) * 16 * 2, dtype=torch.float32, device=device)
        _lib.check(_lib.lib().s3r_rope_table(C.c_void_p(t.data_ptr()), ROPE_MAX_POS, float(base),
                                            C.c_void_p(torch.cuda.current_stream(device).cuda_stream)), "s3r_rope_table")
        _rope_tables[key] = t
    return t


def linear(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None,
           residual: Optional[torch.Tensor] = None, gelu: bool = False, out_dtype: torch.dtype = torch.bfloat16,
           rope_pos: Optional[torch.Tensor] = None, rope_cols: int = 0, rope_base: float = 100.0,
           split_k: bool = False, relu: bool = False):
    """rope_pos [..., 2] int64 (one (y, x) per row of x) + rope_cols: RoPE-2D (head_dim 64) is applied to output columns
    [0, rope_cols) inside the GEMM epilogue (q and k parts of a qkv projection), replacing a separate rope pass.
    split_k: let grids smaller than the machine split the K range over several CTAs (deterministic last-CTA fix-up
    through an fp32 workspace).  Off by default: on B200 the extra L2 round trips of the fix-up (partials, fence, ticket)
    cost more than the shorter K loops save for the M = 257/514 shapes measured (scripts/bench_gemm.py)."""
    if x.device.type != "cuda":
        raise _lib.S3RError("styl3r_b200.gemm.linear needs CUDA tensors (no CPU fallback)")
    if x.dtype != torch.bfloat16 or weight.dtype != torch.bfloat16:
        raise _lib.S3RError("styl3r_b200.gemm.linear expects bf16 activations and weights")
    N, K = weight.shape
    lead = x.shape[:-1]
    x2 = x.reshape(-1, K)
    if x2.stride(1) != 1:
        x2 = x2.contiguous()
    w = weight if weight.stride(1) == 1 else weight.contiguous()
    M = x2.shape[0]
    out = torch.empty((M, N), dtype=out_dtype, device=x.device)
    flags = 0
    bp = rp = None
    ldr = 0
    if bias is not None:
        flags |= EPI_BIAS
        bias = bias if bias.dtype == torch.bfloat16 else bias.to(torch.bfloat16)
        bp = C.c_void_p(bias.data_ptr())
    if gelu:
        flags |= EPI_GELU
    if relu:
        flags |= EPI_RELU
    if residual is not None:
        flags |= EPI_RESIDUAL
        r2 = residual.reshape(-1, N)
        if r2.dtype == torch.float32:  # fp32 residual stream (inference layout of the ViT trunks)
            if r2.stride(1) != 1 or r2.stride(0) % 4:
                r2 = r2.contiguous()
            flags |= EPI_RES_F32
        elif r2.stride(1) != 1 or r2.dtype != torch.bfloat16:
            r2 = r2.to(torch.bfloat16).contiguous()
        rp, ldr = C.c_void_p(r2.data_ptr()), r2.stride(0)
    pp = tp = None
    if rope_pos is not None:
        flags |= EPI_ROPE
        rp2 = rope_pos.reshape(-1, 2)
        if rp2.dtype != torch.int64 or not rp2.is_contiguous():
            rp2 = rp2.to(torch.int64).contiguous()
        if rp2.shape[0] != M:
            raise _lib.S3RError("rope_pos must hold one (y, x) per row")
        table = rope_table(x.device, rope_base)
        pp, tp = C.c_void_p(rp2.data_ptr()), C.c_void_p(table.data_ptr())
    if out_dtype == torch.float32:
        flags |= EPI_OUT_F32
    elif out_dtype != torch.bfloat16:
        raise _lib.S3RError("out_dtype must be bf16 or fp32")
    ws = _workspace(x.device) if split_k else None
    _lib.check(_lib.lib().s3r_gemm_bf16_rope(C.c_void_p(x2.data_ptr()), C.c_void_p(w.data_ptr()), bp, rp,
                                             C.c_void_p(out.data_ptr()), M, N, K, x2.stride(0), w.stride(0),
                                             out.stride(0), ldr, flags, pp, tp, int(rope_cols), ROPE_MAX_POS,
                                             None if ws is None else C.c_void_p(ws.data_ptr()),
                                             0 if ws is None else ws.numel(),
                                             C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
               "s3r_gemm_bf16")
    return out.reshape(*lead, N)


EPI_DGELU = 128


def gemm_majors(a: torch.Tensor, b: torch.Tensor, M: int, N: int, K: int, a_mn_major: bool, b_mn_major: bool,
                bias: Optional[torch.Tensor] = None, aux: Optional[torch.Tensor] = None, dgelu: bool = False,
                out_dtype: torch.dtype = torch.bfloat16, gelu: bool = False, pre_out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """C[M,N] = opA . opB^T (+ bias) (+ aux | * gelu'(aux)) with either operand MN-major (`s3r_gemm_bf16_majors`):
    a is [M,K] (a_mn_major False) or [K,M] (True); b is [N,K] or [K,N].  2-D bf16 tensors with unit inner stride."""
    if a.dtype != torch.bfloat16 or b.dtype != torch.bfloat16 or a.device.type != "cuda":
        raise _lib.S3RError("gemm_majors expects bf16 CUDA operands")
    if a.stride(1) != 1:
        a = a.contiguous()
    if b.stride(1) != 1:
        b = b.contiguous()
    out = torch.empty((M, N), dtype=out_dtype, device=a.device)
    flags = 0
    bp = xp = None
    ldx = 0
    if bias is not None:
        flags |= EPI_BIAS
        bias = bias if bias.dtype == torch.bfloat16 else bias.to(torch.bfloat16)
        bp = C.c_void_p(bias.data_ptr())
    if aux is not None:
        flags |= EPI_DGELU if dgelu else EPI_RESIDUAL
        if aux.stride(1) != 1 or aux.dtype != torch.bfloat16:
            aux = aux.to(torch.bfloat16).contiguous()
        xp, ldx = C.c_void_p(aux.data_ptr()), aux.stride(0)
    if out_dtype == torch.float32:
        flags |= EPI_OUT_F32
    if gelu:
        flags |= EPI_GELU
    if pre_out is not None and (pre_out.shape != (M, N) or pre_out.stride() != out.stride() or pre_out.dtype != torch.bfloat16):
        raise _lib.S3RError("pre_out must be a contiguous bf16 [M, N] tensor")
    _lib.check(_lib.lib().s3r_gemm_bf16_majors(C.c_void_p(a.data_ptr()), C.c_void_p(b.data_ptr()), bp, xp,
                                               C.c_void_p(out.data_ptr()), M, N, K, a.stride(0), b.stride(0), out.stride(0),
                                               ldx, flags, int(a_mn_major), int(b_mn_major),
                                               None if pre_out is None else C.c_void_p(pre_out.data_ptr()),
                                               C.c_void_p(torch.cuda.current_stream(a.device).cuda_stream)),
               "s3r_gemm_bf16_majors")
    return out


def linear_dgrad(dy: torch.Tensor, weight: torch.Tensor, pre_gelu: Optional[torch.Tensor] = None) -> torch.Tensor:
    """dX[M,Kin] = dY[M,Nout] . W[Nout,Kin]  (optionally * gelu'(pre_gelu): the layer in front ended in a fused GELU)."""
    M, Nout = dy.shape
    return gemm_majors(dy, weight, M, weight.shape[1], Nout, False, True, aux=pre_gelu, dgelu=pre_gelu is not None)


def linear_wgrad(dy: torch.Tensor, x: torch.Tensor, out_dtype: torch.dtype = torch.float32) -> torch.Tensor:
    """dW[Nout,Kin] = dY[M,Nout]^T . X[M,Kin], fp32 by default (it lands in an fp32 .grad)."""
    M, Nout = dy.shape
    return gemm_majors(dy, x, Nout, x.shape[1], M, True, True, out_dtype=out_dtype)
