"""Transformer pieces of the CroCo / MASt3R ViT used by Styl3R, with the reference's parameter registry
 so that checkpoints load strictly:

  Block           norm1, attn.{qkv, proj}, norm2, mlp.{fc1, fc2}                 croco/blocks.py:136-152
  DecoderBlock    + cross_attn.{projq, projk, projv, proj}, norm3, norm_y          croco/blocks.py:202-222
  PatchEmbed      proj (Conv2d k16 s16), integer (y, x) positions                  croco/patch_embed.py:19-29

Differences from the reference implementation (results are the same): RoPE is applied in place on the q / k
slices of the packed qkv tensor by our CUDA kernel (no transposes / copies), attention goes through
styl3r_b200.ops.memory_efficient_attention, and the context K/V projections are computed once per layer.
"""
from __future__ import annotations

import torch
from torch import Tensor, nn

from .. import gemm as _gemm
from ..curope import cuRoPE2D_func
from .. import ops as _ops
from ..ops import memory_efficient_attention
from ..streams import fork_join
from . import train_ops as _tops

LN_EPS = 1e-6  # croco.py:34

# bf16 TRAINING layout (`EncoderNoPoSplatMultiTokenStyle.to_training()`): with autograd enabled the ViT trunks run the
# autograd functions of train_ops.py - forward AND backward on the tcgen05 GEMM (dgrad / wgrad with MN-major operands),
# the tcgen05 attention kernel + batched-GEMM attention backward and our LayerNorm kernels; parameters stay fp32.
# False (default): autograd runs the reference's fp32 / TF32 torch ops (the numerics the golden tests pin).
TRAIN_BF16 = False
TRAIN_BACKWARD_KIND = "torch autograd over the reference's fp32/TF32 ops (cuBLAS / cuDNN / SDPA)"


def _train(x: Tensor) -> bool:
    return TRAIN_BF16 and torch.is_grad_enabled() and x.is_cuda


def _bf16(x: Tensor) -> Tensor:
    return x if x.dtype == torch.bfloat16 else x.to(torch.bfloat16)


def _rope(t_bnhd: Tensor, pos: Tensor, base: float) -> Tensor:
    """In-place RoPE-2D on a [B,N,H,D] (possibly strided) tensor."""
    return cuRoPE2D_func.apply(t_bnhd, pos if pos.is_contiguous() else pos.contiguous(), base, 1.0)


def _fast(layer: nn.Linear, x: Tensor) -> bool:
    """bf16 inference layout (to_inference) without autograd: the tcgen05 kernels are used."""
    return (x.dtype == torch.bfloat16 and layer.weight.dtype == torch.bfloat16 and not torch.is_grad_enabled()
            and not _ops.FORCE_LIBRARY)


# Optional fp32 residual stream for the inference layout (patch embedding -> blocks -> decoder blocks): the proj / fc2
# GEMMs add an fp32 residual and write fp32, LayerNorm reads fp32 and emits the bf16 GEMM operand, so that only GEMM /
# attention OPERANDS are rounded to bf16 instead of the stream itself.  Measured on B200 (scripts/parity_probe.py,
# cfg2): error vs the fp32 golden drops only from 0.0074 to 0.0067 sigma (means, mean) / 0.038 to 0.029 (covariances) -
# the operand rounding dominates, not its accumulation in the stream - while the encoder slows from 7.1 to 8.0 ms
# (fp32 stream traffic on the critical chain).  Off by default.
STREAM_FP32 = False


def _lin(layer: nn.Linear, x: Tensor, residual: Tensor | None = None, gelu: bool = False, stream: bool = False) -> Tensor:
    """y = act(layer(x)) + residual.  bf16 inference layout (to_inference): one tcgen05 GEMM with the bias, exact
    GELU and residual add fused in its epilogue; otherwise (fp32 / training) the torch ops of the reference.
    `stream`: the result starts a residual stream (fp32 in the inference layout when STREAM_FP32)."""
    if _fast(layer, x):
        f32 = STREAM_FP32 and ((residual is not None and residual.dtype == torch.float32) or stream)
        return _gemm.linear(x, layer.weight, layer.bias, residual=residual, gelu=gelu,
                            out_dtype=torch.float32 if f32 else torch.bfloat16)
    if _train(x) and not gelu and layer.in_features % 8 == 0 and layer.out_features % 8 == 0:
        return _tops.linear(_bf16(x), layer, residual=None if residual is None else _bf16(residual))
    y = layer(x)
    if gelu:
        y = torch.nn.functional.gelu(y)
    return y if residual is None else residual + y


def _ln(norm: nn.LayerNorm, x: Tensor) -> Tensor:
    """LayerNorm: bf16 inference layout -> `s3r_layernorm_bf16` (one warp per row, fp32 statistics, launched with
    programmatic dependent launch); otherwise the torch module."""
    C_ = x.shape[-1]
    if _train(x) and C_ % 256 == 0 and C_ <= 1024:
        return _tops.layer_norm(_bf16(x), norm)
    if (not _ops.FORCE_LIBRARY and x.dtype in (torch.bfloat16, torch.float32) and norm.weight.dtype == torch.bfloat16
            and not torch.is_grad_enabled() and x.is_cuda
            and C_ % 256 == 0 and C_ <= 2048):
        import ctypes as C
        from .. import _lib
        x2 = x.reshape(-1, C_)
        if x2.stride(1) != 1 or x2.stride(0) % 8:
            x2 = x2.contiguous()
        y = torch.empty((x2.shape[0], C_), dtype=torch.bfloat16, device=x.device)
        fn = _lib.lib().s3r_layernorm_f32_bf16 if x.dtype == torch.float32 else _lib.lib().s3r_layernorm_bf16
        _lib.check(fn(C.c_void_p(x2.data_ptr()), C.c_void_p(norm.weight.data_ptr()),
                                                 C.c_void_p(norm.bias.data_ptr()), C.c_void_p(y.data_ptr()), x2.shape[0], C_,
                                                 x2.stride(0), float(norm.eps),
                                                 C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
                   "s3r_layernorm_bf16")
        return y.view(*x.shape)
    return norm(x)


class Mlp(nn.Module):
    def __init__(self, dim: int, hidden: int):
        super().__init__()
        self.fc1 = nn.Linear(dim, hidden)
        self.act = nn.GELU()
        self.fc2 = nn.Linear(hidden, dim)

    def forward(self, x: Tensor, residual: Tensor | None = None) -> Tensor:
        if _train(x):
            return _tops.mlp(_bf16(x), self.fc1, self.fc2, residual=None if residual is None else _bf16(residual))
        return _lin(self.fc2, _lin(self.fc1, x, gelu=True), residual=residual)


class Attention(nn.Module):
    def __init__(self, dim: int, num_heads: int, rope_base: float):
        super().__init__()
        self.num_heads, self.scale, self.rope_base = num_heads, (dim // num_heads) ** -0.5, rope_base
        self.qkv = nn.Linear(dim, dim * 3, bias=True)
        self.proj = nn.Linear(dim, dim)

    def forward(self, x: Tensor, xpos: Tensor, residual: Tensor | None = None) -> Tensor:
        B, N, C = x.shape
        if _train(x) and C // self.num_heads == 64:
            qkv = _tops.linear(_bf16(x), self.qkv, rope_pos=xpos, rope_cols=2 * C, rope_base=self.rope_base)
            qkv = qkv.view(B, N, 3, self.num_heads, C // self.num_heads)
            o = _tops.attention(qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2], self.scale)
            return _lin(self.proj, o.reshape(B, N, C), residual=residual)
        if _fast(self.qkv, x) and C // self.num_heads == 64:
            # RoPE on the q and k thirds happens in the GEMM epilogue (fp32, before the bf16 rounding)
            qkv = _gemm.linear(x, self.qkv.weight, self.qkv.bias, rope_pos=xpos, rope_cols=2 * C, rope_base=self.rope_base)
            qkv = qkv.view(B, N, 3, self.num_heads, C // self.num_heads)
            q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
        else:
            qkv = _lin(self.qkv, x).view(B, N, 3, self.num_heads, C // self.num_heads)
            q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]      # [B,N,H,D] views, stride(2) == D
            q, k = _rope(q, xpos, self.rope_base), _rope(k, xpos, self.rope_base)
        o = memory_efficient_attention(q, k, v, scale=self.scale)
        return _lin(self.proj, o.reshape(B, N, C), residual=residual)


class CrossAttention(nn.Module):
    def __init__(self, dim: int, num_heads: int, rope_base: float):
        super().__init__()
        self.num_heads, self.scale, self.rope_base = num_heads, (dim // num_heads) ** -0.5, rope_base
        self.projq = nn.Linear(dim, dim, bias=True)
        self.projk = nn.Linear(dim, dim, bias=True)
        self.projv = nn.Linear(dim, dim, bias=True)
        self.proj = nn.Linear(dim, dim)

    def project_q(self, query: Tensor, qpos: Tensor) -> Tensor:
        B, Nq, C = query.shape
        H, D = self.num_heads, C // self.num_heads
        if _train(query) and D == 64:
            return _tops.linear(_bf16(query), self.projq, rope_pos=qpos, rope_cols=C, rope_base=self.rope_base).view(B, Nq, H, D)
        if _fast(self.projq, query) and D == 64:
            return _gemm.linear(query, self.projq.weight, self.projq.bias, rope_pos=qpos, rope_cols=C,
                                rope_base=self.rope_base).view(B, Nq, H, D)
        return _rope(_lin(self.projq, query).view(B, Nq, H, D), qpos, self.rope_base)

    def project_kv(self, key: Tensor, value: Tensor, kpos: Tensor):
        """K (with RoPE) and V of the context tokens - independent of the query stream, so a DecoderBlock can compute
        them concurrently with its self-attention."""
        B, Nk, C = key.shape
        H, D = self.num_heads, C // self.num_heads
        if _train(key) and D == 64:
            k = _tops.linear(_bf16(key), self.projk, rope_pos=kpos, rope_cols=C, rope_base=self.rope_base).view(B, Nk, H, D)
            return k, _lin(self.projv, value).view(B, value.shape[1], H, D)
        if _fast(self.projk, key) and D == 64 and key is value:
            # one GEMM for both projections (N = 2C, RoPE on the K half only): the batch-1 forward is bound by the number
            # of kernel launches the graph has to dispatch, not by their FLOPs
            w, b_ = self._kv_operands()
            kv = _gemm.linear(key, w, b_, rope_pos=kpos, rope_cols=C, rope_base=self.rope_base).view(B, Nk, 2, H, D)
            return kv[:, :, 0], kv[:, :, 1]
        if _fast(self.projk, key) and D == 64:
            k = _gemm.linear(key, self.projk.weight, self.projk.bias, rope_pos=kpos, rope_cols=C,
                             rope_base=self.rope_base).view(B, Nk, H, D)
        else:
            k = _rope(_lin(self.projk, key).view(B, Nk, H, D), kpos, self.rope_base)
        return k, _lin(self.projv, value).view(B, value.shape[1], H, D)

    def _kv_operands(self):
        """[projk; projv] stacked weight / bias for the fused K/V projection, rebuilt when the parameters change."""
        key = (self.projk.weight.data_ptr(), self.projk.weight._version, self.projv.weight.data_ptr(),
               self.projv.weight._version, self.projk.bias._version, self.projv.bias._version)
        cache = getattr(self, "_kv_cache", None)
        if cache is None or cache[0] != key:
            w = torch.cat((self.projk.weight.detach(), self.projv.weight.detach()), dim=0).contiguous()
            b_ = torch.cat((self.projk.bias.detach(), self.projv.bias.detach()), dim=0).contiguous()
            object.__setattr__(self, "_kv_cache", (key, w, b_))
            cache = self._kv_cache
        return cache[1], cache[2]

    def attend(self, q: Tensor, k: Tensor, v: Tensor, residual: Tensor | None = None) -> Tensor:
        B, Nq, H, D = q.shape
        if _train(q) and D == 64:
            o = _tops.attention(q, k, v, self.scale)
        else:
            o = memory_efficient_attention(q, k, v, scale=self.scale)
        return _lin(self.proj, o.reshape(B, Nq, H * D), residual=residual)

    def forward(self, query: Tensor, key: Tensor, value: Tensor, qpos: Tensor, kpos: Tensor,
                residual: Tensor | None = None) -> Tensor:
        q = self.project_q(query, qpos)
        k, v = self.project_kv(key, value, kpos)
        return self.attend(q, k, v, residual=residual)


class Block(nn.Module):
    def __init__(self, dim: int, num_heads: int, rope_base: float, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = Attention(dim, num_heads, rope_base)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))

    def forward(self, x: Tensor, xpos: Tensor) -> Tensor:
        x = self.attn(_ln(self.norm1, x), xpos, residual=x)
        return self.mlp(_ln(self.norm2, x), residual=x)


class DecoderBlock(nn.Module):
    def __init__(self, dim: int, num_heads: int, rope_base: float, mlp_ratio: float = 4.0):
        super().__init__()
        self.norm1 = nn.LayerNorm(dim, eps=LN_EPS)
        self.attn = Attention(dim, num_heads, rope_base)
        self.cross_attn = CrossAttention(dim, num_heads, rope_base)
        self.norm2 = nn.LayerNorm(dim, eps=LN_EPS)
        self.norm3 = nn.LayerNorm(dim, eps=LN_EPS)
        self.mlp = Mlp(dim, int(dim * mlp_ratio))
        self.norm_y = nn.LayerNorm(dim, eps=LN_EPS)

    def forward(self, x: Tensor, y: Tensor, xpos: Tensor, ypos: Tensor, parallel: bool = False) -> Tensor:
        """`parallel`: the context branch (norm_y -> projk/projv) runs concurrently with the self-attention branch
        (streams.fork_join); same kernels and operands either way, so the result is identical."""
        def self_branch():
            x1 = self.attn(_ln(self.norm1, x), xpos, residual=x)
            return x1, self.cross_attn.project_q(_ln(self.norm2, x1), xpos)

        def ctx_branch():
            y_ = _ln(self.norm_y, y)
            return self.cross_attn.project_kv(y_, y_, ypos)

        (x1, q), (k, v) = fork_join([self_branch, ctx_branch], x.device, parallel=parallel)
        x2 = self.cross_attn.attend(q, k, v, residual=x1)
        return self.mlp(_ln(self.norm3, x2), residual=x2)


class PatchEmbed(nn.Module):
    """Conv2d(3 -> dim, k = s = patch) + flatten; positions are the integer (y, x) patch coordinates."""

    def __init__(self, patch_size: int, in_chans: int, dim: int):
        super().__init__()
        self.patch_size = (patch_size, patch_size)
        self.proj = nn.Conv2d(in_chans, dim, kernel_size=patch_size, stride=patch_size)

    def forward(self, img: Tensor):
        B, _, H, W = img.shape
        ph, pw = self.patch_size
        assert H % ph == 0, f"Input image height ({H}) is not a multiple of patch size ({ph})."
        assert W % pw == 0, f"Input image width ({W}) is not a multiple of patch size ({pw})."
        gh, gw = H // ph, W // pw
        if max(gh + 1, gw) > _gemm.ROPE_MAX_POS:  # (+1: the intrinsics token sits at row gh)
            # the fused-RoPE GEMM epilogue looks positions up in a (ROPE_MAX_POS + 1)-entry table and would clamp
            raise ValueError(f"patch grid {gh}x{gw} exceeds the RoPE table ({_gemm.ROPE_MAX_POS} positions per axis): "
                             f"images above {_gemm.ROPE_MAX_POS * ph} px per side are not supported")
        ys, xs = torch.meshgrid(torch.arange(gh, device=img.device), torch.arange(gw, device=img.device), indexing="ij")
        pos = torch.stack((ys.reshape(-1), xs.reshape(-1)), dim=-1)[None].expand(B, -1, -1).contiguous()
        w = self.proj.weight
        if img.dtype == torch.bfloat16 and w.dtype == torch.bfloat16 and not torch.is_grad_enabled() and not _ops.FORCE_LIBRARY:
            # kernel == stride: the convolution is a GEMM over non-overlapping patches, columns ordered (c, kh, kw)
            # like weight.flatten(1) - one gather copy + the tcgen05 GEMM instead of a cuDNN conv with layout transposes
            cols = img.reshape(B, -1, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, -1)
            return _gemm.linear(cols, w.flatten(1), self.proj.bias,
                                out_dtype=torch.float32 if STREAM_FP32 else torch.bfloat16).view(B, gh * gw, -1), pos
        if _train(img):
            cols = _bf16(img).reshape(B, -1, gh, ph, gw, pw).permute(0, 2, 4, 1, 3, 5).reshape(B * gh * gw, -1)
            y = _tops.LinearFn.apply(cols, w.flatten(1), self.proj.bias, None, None, 0, 100.0)
            return y.view(B, gh * gw, -1), pos
        x = self.proj(img)
        return x.flatten(2).transpose(1, 2), pos


class CroCoTrunk(nn.Module):
    """ViT-L/16 encoder (24 x 1024 x 16 heads) + decoder_embed + 12 x 768 x 12-head DecoderBlocks, RoPE base 100
    (`croco_params['ViTLarge_BaseDecoder']`, backbone_croco_multiview.py:21-32; croco.py:21-84).  Holds exactly the
    parameters of the reference's CroCoNet subclasses, including the unused `mask_token`."""

    enc_depth, dec_depth, enc_embed_dim, dec_embed_dim, enc_heads, dec_heads, rope_base = 24, 12, 1024, 768, 16, 12, 100.0

    def __init__(self, second_decoder: bool, intrinsics_token: bool):
        super().__init__()
        E, Dd = self.enc_embed_dim, self.dec_embed_dim
        self.patch_embed = PatchEmbed(16, 3, E)
        self.enc_blocks = nn.ModuleList([Block(E, self.enc_heads, self.rope_base) for _ in range(self.enc_depth)])
        self.enc_norm = nn.LayerNorm(E, eps=LN_EPS)
        self.mask_token = nn.Parameter(torch.zeros(1, 1, Dd))
        self.decoder_embed = nn.Linear(E, Dd, bias=True)
        self.dec_blocks = nn.ModuleList([DecoderBlock(Dd, self.dec_heads, self.rope_base) for _ in range(self.dec_depth)])
        self.dec_norm = nn.LayerNorm(Dd, eps=LN_EPS)
        if second_decoder:
            self.dec_blocks2 = nn.ModuleList([DecoderBlock(Dd, self.dec_heads, self.rope_base) for _ in range(self.dec_depth)])
        if intrinsics_token:
            self.intrinsic_encoder = nn.Linear(9, E)
        self._init_weights()

    def _init_weights(self):
        # croco.py:112-127: xavier-uniform linears, unit LayerNorms, N(0, .02) mask token, xavier patch projection
        w = self.patch_embed.proj.weight.data
        nn.init.xavier_uniform_(w.view(w.shape[0], -1))
        nn.init.normal_(self.mask_token, std=0.02)
        for m in self.modules():
            if isinstance(m, nn.Linear):
                nn.init.xavier_uniform_(m.weight)
                if m.bias is not None:
                    nn.init.constant_(m.bias, 0)
            elif isinstance(m, nn.LayerNorm):
                nn.init.constant_(m.bias, 0)
                nn.init.constant_(m.weight, 1.0)

    def encode(self, img: Tensor, extra_token: Tensor | None = None):
        x, pos = self.patch_embed(img)
        if extra_token is not None:  # intrinsics token appended at position (grid_h, 0)  (…multiview.py:131-135)
            x = torch.cat((x, extra_token.to(x.dtype)), dim=1)
            tok_pos = pos[:, :1].clone()
            tok_pos[:, :, 0] += pos[:, -1:, 0] + 1
            pos = torch.cat((pos, tok_pos), dim=1)
        for blk in self.enc_blocks:
            x = blk(x, pos)
        return _ln(self.enc_norm, x), pos
