"""`EncoderNoPoSplatMultiTokenStyle` — the production Styl3R encoder (backbone `croco_multi` + token stylizer + DPT
heads + unified Gaussian adapter) behind the reference's module surface
(src/model/encoder/encoder_noposplat_multi_token_style.py:46-263, src/model/encoder/__init__.py:20-25):

    encoder, _ = get_encoder(cfg)
    gaussians  = encoder(context, style, global_step=0, visualization_dump=None)   # -> Gaussians

with `context = {"image": [b,v,3,h,w] in [-1,1], "intrinsics": [b,v,3,3] normalised, ...}` and
`style = {"image": [b,3,h,w]}`.  The parameter registry equals the reference's (SURVEY.md Appendix C;
tests/golden/encoder_state_manifest.json), so `load_state_dict(strict=True)` works with existing checkpoints.

B200 inference layout (`to_inference(torch.bfloat16)` + `GraphedEncoder`): every Linear (+ bias / GELU / residual /
RoPE-2D) on the tcgen05 GEMM, attention on the tcgen05 attention kernel, DPT convolutions on the tcgen05 implicit-GEMM
convolution (bf16 NHWC), LayerNorm / bilinear x2 / head-epilogue -> Gaussians as our own kernels, independent branches
on concurrent streams, the whole forward replayed as one CUDA graph.  With autograd enabled (training)
the same modules run the reference's torch ops in fp32.  CUDA only — there is no CPU fallback.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Any, List, Literal, Optional

import torch
from torch import Tensor, nn

from .. import _lib
from ..streams import fork_join
from .dpt import HOOKS, PixelwiseDPT
from .vit import CroCoTrunk, _lin, _ln


@dataclass
class Gaussians:  # src/model/types.py:7-12
    means: Tensor         # [batch, gaussian, 3]
    covariances: Tensor   # [batch, gaussian, 3, 3]
    harmonics: Tensor     # [batch, gaussian, 3, d_sh]
    opacities: Tensor     # [batch, gaussian]


@dataclass
class OpacityMappingCfg:
    initial: float = 0.0
    final: float = 0.0
    warm_up: int = 1


@dataclass
class GaussianAdapterCfg:
    gaussian_scale_min: float = 0.5
    gaussian_scale_max: float = 15.0
    sh_degree: int = 0


@dataclass
class BackboneCrocoCfg:
    name: Literal["croco", "croco_multi"] = "croco_multi"
    model: str = "ViTLarge_BaseDecoder"
    patch_embed_cls: str = "PatchEmbedDust3R"
    asymmetry_decoder: bool = True
    intrinsics_embed_loc: str = "encoder"
    intrinsics_embed_degree: int = 4
    intrinsics_embed_type: str = "token"


@dataclass
class TokenStylizerCfg:
    model: str = "ViTLarge_BaseDecoder"
    patch_embed_cls: str = "PatchEmbedDust3R"
    pretrained_weights: str = ""


@dataclass
class EncoderNoPoSplatTokenStyleCfg:
    name: str = "noposplat_multi_token_style"
    d_feature: int = 128
    num_monocular_samples: int = 32
    backbone: BackboneCrocoCfg = field(default_factory=BackboneCrocoCfg)
    token_stylizer: TokenStylizerCfg = field(default_factory=TokenStylizerCfg)
    structure_builder: Any = None
    visualizer: Any = None
    gaussian_adapter: GaussianAdapterCfg = field(default_factory=GaussianAdapterCfg)
    apply_bounds_shim: bool = True
    opacity_mapping: OpacityMappingCfg = field(default_factory=OpacityMappingCfg)
    gaussians_per_pixel: int = 1
    num_surfaces: int = 1
    gs_params_head_type: str = "dpt_gs"
    gs_sh_head_type: str = "dpt"
    input_mean: tuple = (0.5, 0.5, 0.5)
    input_std: tuple = (0.5, 0.5, 0.5)
    pretrained_weights: str = ""
    pose_free: bool = True
    stylized: bool = False


def _feature(t: Tensor, dtype: torch.dtype) -> Tensor:
    """Feature maps handed to the DPT heads carry the trunk's operand dtype (the fp32 residual stream of the inference
    layout is rounded once here)."""
    return t if t.dtype == dtype else t.to(dtype)


class AsymmetricCroCoMulti(CroCoTrunk):
    """Backbone: shared ViT-L encoder over all context views (+ intrinsics token) and two cross-view decoders —
    `dec_blocks` for view 0, `dec_blocks2` for views 1..v-1 (backbone_croco_multiview.py:51-227)."""

    def __init__(self, cfg: BackboneCrocoCfg):
        if cfg.model != "ViTLarge_BaseDecoder" or cfg.intrinsics_embed_loc != "encoder" or cfg.intrinsics_embed_type != "token":
            raise NotImplementedError("production configuration only: ViTLarge_BaseDecoder + intrinsics token in the encoder")
        super().__init__(second_decoder=cfg.asymmetry_decoder, intrinsics_token=True)
        self.asymmetric = cfg.asymmetry_decoder

    def load_state_dict(self, ckpt, **kw):
        """Like the reference (…multiview.py:94-101): duplicate `dec_blocks` into `dec_blocks2` when absent."""
        ckpt = dict(ckpt)
        if self.asymmetric and not any(k.startswith("dec_blocks2") for k in ckpt):
            for k, v in list(ckpt.items()):
                if k.startswith("dec_blocks"):
                    ckpt[k.replace("dec_blocks", "dec_blocks2")] = v
        return super().load_state_dict(ckpt, **kw)

    _others_idx: dict = {}

    @classmethod
    def _others(cls, x: Tensor) -> Tensor:
        """[b,v,l,c] -> [b,v,(v-1)*l,c]: for every view the tokens of all *other* views, in view order."""
        b, v, l, c = x.shape
        key = (v, x.device)
        if key not in cls._others_idx:  # built once per (v, device): no host->device copy inside a CUDA-graph capture
            cls._others_idx[key] = torch.tensor([[j for j in range(v) if j != i] for i in range(v)], device=x.device)
        return x[:, cls._others_idx[key]].reshape(b, v, (v - 1) * l, c)

    def encode_views(self, context: dict):
        """Shared ViT-L encoder over the b*v context images (+ intrinsics token): (feat [b,v,l,1024], pos [b,v,l,2])."""
        img = context["image"]
        b, v = img.shape[:2]
        tok = self.intrinsic_encoder(context["intrinsics"].flatten(2)).reshape(b * v, 1, -1)
        feat, pos = self.encode(img.reshape(b * v, *img.shape[2:]), tok)
        return feat.reshape(b, v, *feat.shape[1:]), pos.reshape(b, v, *pos.shape[1:])

    def decode_views(self, feat: Tensor, pos: Tensor, parallel: bool = False, hooked_only: bool = False):
        """The two cross-view decoders; `parallel`: view 0 (`dec_blocks`) and views >= 1 (`dec_blocks2`) of a layer
        are independent given the previous layer's output and run as concurrent branches (streams.fork_join).
        `hooked_only` (bf16 inference layout): only the layers the DPT heads hook are assembled into [b, v, l, c]
        tensors (the other list entries are None); the two branches keep their own contiguous tensors from layer to
        layer and the context of views >= 1 is gathered by one kernel (`s3r_gather_other_views_bf16`) instead of
        cat + index + two reshape copies per layer."""
        b, v = feat.shape[:2]
        outs = [feat]
        cur = _lin(self.decoder_embed, feat, stream=True)
        pos_ctx = self._others(pos)
        blocks2 = self.dec_blocks2 if self.asymmetric else self.dec_blocks
        if (hooked_only and v >= 2 and cur.is_cuda and cur.dtype == torch.bfloat16 and not torch.is_grad_enabled()
                and cur.shape[-1] % 8 == 0):
            import ctypes as C
            l, c = cur.shape[2:]
            L = _lib.lib()
            cur0, cur1 = cur[:, 0].contiguous(), cur[:, 1:].contiguous()           # [b, l, c], [b, v-1, l, c]
            pos0, pctx0 = pos[:, 0].contiguous(), pos_ctx[:, 0].contiguous()   # (a strided slice is re-copied by every RoPE call)
            pos1 = pos[:, 1:].reshape(b * (v - 1), *pos.shape[2:])
            pctx1 = pos_ctx[:, 1:].reshape(b * (v - 1), *pos_ctx.shape[2:])
            n_layers = len(self.dec_blocks)
            for li, (blk1, blk2) in enumerate(zip(self.dec_blocks, blocks2)):
                ctx0 = cur1.view(b, (v - 1) * l, c)
                ctx1 = torch.empty(b * (v - 1), (v - 1) * l, c, dtype=cur.dtype, device=cur.device)
                _lib.check(L.s3r_gather_other_views_bf16(C.c_void_p(cur0.data_ptr()), C.c_void_p(cur1.data_ptr()),
                                                         C.c_void_p(ctx1.data_ptr()), b, v, l, c,
                                                         C.c_void_p(torch.cuda.current_stream(cur.device).cuda_stream)),
                           "s3r_gather_other_views_bf16")
                x1 = cur1.view(b * (v - 1), l, c)
                r0, r1 = fork_join([lambda: blk1(cur0, ctx0, pos0, pctx0, parallel=parallel),
                                    lambda: blk2(x1, ctx1, pos1, pctx1, parallel=parallel)], feat.device, parallel=parallel)
                cur0, cur1 = r0, r1.view(b, v - 1, l, c)
                last = li + 1 == n_layers
                outs.append(torch.cat((cur0[:, None], cur1), dim=1) if ((li + 1) in HOOKS or last) else None)
            outs[-1] = _ln(self.dec_norm, outs[-1])
            return [None if o is None else (_feature(o[:, :, :-1], feat.dtype) if i in HOOKS else o[:, :, :-1])
                    for i, o in enumerate(outs)]
        # loop-invariant position slices (a [b, v-1, ...] slice of a batched tensor is a copy: once, not once per layer)
        pos0, pctx0 = pos[:, 0], pos_ctx[:, 0]
        pos1 = pos[:, 1:].reshape(b * (v - 1), *pos.shape[2:]) if v > 1 else None
        pctx1 = pos_ctx[:, 1:].reshape(b * (v - 1), *pos_ctx.shape[2:]) if v > 1 else None
        for blk1, blk2 in zip(self.dec_blocks, blocks2):
            ctx = self._others(cur)
            branches = [lambda: blk1(cur[:, 0], ctx[:, 0], pos0, pctx0, parallel=parallel)]
            if v > 1:
                branches.append(lambda: blk2(
                    cur[:, 1:].reshape(b * (v - 1), *cur.shape[2:]), ctx[:, 1:].reshape(b * (v - 1), *ctx.shape[2:]),
                    pos1, pctx1, parallel=parallel))
            res = fork_join(branches, feat.device, parallel=parallel)
            parts = [res[0][:, None]]
            if v > 1:
                parts.append(res[1].reshape(b, v - 1, *res[1].shape[1:]))
            cur = torch.cat(parts, dim=1)
            outs.append(cur)
        outs[-1] = _ln(self.dec_norm, outs[-1])
        # drop the intrinsics token; only the layers the DPT heads hook are converted
        return [_feature(o[:, :, :-1], feat.dtype) if i in HOOKS else o[:, :, :-1] for i, o in enumerate(outs)]

    def forward(self, context: dict):
        img = context["image"]
        h, w = img.shape[-2:]
        feat, pos = self.encode_views(context)
        return feat, pos, self.decode_views(feat, pos), (h, w), img


class TokenStylizer(CroCoTrunk):
    """Second ViT-L on the style image; 12 DecoderBlocks where all content tokens of a scene self-attend jointly and
    cross-attend to the 256 style tokens (token_stylizer/token_stylizer.py:36-154)."""

    def __init__(self, cfg: TokenStylizerCfg):
        super().__init__(second_decoder=False, intrinsics_token=False)

    def encode_style(self, style: dict):
        """Style ViT-L + decoder_embed: (y [b,256,768], spos [b,256,2]) - independent of the content branch."""
        sfeat, spos = self.encode(style["image"])
        return _lin(self.decoder_embed, sfeat), spos

    def forward(self, style: dict, content_feat: Tensor, content_pos: Tensor) -> List[Tensor]:
        return self.decode(*self.encode_style(style), content_feat, content_pos)

    def decode(self, y: Tensor, spos: Tensor, content_feat: Tensor, content_pos: Tensor,
               parallel: bool = False) -> List[Tensor]:
        b, v, l, _ = content_feat.shape
        outs = [content_feat]
        x = _lin(self.decoder_embed, content_feat.reshape(b, v * l, -1), stream=True)
        xpos = content_pos.reshape(b, v * l, 2)
        for blk in self.dec_blocks:
            x = blk(x, y, xpos, spos, parallel=parallel)
            outs.append(x.reshape(b, v, l, -1))
        outs[-1] = _ln(self.dec_norm, x).reshape(b, v, l, -1)
        return [_feature(o[:, :, :-1], content_feat.dtype) if i in HOOKS else o[:, :, :-1] for i, o in enumerate(outs)]


class UnifiedGaussianAdapter(nn.Module):
    """Parameter-free; `sh_mask` is a non-persistent buffer as in the reference (gaussian_adapter.py:35-48)."""

    def __init__(self, cfg: GaussianAdapterCfg):
        super().__init__()
        self.cfg = cfg
        mask = torch.ones((self.d_sh,), dtype=torch.float32)
        for degree in range(1, cfg.sh_degree + 1):
            mask[degree ** 2:(degree + 1) ** 2] = 0.1 * 0.25 ** degree
        self.register_buffer("sh_mask", mask, persistent=False)

    @property
    def d_sh(self) -> int:
        return (self.cfg.sh_degree + 1) ** 2

    @property
    def d_in(self) -> int:
        return 7 + 3 * self.d_sh


class EncoderNoPoSplatMultiTokenStyle(nn.Module):
    def __init__(self, cfg: EncoderNoPoSplatTokenStyleCfg):
        super().__init__()
        if not cfg.pose_free or cfg.gs_params_head_type != "dpt_gs" or cfg.num_surfaces != 1:
            raise NotImplementedError("production configuration only: pose_free, dpt_gs heads, 1 surface")
        self.cfg = cfg
        self.backbone = AsymmetricCroCoMulti(cfg.backbone)
        self.gaussian_adapter = UnifiedGaussianAdapter(cfg.gaussian_adapter)
        self.pose_free = True
        self.patch_size = self.backbone.patch_embed.patch_size[0]
        d_sh = self.gaussian_adapter.d_sh
        self.raw_gs_dim = 1 + self.gaussian_adapter.d_in
        self.gs_params_head_type = cfg.gs_params_head_type
        self.downstream_head1 = PixelwiseDPT("pts3d", 3)
        self.downstream_head2 = PixelwiseDPT("pts3d", 3)
        self.gaussian_param_head = PixelwiseDPT("gs_params", self.raw_gs_dim - 3 * d_sh)
        self.gaussian_param_head2 = PixelwiseDPT("gs_params", self.raw_gs_dim - 3 * d_sh)
        self.stylized = cfg.stylized
        self.token_stylizer = TokenStylizer(cfg.token_stylizer)
        self.gaussian_appearance_head = PixelwiseDPT("gs_sh", 3 * d_sh)

    def to_inference(self, vit_dtype: torch.dtype = torch.bfloat16, heads: str = "tcgen05", branches: bool = True):
        """Inference layout for B200: ViT trunks (backbone, token stylizer) hold `vit_dtype` weights (bf16 operands,
        fp32 accumulation in the GEMMs / attention; no per-call autocast weight casts).  DPT heads:
        `heads="tcgen05"` (default with bf16 trunks) runs the whole pyramid in bf16 NHWC on the implicit-GEMM
        convolution kernel (fp32 accumulation; `dpt.dpt_forward_nhwc`), `heads="cudnn"` keeps them fp32 (TF32
        convolutions — the reference disables autocast there, encoder…style.py:150) in channels_last so cuDNN runs
        NHWC kernels without layout transposes.  `branches`: independent sub-graphs (content ViT | style ViT, backbone
        decoder | stylizer decoder, dec_blocks | dec_blocks2, the 5 DPT pyramids of every view) run as concurrent
        stream branches (streams.fork_join; graph edges under GraphedEncoder) - at batch 1 the forward is bound by the
        length of its kernel chain, not by FLOPs.  Parameters stay registered as they are: checkpoints still load
        strictly (load_state_dict casts)."""
        if heads not in ("tcgen05", "cudnn"):
            raise ValueError(heads)
        self.backbone.to(vit_dtype)
        self.token_stylizer.to(vit_dtype)
        self._nhwc_heads = heads == "tcgen05" and vit_dtype == torch.bfloat16
        self._branches = bool(branches)
        for head in (self.downstream_head1, self.downstream_head2, self.gaussian_param_head, self.gaussian_param_head2,
                     self.gaussian_appearance_head):
            head.to(memory_format=torch.channels_last)
            head.dpt._prep = None  # bf16 operands are rebuilt from the (possibly reloaded) parameters on first use
        return self

    def to_training(self, vit_dtype: torch.dtype = torch.bfloat16):
        """bf16 TRAINING layout (BASELINE cfg5): fp32 parameters (the optimiser's master copy), bf16 activations in the
        ViT trunks, forward and backward of every trunk Linear / LayerNorm / attention on our kernels through the
        autograd functions of `train_ops.py` (tcgen05 GEMM dgrad / wgrad with MN-major operands, batched-GEMM attention
        backward, LayerNorm backward).  The DPT heads and the adapter keep the reference's fp32 torch ops under autograd
        (convolution wgrad is not written yet - DESIGN.md §7).  `to_training(None)` restores the fp32 / TF32 torch path
        that the golden tests pin.  Process-wide switch (vit.TRAIN_BF16)."""
        from . import vit
        if vit_dtype not in (None, torch.bfloat16):
            raise ValueError("to_training supports torch.bfloat16 or None")
        vit.TRAIN_BF16 = vit_dtype is not None
        vit.TRAIN_BACKWARD_KIND = (
            "ours: tcgen05 GEMM dgrad / wgrad (MN-major operands) + gelu' epilogue, tcgen05 attention forward + batched-GEMM "
            "attention backward, LayerNorm forward / backward kernels for the ViT trunks; DPT heads on cuDNN autograd"
            if vit.TRAIN_BF16 else "torch autograd over the reference's fp32/TF32 ops (cuBLAS / cuDNN / SDPA)")
        return self

    def opacity_exponent(self, global_step: int) -> float:
        m = self.cfg.opacity_mapping
        return 2.0 ** (m.initial + min(global_step / m.warm_up, 1) * (m.final - m.initial))

    def forward(self, context: dict, style: dict, global_step: int = 0,
                visualization_dump: Optional[dict] = None) -> Gaussians:
        img = context["image"]
        if img.device.type != "cuda":
            raise _lib.S3RError("styl3r_b200 encoder needs CUDA tensors (no CPU fallback)")
        b, v, _, h, w = img.shape
        if w < h:
            raise NotImplementedError("portrait inputs: transpose to landscape first (reference transpose_to_landscape)")
        vit_dtype = self.backbone.patch_embed.proj.weight.dtype
        if img.dtype != vit_dtype:  # to_inference(): bf16 trunks
            context = {**context, "image": img.to(vit_dtype), "intrinsics": context["intrinsics"].to(vit_dtype)}
            style = {**style, "image": style["image"].to(vit_dtype)}
        par = getattr(self, "_branches", False) and not torch.is_grad_enabled()
        shape = (h, w)
        (enc_feat, enc_pos), (sty_y, sty_pos) = fork_join(
            [lambda: self.backbone.encode_views(context), lambda: self.token_stylizer.encode_style(style)], img.device,
            parallel=par)
        from .dpt import nhwc_supported
        nhwc = getattr(self, "_nhwc_heads", False) and not torch.is_grad_enabled() and nhwc_supported(shape)
        dec_feat, sty_feat = fork_join(
            [lambda: self.backbone.decode_views(enc_feat, enc_pos, parallel=par, hooked_only=nhwc),
             lambda: self.token_stylizer.decode(sty_y, sty_pos, enc_feat, enc_pos, parallel=par)], img.device, parallel=par)
        HW, G, d_sh = h * w, v * h * w, self.gaussian_adapter.d_sh
        dev = img.device
        means = torch.empty(b, G, 3, device=dev)
        cov = torch.empty(b, G, 3, 3, device=dev)
        harm = torch.empty(b, G, 3, d_sh, device=dev)
        opac = torch.empty(b, G, device=dev)
        scales = torch.empty(b, G, 3, device=dev) if visualization_dump is not None else None
        rots = torch.empty(b, G, 4, device=dev) if visualization_dump is not None else None
        L = _lib.lib()
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())

        def head_branch(head, feats, i, with_img):
            """One DPT pyramid of view i: bf16 NHWC on the tcgen05 implicit-GEMM convolutions (pixel-major fp32 rows
            out), or the fp32 module (planar NCHW out)."""
            def run():
                with torch.autocast("cuda", enabled=False):
                    if nhwc:
                        return head.forward_nhwc([t[:, i] for t in feats], shape, img[:, i, :3] if with_img else None)
                    return head([t[:, i].float() for t in feats], shape,
                                img[:, i, :3].float() if with_img else None).contiguous()
            return run

        if nhwc:
            # views >= 1 share `downstream_head2` / `gaussian_param_head2` and every view shares the appearance head
            # (encoder...style.py:154-176 runs them view by view): batch those views (view-major) into one pyramid each -
            # fewer, larger kernels; the graph's node count, not FLOPs, bounds the batch-1 forward
            from .dpt import HOOKS

            def batched(head, feats, views, with_img):
                def run():
                    with torch.autocast("cuda", enabled=False):
                        toks = [None] * len(feats)
                        for hk in HOOKS:
                            t = feats[hk]
                            toks[hk] = t[:, views[0]] if len(views) == 1 else torch.cat([t[:, i] for i in views], dim=0)
                        im = None
                        if with_img:
                            im = img[:, views[0], :3] if len(views) == 1 else torch.cat([img[:, i, :3] for i in views], dim=0)
                        return head.forward_nhwc(toks, shape, im)
                return run

            rest = list(range(1, v))
            branches = [batched(self.gaussian_appearance_head, sty_feat, list(range(v)), False),
                        batched(self.downstream_head1, dec_feat, [0], False),
                        batched(self.gaussian_param_head, dec_feat, [0], True)]
            if rest:
                branches += [batched(self.downstream_head2, dec_feat, rest, False),
                             batched(self.gaussian_param_head2, dec_feat, rest, True)]
            res = fork_join(branches, dev, parallel=par)
            rows = b * HW
            raw = []
            for i in range(v):
                pts_i = res[1] if i == 0 else res[3][(i - 1) * rows:i * rows]
                prm_i = res[2] if i == 0 else res[4][(i - 1) * rows:i * rows]
                raw += [pts_i, prm_i, res[0][i * rows:(i + 1) * rows]]
        else:
            branches = []
            for i in range(v):
                branches += [head_branch(self.downstream_head1 if i == 0 else self.downstream_head2, dec_feat, i, False),
                             head_branch(self.gaussian_param_head if i == 0 else self.gaussian_param_head2, dec_feat, i, True),
                             head_branch(self.gaussian_appearance_head, sty_feat, i, False)]
            raw = fork_join(branches, dev, parallel=par)
        if torch.is_grad_enabled() and any(t.requires_grad for t in raw):
            # training: differentiable restatement of the adapter with torch ops (the fused kernel has no backward)
            g = self._adapter_autograd(raw, b, v, h, w, global_step)
            if visualization_dump is not None:
                visualization_dump.update(depth=g[0][..., 2].reshape(b, v, h, w, 1, 1), scales=g[4], rotations=g[5],
                                          means=g[0].reshape(b, v, h, w, 1, 3), opacities=g[3].reshape(b, v, h, w, 1, 1))
            return Gaussians(g[0], g[1], g[2], g[3])
        st = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        expo = float(self.opacity_exponent(global_step))
        for i in range(v):
            pts_raw, prm, app = raw[3 * i:3 * i + 3]
            if nhwc:
                _lib.check(L.s3r_gaussian_adapter_nhwc(p(pts_raw), p(prm), p(app), pts_raw.shape[1], prm.shape[1],
                                                       app.shape[1], p(self.gaussian_adapter.sh_mask), b, HW, d_sh, i, G,
                                                       expo, p(means), p(cov), p(harm), p(opac), p(scales), p(rots), st),
                           "s3r_gaussian_adapter_nhwc")
            else:
                _lib.check(L.s3r_gaussian_adapter(p(pts_raw), p(prm), p(app), p(self.gaussian_adapter.sh_mask), b, HW, d_sh,
                                                  i, G, expo, p(means), p(cov), p(harm), p(opac), p(scales), p(rots), st),
                           "s3r_gaussian_adapter")
        if visualization_dump is not None:  # keys consumed by export_ply (infer_model_re10k.py:542-557)
            visualization_dump["depth"] = means[..., 2].reshape(b, v, h, w, 1, 1)
            visualization_dump["scales"] = scales
            visualization_dump["rotations"] = rots
            visualization_dump["means"] = means.reshape(b, v, h, w, 1, 3)
            visualization_dump["opacities"] = opac.reshape(b, v, h, w, 1, 1)
        return Gaussians(means, cov, harm, opac)

    def _adapter_autograd(self, raw, b: int, v: int, h: int, w: int, global_step: int):
        """Head outputs -> Gaussians with autograd (same arithmetic as csrc/adapter.cu; reference:
        encoder_noposplat_multi_token_style.py:178-251, postprocess.py:45-61, gaussian_adapter.py:122-153,
        gaussians.py:8-44).  raw = per view (pts [b,3,h,w], params [b,8,h,w], app [b,3*d_sh,h,w])."""
        d_sh, expo = self.gaussian_adapter.d_sh, float(self.opacity_exponent(global_step))
        flat = lambda t: t.flatten(2).transpose(1, 2)                               # [b, HW, C]
        pts = torch.stack([flat(raw[3 * i]) for i in range(v)], dim=1).flatten(1, 2)        # [b, G, 3]
        prm = torch.stack([flat(raw[3 * i + 1]) for i in range(v)], dim=1).flatten(1, 2)    # [b, G, 8]
        app = torch.stack([flat(raw[3 * i + 2]) for i in range(v)], dim=1).flatten(1, 2)    # [b, G, 3*d_sh]
        d = pts.norm(dim=-1, keepdim=True)
        means = pts * (torch.expm1(d) / d.clamp(min=1e-8))
        pdf = torch.sigmoid(prm[..., 0])
        opac = pdf if expo == 1.0 else 0.5 * (1 - (1 - pdf) ** expo + pdf ** (1 / expo))
        scales = (0.001 * torch.nn.functional.softplus(prm[..., 1:4])).clamp(max=0.3)
        q = prm[..., 4:8]
        q = q / (q.norm(dim=-1, keepdim=True) + 1e-8)
        qi, qj, qk, qr = q.unbind(-1)                                               # xyzw
        two_s = 2.0 / ((q * q).sum(-1) + 1e-8)
        R = torch.stack((1 - two_s * (qj * qj + qk * qk), two_s * (qi * qj - qk * qr), two_s * (qi * qk + qj * qr),
                         two_s * (qi * qj + qk * qr), 1 - two_s * (qi * qi + qk * qk), two_s * (qj * qk - qi * qr),
                         two_s * (qi * qk - qj * qr), two_s * (qj * qk + qi * qr), 1 - two_s * (qi * qi + qj * qj)),
                        dim=-1).reshape(*q.shape[:-1], 3, 3)
        M = R * scales[..., None, :]
        # Sigma = M M^T as a broadcast product + sum: torch's batched matmul sends 1.3 M 3x3 problems to a library bmm that
        # takes 24 ms per call (forward and again in backward) at cfg5's batch - 14 % of the training step
        cov = (M.unsqueeze(-2) * M.unsqueeze(-3)).sum(-1)
        harm = app.reshape(*app.shape[:-1], 3, d_sh) * self.gaussian_adapter.sh_mask
        return means, cov, harm, opac, scales, q

    def get_data_shim(self):
        mean, std = self.cfg.input_mean, self.cfg.input_std

        def data_shim(batch):
            """apply_normalize_shim (src/dataset/shims/normalize_shim.py:21-27): context images -> (x - mean) / std."""
            ctx = dict(batch["context"])
            m = torch.tensor(mean, device=ctx["image"].device).view(1, 1, 3, 1, 1)
            s = torch.tensor(std, device=ctx["image"].device).view(1, 1, 3, 1, 1)
            ctx["image"] = (ctx["image"] - m) / s
            return {**batch, "context": ctx}

        return data_shim


class GraphedEncoder:
    """CUDA-graph replay of `encoder(context, style)` for inference: the ~3000 small launches of one forward
    (84 self- and 36 cross-attention modules, 5 DPT pyramids per view) are captured once per input shape and
    replayed with a single launch, which removes the Python / launch overhead that dominates at b=1.

        fast = GraphedEncoder(encoder, autocast_dtype=torch.bfloat16)
        gaussians = fast(context, style)           # tensors are copied into static buffers, outputs are static

    Outputs are views of static buffers: they are overwritten by the next call with the same shape."""

    def __init__(self, encoder: EncoderNoPoSplatMultiTokenStyle, autocast_dtype: Optional[torch.dtype] = None,
                 global_step: int = 0):
        self.encoder, self.autocast_dtype, self.global_step = encoder, autocast_dtype, global_step
        self._graphs: dict = {}

    def _run(self, context, style):
        with torch.no_grad(), torch.autocast("cuda", dtype=self.autocast_dtype or torch.bfloat16,
                                            enabled=self.autocast_dtype is not None):
            return self.encoder(context, style, self.global_step)

    def __call__(self, context: dict, style: dict) -> Gaussians:
        img, K, sty = context["image"], context["intrinsics"], style["image"]
        dev = next(self.encoder.parameters()).device  # inputs may live in (pinned) host memory: they are copied in
        key = (tuple(img.shape), tuple(sty.shape))
        if key not in self._graphs:
            static = dict(img=img.to(dev, copy=True), K=K.to(dev, copy=True), sty=sty.to(dev, copy=True))
            ctx, st = {"image": static["img"], "intrinsics": static["K"]}, {"image": static["sty"]}
            side = torch.cuda.Stream(device=dev)
            side.wait_stream(torch.cuda.current_stream(dev))
            with torch.cuda.stream(side):
                for _ in range(2):  # warm-up outside capture: lazy inits, cuDNN/cuBLAS plan selection, index caches
                    self._run(ctx, st)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                out = self._run(ctx, st)
            self._graphs[key] = (graph, static, out)
        graph, static, out = self._graphs[key]
        static["img"].copy_(img, non_blocking=True)
        static["K"].copy_(K, non_blocking=True)
        static["sty"].copy_(sty, non_blocking=True)
        graph.replay()
        return out


ENCODERS = {"noposplat_multi_token_style": (EncoderNoPoSplatMultiTokenStyle, None)}


def get_encoder(cfg):
    encoder, visualizer = ENCODERS[cfg.name]
    return encoder(cfg), visualizer
