"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement of the encoder forward: the module tree of `styl3r_b200.encoder` - whose
parameter registry equals the reference's - evaluated with plain PyTorch ops on the host: the reference's own PyTorch
`RoPE2D` arithmetic (src/model/encoder/backbone/croco/pos_embed.py:112-159) instead of the CUDA RoPE kernel, SDPA for
`xformers.ops.memory_efficient_attention` (croco/blocks.py:126-130,192-196), the torch modules of the DPT heads and the
torch restatement of the adapter (gaussian_adapter.py:122-153, gaussians.py:8-44, postprocess.py:45-61).

PINNED: tests/test_encoder_cpu.py compares it against tests/golden/encoder_golden.npz, which was produced by the
REFERENCE encoder itself (tests/golden/make_encoder_golden.py).  Also the `cpu_baseline` of the encoder in bench.py
(kind "port": the reference's Python cannot travel to the GPU box)."""
from __future__ import annotations

from unittest import mock

import torch


def rope2d_bnhd(tokens: torch.Tensor, positions: torch.Tensor, base: float) -> torch.Tensor:
    """RoPE2D.forward (pos_embed.py:142-159) on the [B, N, H, D] layout our blocks use (the reference transposes to
    [B, H, N, D] around the call, blocks.py:104-106)."""
    t = tokens.transpose(1, 2)                                  # [B, H, N, D]
    D = t.shape[-1] // 2
    inv_freq = 1.0 / (base ** (torch.arange(0, D, 2, device=t.device).float() / D))
    pos_t = torch.arange(int(positions.max()) + 1, device=t.device, dtype=inv_freq.dtype)
    freqs = torch.einsum("i,j->ij", pos_t, inv_freq).to(t.dtype)
    freqs = torch.cat((freqs, freqs), dim=-1)
    cos, sin = freqs.cos(), freqs.sin()

    def rot_half(x):
        x1, x2 = x[..., : x.shape[-1] // 2], x[..., x.shape[-1] // 2:]
        return torch.cat((-x2, x1), dim=-1)

    def rope1d(x, pos1d):
        c = torch.nn.functional.embedding(pos1d, cos)[:, None, :, :]
        s = torch.nn.functional.embedding(pos1d, sin)[:, None, :, :]
        return x * c + rot_half(x) * s

    y, x = t.chunk(2, dim=-1)
    out = torch.cat((rope1d(y, positions[:, :, 0]), rope1d(x, positions[:, :, 1])), dim=-1)
    return out.transpose(1, 2)


@torch.no_grad()
def encoder_forward(enc, context: dict, style: dict, global_step: int = 0):
    """(means [b,G,3], covariances [b,G,3,3], harmonics [b,G,3,d_sh], opacities [b,G], scales, rotations) on the host."""
    img = context["image"]
    b, v, _, h, w = img.shape
    with mock.patch("styl3r_b200.encoder.vit._rope", rope2d_bnhd):
        feat, pos = enc.backbone.encode_views(context)
        sty_y, sty_pos = enc.token_stylizer.encode_style(style)
        dec_feat = enc.backbone.decode_views(feat, pos)
        sty_feat = enc.token_stylizer.decode(sty_y, sty_pos, feat, pos)
    raw = []
    for i in range(v):
        toks = [t[:, i].float() for t in dec_feat]
        raw.append((enc.downstream_head1 if i == 0 else enc.downstream_head2)(toks, (h, w)))
        raw.append((enc.gaussian_param_head if i == 0 else enc.gaussian_param_head2)(toks, (h, w), img[:, i, :3].float()))
        raw.append(enc.gaussian_appearance_head([t[:, i].float() for t in sty_feat], (h, w)))
    return enc._adapter_autograd(raw, b, v, h, w, global_step)


def time_encoder(n_threads: int, runs: int = 2, seed: int = 0):
    """Seconds per scene of the cfg1 workload (b=1, v=2, 256x256 + style image, random weights) on `n_threads` host
    threads: 1 warm-up + `runs` timed forwards, median."""
    import time
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
    torch.set_num_threads(n_threads)
    torch.manual_seed(seed)
    enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
    enc = enc.eval()
    g = torch.Generator().manual_seed(1234)
    ctx = {"image": torch.rand(1, 2, 3, 256, 256, generator=g) * 2 - 1,
           "intrinsics": torch.tensor([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]]).expand(1, 2, 3, 3).contiguous()}
    sty = {"image": torch.rand(1, 3, 256, 256, generator=g) * 2 - 1}
    encoder_forward(enc, ctx, sty)
    ts = []
    for _ in range(runs):
        t0 = time.perf_counter()
        encoder_forward(enc, ctx, sty)
        ts.append(time.perf_counter() - t0)
    return sorted(ts)[len(ts) // 2]
