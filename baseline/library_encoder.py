"""GPU comparator for the encoder (bench.py key `encoder_comparator`): the SAME nn.Module tree as styl3r_b200.encoder,
routed through the torch libraries instead of our kernels (`styl3r_b200.ops.FORCE_LIBRARY`).  Comparator only - never
imported by the product.

  library_bf16_graph   bf16 ViT trunks on cuBLAS (nn.Linear), SDPA (flash / cuDNN attention), ATen LayerNorm / GELU; DPT
                       heads fp32 channels_last on cuDNN with TF32 (the reference switches autocast off for the heads,
                       encoder_noposplat_multi_token_style.py:150); same concurrent stream branches; CUDA-graph replay
  library_tf32_eager   fp32 modules, TF32 matmuls / convolutions, eager launches - the reference's own configuration
                       (croco.py:13) on this GPU
"""
from __future__ import annotations

import torch


def _time(fn, iters):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def time_library_encoder(dev, ctx, sty, iters=10):
    from styl3r_b200 import ops
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, GraphedEncoder, get_encoder
    res = {}
    tf32 = (torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32)
    torch.backends.cuda.matmul.allow_tf32 = True
    torch.backends.cudnn.allow_tf32 = True
    try:
        with torch.no_grad():
            torch.manual_seed(0)
            enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
            enc = enc.to(dev).eval()
            res["library_tf32_eager_ms"] = _time(lambda: enc(ctx, sty), max(3, iters // 2))
            enc = enc.to_inference(torch.bfloat16, heads="cudnn")
            ops.FORCE_LIBRARY = True
            fast = GraphedEncoder(enc)
            res["library_bf16_graph_ms"] = _time(lambda: fast(ctx, sty), iters)
            del fast, enc
    finally:
        ops.FORCE_LIBRARY = False
        torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.allow_tf32 = tf32
        torch.cuda.empty_cache()
    return res
