"""Development aid (GPU): critical-path breakdown of the graphed encoder forward (cfg2 / cfg3) - CUDA graphs of
(a) the two ViT-L encoders, (b) + the decoders, (c) the full forward; differences = stage times."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
from styl3r_b200.streams import fork_join
from tests.encoder_weights import make_inputs

enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval().to_inference(torch.bfloat16)

def graph_time(fn, iters=10):
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side), torch.no_grad():
        for _ in range(2): fn()
    torch.cuda.current_stream().wait_stream(side); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.no_grad(), torch.cuda.graph(g):
        fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

import ctypes
from styl3r_b200 import _lib
shallow = int(sys.argv[1]) if len(sys.argv) > 1 else 0
_lib.check(_lib.lib().s3r_set_tunable(6, shallow))
print("gemm shallow ring:", shallow)
for (b, v) in ((1, 2), (4, 4)):
    context, style = make_inputs(b, v, 256, seed=1, device="cuda")
    ctx = {**context, "image": context["image"].to(torch.bfloat16), "intrinsics": context["intrinsics"].to(torch.bfloat16)}
    sty = {"image": style["image"].to(torch.bfloat16)}
    def stage_a():
        return fork_join([lambda: enc.backbone.encode_views(ctx), lambda: enc.token_stylizer.encode_style(sty)])
    def stage_b():
        (f, p), (sy, sp) = stage_a()
        return fork_join([lambda: enc.backbone.decode_views(f, p, parallel=True), lambda: enc.token_stylizer.decode(sy, sp, f, p, parallel=True)])
    def content_only():
        return enc.backbone.encode_views(ctx)
    def full():
        return enc(context, style)
    ta, tb, tc, t1 = graph_time(stage_a), graph_time(stage_b), graph_time(full), graph_time(content_only)
    print(f"b={b} v={v}: content ViT alone {t1:.2f} ms | both ViT-L encoders {ta:.2f} | + decoders {tb:.2f} (decoders {tb-ta:.2f}) | full {tc:.2f} (heads + adapter {tc-tb:.2f})", flush=True)
