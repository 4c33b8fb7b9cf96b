"""Development aid (GPU): the three 12-layer decoders of the cfg2 forward under CUDA-graph replay - each alone, and
together with / without the stream branches."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200 import _lib
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
from styl3r_b200.streams import fork_join
from tests.encoder_weights import make_inputs
enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval().to_inference(torch.bfloat16)
def graph_time(fn, iters=10):
    side = torch.cuda.Stream(); side.wait_stream(torch.cuda.current_stream())
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
context, style = make_inputs(1, 2, 256, seed=1, device="cuda")
ctx = {**context, "image": context["image"].to(torch.bfloat16), "intrinsics": context["intrinsics"].to(torch.bfloat16)}
sty = {"image": style["image"].to(torch.bfloat16)}
with torch.no_grad():
    f, p = enc.backbone.encode_views(ctx)
    sy, sp = enc.token_stylizer.encode_style(sty)
for pdl in (1, 0):
    _lib.check(_lib.lib().s3r_set_tunable(5, pdl))
    t_bb_seq = graph_time(lambda: enc.backbone.decode_views(f, p, parallel=False))
    t_bb_par = graph_time(lambda: enc.backbone.decode_views(f, p, parallel=True))
    t_st_seq = graph_time(lambda: enc.token_stylizer.decode(sy, sp, f, p, parallel=False))
    t_st_par = graph_time(lambda: enc.token_stylizer.decode(sy, sp, f, p, parallel=True))
    t_all = graph_time(lambda: fork_join([lambda: enc.backbone.decode_views(f, p, parallel=True), lambda: enc.token_stylizer.decode(sy, sp, f, p, parallel=True)]))
    blk = enc.backbone.dec_blocks[0]
    x = torch.randn(1, 257, 768, device="cuda").to(torch.bfloat16); pos = p[:, 0]
    t_blk = graph_time(lambda: blk(x, x, pos, pos, parallel=False))
    t_blk_p = graph_time(lambda: blk(x, x, pos, pos, parallel=True))
    print(f"pdl={pdl}: backbone decoder seq {t_bb_seq:.2f} / branches {t_bb_par:.2f} ms | stylizer decoder seq {t_st_seq:.2f} / branches {t_st_par:.2f} | both {t_all:.2f} | one DecoderBlock (M=257) seq {t_blk*1e3:.0f} us / kv-branch {t_blk_p*1e3:.0f} us", flush=True)
