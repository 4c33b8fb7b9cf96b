"""Development aid (GPU): graphed encoder time for cfg2 / cfg3 with PDL on / off."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200 import _lib
L = _lib.lib()
def t_ms(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder, GraphedEncoder
from tests.encoder_weights import make_inputs
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval()
enc.to_inference(torch.bfloat16)
for pdl in (0, 1):
    _lib.check(L.s3r_set_tunable(5, pdl))
    for (b, v) in ((1, 2), (4, 4)):
        context, style = make_inputs(b, v, 256, seed=1, device="cuda")
        fast = GraphedEncoder(enc)
        out = fast(context, style); torch.cuda.synchronize()
        ms = t_ms(lambda: fast(context, style), 10)
        flops = {2: 1270.8e9, 4: 2437.2e9}[v] * b
        print(f"GRAPH encoder pdl={pdl} b={b} v={v}: {ms:.2f} ms {flops/ms/1e9:.1f} TFLOP/s  means|mean| {out.means.abs().mean().item():.4f}", flush=True)
