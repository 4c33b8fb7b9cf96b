"""Encoder timing on one GPU (development aid): fp32 / tf32 / bf16-autocast, cfg2 (b=1,v=2) and cfg3 (b=4,v=4)."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
from tests.encoder_weights import make_inputs

enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
enc = enc.cuda().eval()

def run(b, v, mode, iters=5):
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    context, style = make_inputs(b, v, 256, seed=1, device="cuda")
    def step():
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
            return enc(context, style)
    for _ in range(2): step()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): step()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = {2: 1270.8e9, 4: 2437.2e9}.get(v, 0) * b
    print(f"b={b} v={v} {mode}: {ms:.2f} ms/scene-batch  {flops / ms / 1e9:.1f} TFLOP/s", flush=True)

def run_graph(b, v, mode, iters=10):
    from styl3r_b200.encoder import GraphedEncoder
    torch.backends.cuda.matmul.allow_tf32 = mode != "fp32"
    torch.backends.cudnn.allow_tf32 = mode != "fp32"
    context, style = make_inputs(b, v, 256, seed=1, device="cuda")
    fast = GraphedEncoder(enc, autocast_dtype=torch.bfloat16 if mode == "bf16" else None)
    ref = fast(context, style).means.clone()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fast(context, style)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = {2: 1270.8e9, 4: 2437.2e9}.get(v, 0) * b
    with torch.no_grad():
        eager = enc(context, style).means
    print(f"GRAPH b={b} v={v} {mode}: {ms:.2f} ms  {flops / ms / 1e9:.1f} TFLOP/s  (graph vs eager fp32 max rel diff "
          f"{float((ref - eager).abs().max() / eager.abs().max()):.2e})", flush=True)

for mode in ("fp32", "tf32", "bf16"):
    run(1, 2, mode)
for mode in ("tf32", "bf16"):
    run_graph(1, 2, mode)
run_graph(4, 4, "bf16", iters=3)
enc.to_inference(torch.bfloat16)
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
for (b, v) in ((1, 2), (4, 4)):
    from styl3r_b200.encoder import GraphedEncoder
    context, style = make_inputs(b, v, 256, seed=1, device="cuda")
    fast = GraphedEncoder(enc)
    fast(context, style); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5): fast(context, style)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    flops = {2: 1270.8e9, 4: 2437.2e9}[v] * b
    print(f"GRAPH to_inference(bf16 trunks, channels_last heads) b={b} v={v}: {ms:.2f} ms {flops / ms / 1e9:.1f} TFLOP/s", flush=True)
