"""Development aid (GPU): kernel timeline (start, duration, stream) of ONE graph replay of the encoder forward, from
the torch profiler (CUPTI), written as compact JSON for offline critical-path analysis."""
import json, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, GraphedEncoder, get_encoder
from tests.encoder_weights import make_inputs
b, v = (int(sys.argv[1]), int(sys.argv[2])) if len(sys.argv) > 2 else (1, 2)
enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval().to_inference(torch.bfloat16)
context, style = make_inputs(b, v, 256, seed=1, device="cuda")
fast = GraphedEncoder(enc)
for _ in range(3): fast(context, style)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    fast(context, style); torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
rows = sorted(((e.time_range.start, e.time_range.end - e.time_range.start, getattr(e, "stream", -1) if hasattr(e, "stream") else -1, e.name[:60]) for e in ev))
t0 = rows[0][0]
out = [[round(s - t0, 2), round(d, 2), st, n] for s, d, st, n in rows]
Path("gpurun_out").mkdir(exist_ok=True)
json.dump(out, open(f"gpurun_out/enc_timeline_b{b}v{v}.json", "w"))
print(len(out), "kernels, span", out[-1][0] + out[-1][1], "us")
