"""Per-kernel device-time breakdown of one encoder forward (torch profiler, eager launches, branches off so that the
kernel times are not inflated by concurrency).  usage: profile_encoder.py [inf|bf16|fp32] [b] [v]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from torch.profiler import profile, ProfilerActivity
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
from tests.encoder_weights import make_inputs
mode = sys.argv[1] if len(sys.argv) > 1 else "inf"
b = int(sys.argv[2]) if len(sys.argv) > 2 else 1
v = int(sys.argv[3]) if len(sys.argv) > 3 else 2
enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval()
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
if mode == "inf":
    enc.to_inference(torch.bfloat16, branches=False)
context, style = make_inputs(b, v, 256, seed=1, device="cuda")
def step():
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.bfloat16, enabled=(mode == "bf16")):
        return enc(context, style)
for _ in range(2): step()
torch.cuda.synchronize()
SHAPES = len(sys.argv) > 4
with profile(activities=[ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if SHAPES else []), record_shapes=SHAPES) as prof:
    step(); torch.cuda.synchronize()
if SHAPES:  # which ATen ops (with input shapes) own the glue kernels
    evs = [e for e in prof.key_averages(group_by_input_shape=True) if e.key.startswith("aten::") and e.device_time_total > 0]
    for e in sorted(evs, key=lambda e: -e.device_time_total)[:40]:
        print(f"{e.device_time_total/1000:8.3f} ms n={e.count:4d} {e.key[:32]:32s} {str(e.input_shapes)[:140]}")
    sys.exit(0)
ev = prof.key_averages()
rows = sorted(ev, key=lambda e: -e.device_time_total)[:32]
tot = sum(e.device_time_total for e in ev)
print(f"mode={mode} b={b} v={v}: total device time {tot/1000:.2f} ms, {sum(e.count for e in ev)} kernels")
for e in rows:
    print(f"{e.device_time_total/1000:8.3f} ms {100*e.device_time_total/tot:5.1f}% n={e.count:4d} avg {e.device_time_total/e.count:7.1f} us  {e.key[:100]}")
