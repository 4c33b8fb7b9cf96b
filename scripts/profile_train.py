"""Per-kernel device-time breakdown of one cfg5 training step (torch profiler).  usage: profile_train.py [batch]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import numpy as np
import torch
from torch.profiler import profile, ProfilerActivity
import bench_legs as bl
from styl3r_b200 import synthetic as syn
from styl3r_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
from styl3r_b200.train import IdentityLoss, LossStyle, LossStyleCfg, LossStyleCfgWrapper, TrainStep
batch = int(sys.argv[1]) if len(sys.argv) > 1 else 10
dev = torch.device("cuda"); HW = 256; V = 4
enc = bl.make_encoder(dev, inference=False); enc.to_training(torch.bfloat16)
dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True)).to(dev)
step = TrainStep(enc, dec, [LossStyle(LossStyleCfgWrapper(LossStyleCfg(10.0))).to(dev)], IdentityLoss().to(dev))
g = torch.Generator().manual_seed(7)
K = torch.tensor([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]])
scs = [syn.make_scene(seed=100 + s, v=2, V=V, hw=8) for s in range(batch)]
bd = {"context": {"image": torch.rand(batch, 2, 3, HW, HW, generator=g).to(dev), "intrinsics": K.expand(batch, 2, 3, 3).contiguous().to(dev),
                  "extrinsics": torch.as_tensor(np.stack([s["context_extrinsics"] for s in scs])).to(dev),
                  "near": torch.full((batch, 2), 0.1, device=dev), "far": torch.full((batch, 2), 100.0, device=dev)},
      "target": {"image": torch.rand(batch, V, 3, HW, HW, generator=g).to(dev), "intrinsics": K.expand(batch, V, 3, 3).contiguous().to(dev),
                 "extrinsics": torch.as_tensor(np.stack([s["extrinsics"] for s in scs])).to(dev),
                 "near": torch.full((batch, V), 0.1, device=dev), "far": torch.full((batch, V), 100.0, device=dev)},
      "style": {"image": torch.rand(batch, 3, HW, HW, generator=g).to(dev)}}
enc.train()
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
for _ in range(2): step(bd)
torch.cuda.synchronize()
SHAPES = len(sys.argv) > 2
with profile(activities=[ProfilerActivity.CUDA] + ([ProfilerActivity.CPU] if SHAPES else []), record_shapes=SHAPES) as prof:
    step(bd); torch.cuda.synchronize()
if SHAPES:  # which ATen ops (with input shapes) own the device time
    evs = prof.key_averages(group_by_input_shape=True)
    for e in sorted(evs, key=lambda e: -e.device_time_total)[:45]:
        print(f"{e.device_time_total/1000:8.2f} ms n={e.count:5d} {e.key[:50]:50s} {str(e.input_shapes)[:150]}")
    sys.exit(0)
ev = prof.key_averages()
tot = sum(e.device_time_total for e in ev)
print(f"batch {batch}: total device time {tot/1000:.1f} ms, {sum(e.count for e in ev)} kernels")
for e in sorted(ev, key=lambda e: -e.device_time_total)[:40]:
    print(f"{e.device_time_total/1000:8.2f} ms {100*e.device_time_total/tot:5.1f}% n={e.count:5d} avg {e.device_time_total/e.count:8.1f} us  {e.key[:110]}")
