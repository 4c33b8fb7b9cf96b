"""ncu target: the tcgen05 attention kernel at the encoder's two shape classes (B2 H16 257^2: ViT-L self-attention at cfg2;
B4 H12 1028^2: stylizer decoder self-attention at cfg3), inside a cudaProfilerStart/Stop window.
  ncu --set full --import-source on --clock-control none --profile-from-start off -k regex:s3r_ -o OUT python scripts/ncu_attention.py"""
import sys
from pathlib import Path

sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch

from styl3r_b200.ops import attention_bf16

dev = torch.device("cuda", 0)
q = torch.randn(2, 257, 3, 16, 64, device=dev).to(torch.bfloat16)
q4 = torch.randn(4, 1028, 3, 12, 64, device=dev).to(torch.bfloat16)
for t in (q, q4):
    for _ in range(3):
        attention_bf16(t[:, :, 0], t[:, :, 1], t[:, :, 2], 0.125)
torch.cuda.synchronize()
torch.cuda.profiler.start()
for t in (q, q4):
    attention_bf16(t[:, :, 0], t[:, :, 1], t[:, :, 2], 0.125)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print("done")
