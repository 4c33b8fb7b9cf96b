import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.gemm import linear
for M, N, K in [(4112, 4096, 1024), (514, 3072, 1024), (8192, 8192, 8192)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16)
    for _ in range(3): linear(x, w, b, gelu=True)
torch.cuda.synchronize()
