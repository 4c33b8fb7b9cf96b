import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.gemm import linear
for M, N, K in [(257, 768, 768), (257, 2304, 768), (257, 3072, 768)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16)
    for _ in range(3): linear(x, w, b)
torch.cuda.synchronize()
