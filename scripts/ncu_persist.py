"""ncu target: the persistent CTA-pair GEMM (bias / +gelu) and convolution, 2 launches each."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200 import _lib
from styl3r_b200.gemm import linear
from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
L = _lib.lib()
M, N, K = 4112, 3072, 1024
x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
b = torch.randn(N, device="cuda").to(torch.bfloat16); r = torch.randn(M, N, device="cuda").to(torch.bfloat16)
for _ in range(2): linear(x, w, b)
for _ in range(2): linear(x, w, b, gelu=True)
x2 = torch.randn(4096, 8192, device="cuda").to(torch.bfloat16); w2 = (torch.randn(4096, 8192, device="cuda") / 90).to(torch.bfloat16)
for _ in range(2): linear(x2, w2, None)
for _ in range(2): linear(x, w, b, residual=r)
xc = torch.randn(16, 128, 128, 256, device="cuda").to(torch.bfloat16)
wp = prep_conv_weight((torch.randn(256, 256, 3, 3, device="cuda") / 48).to(torch.bfloat16))
for _ in range(2): conv2d_nhwc(xc, wp, (3, 3), bias=b[:256], relu=True)
torch.cuda.synchronize()
