"""CTA-pair (tcgen05 cta_group::2) GEMM / convolution vs the single-CTA tiles and torch: correctness (max abs error
against an fp32 reference of the same bf16 operands) and CUDA-graph timed launches.  Development aid.
usage: bench_pair.py [gemm|conv|all]"""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
from styl3r_b200 import _lib
from styl3r_b200.gemm import linear
L = _lib.lib()
PAIR = 11
what = sys.argv[1] if len(sys.argv) > 1 else "all"
MODES = ((2, "single"), (1, "pair128"), (3, "pair256"), (4, "persist128"), (5, "persist256"))


def gtime(fn, reps=20):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    torch.cuda.synchronize()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


if what in ("gemm", "all"):
    shapes = [(514, 3072, 1024), (514, 1024, 1024), (1028, 3072, 1024), (4112, 3072, 1024), (4112, 4096, 1024), (4112, 1024, 4096),
              (4112, 1024, 1024), (4112, 2304, 768), (4112, 768, 3072), (4100, 3072, 1024), (8192, 8192, 8192)]
    for M, N, K in shapes:
        x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
        b = torch.randn(N, device="cuda").to(torch.bfloat16); r = torch.randn(M, N, device="cuda").to(torch.bfloat16)
        ref = x.float() @ w.float().t() + b.float()
        reps = 5 if M >= 8192 else 20
        row = f"M={M:5d} N={N:5d} K={K:5d} |"
        for mode, name in MODES:
            L.s3r_set_tunable(PAIR, mode)
            y = linear(x, w, b)
            err = (y.float() - ref).abs().max().item()
            yg = linear(x, w, b, gelu=True); errg = (yg.float() - F.gelu(ref)).abs().max().item()
            yr = linear(x, w, b, residual=r); errr = (yr.float() - (ref + r.float())).abs().max().item()
            t = gtime(lambda: linear(x, w, b), reps); tg = gtime(lambda: linear(x, w, b, gelu=True), reps)
            tr = gtime(lambda: linear(x, w, b, residual=r), reps)
            row += f" {name}: {t:6.1f}/{tg:6.1f}/{tr:6.1f} us ({2.0*M*N*K/t/1e6:5.0f} TF) err {err:.3f}/{errg:.3f}/{errr:.3f} |"
        L.s3r_set_tunable(PAIR, 0)
        t = gtime(lambda: F.linear(x, w, b), reps)
        print(row + f" torch {t:6.1f}", flush=True)

if what in ("conv", "all"):
    from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
    for n, hw, cin, cout in [(2, 16, 256, 256), (2, 32, 256, 256), (2, 64, 256, 256), (2, 128, 256, 256), (4, 256, 256, 256), (16, 128, 256, 256),
                             (16, 64, 256, 256), (2, 256, 128, 128), (16, 256, 128, 128)]:
        x = torch.randn(n, hw, hw, cin, device="cuda").to(torch.bfloat16)
        w = (torch.randn(cout, cin, 3, 3, device="cuda") / (9 * cin) ** 0.5)
        b = torch.randn(cout, device="cuda").to(torch.bfloat16)
        wp = prep_conv_weight(w)
        ref = F.conv2d(x.float().permute(0, 3, 1, 2), w.to(torch.bfloat16).float(), b.float(), padding=1).permute(0, 2, 3, 1)
        row = f"conv n={n:2d} {hw:3d}^2 {cin}->{cout} |"
        fl = 2.0 * n * hw * hw * cin * cout * 9
        for mode, name in MODES:
            L.s3r_set_tunable(PAIR, mode)
            y = conv2d_nhwc(x, wp, (3, 3), bias=b)
            err = (y.float() - ref).abs().max().item()
            t = gtime(lambda: conv2d_nhwc(x, wp, (3, 3), bias=b), 10)
            row += f" {name}: {t:7.1f} us ({fl/t/1e6:5.0f} TF) err {err:.3f} |"
        L.s3r_set_tunable(PAIR, 0)
        print(row, flush=True)
