"""Development aid (GPU): cluster-multicast variants of the tcgen05 GEMM / conv - correctness vs torch and time
(CUDA-graph replay of 50 launches) per (cluster shape, big tile) on the encoder's shapes."""
import subprocess, sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200 import _lib
from styl3r_b200.gemm import linear
L = _lib.lib()

def gtime(fn, reps=50):
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
    return e0.elapsed_time(e1) / reps

shapes = [(257, 2304, 768), (257, 768, 768), (257, 3072, 768), (257, 768, 3072), (514, 3072, 1024), (514, 1024, 1024),
          (514, 4096, 1024), (514, 1024, 4096), (1028, 2304, 768), (4112, 3072, 1024), (4112, 1024, 1024), (4112, 4096, 1024),
          (4112, 1024, 4096), (8192, 8192, 8192)]
only = sys.argv[1:]  # optional subset of cluster codes
for M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16)
    ref = torch.nn.functional.linear(x.float(), w.float(), b.float())
    tt = gtime(lambda: torch.nn.functional.linear(x, w, b), 5 if M >= 8192 else 50)
    line = f"M={M:5d} N={N:5d} K={K:5d} torch {tt*1e3:6.1f} |"
    for big in (0, 1):
        for code in (0, 12, 14, 21, 22, 24):
            if only and str(code) not in only: continue
            if big and M < 2048: continue
            _lib.check(L.s3r_set_tunable(2, code)); _lib.check(L.s3r_set_tunable(3, big))
            y = linear(x, w, b); torch.cuda.synchronize()
            err = (y.float() - ref).abs().max().item()
            ok = "" if err < 0.06 else f"!!ERR {err:.2e}"
            t = gtime(lambda: linear(x, w, b), 5 if M >= 8192 else 50)
            line += f" {'B' if big else ''}c{code}: {t*1e3:6.1f}{ok}"
    print(line, flush=True)
_lib.check(L.s3r_set_tunable(2, 0)); _lib.check(L.s3r_set_tunable(3, 0))
# conv with weight-tile multicast
from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
import torch.nn.functional as F
for (n, h, w_, ci, co) in [(1, 256, 256, 256, 256), (4, 256, 256, 256, 256), (1, 256, 256, 128, 128), (1, 128, 128, 256, 128), (1, 64, 64, 256, 256), (4, 64, 64, 256, 256), (4, 16, 16, 256, 256)]:
    x = torch.randn(n, ci, h, w_, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(co, ci, 3, 3, device="cuda") / (ci * 9) ** 0.5).to(torch.bfloat16)
    ref = F.conv2d(x.float(), wt.float(), None, 1, 1)
    xn, wp = x.permute(0, 2, 3, 1).contiguous(), prep_conv_weight(wt)
    line = f"conv n={n} {h}x{w_} {ci}->{co}:"
    for cl in (0, 1):
        _lib.check(L.s3r_set_tunable(4, cl))
        y = conv2d_nhwc(xn, wp, (3, 3)); torch.cuda.synchronize()
        err = (y.float().permute(0, 3, 1, 2) - ref).abs().max().item()
        t = gtime(lambda: conv2d_nhwc(xn, wp, (3, 3)), 20)
        line += f"  cluster={cl}: {t*1e3:6.1f} us {2.0*n*h*w_*ci*co*9/t/1e9:5.0f} TF/s err {err:.1e}"
    print(line, flush=True)
_lib.check(L.s3r_set_tunable(4, 0))
