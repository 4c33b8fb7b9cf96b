"""tcgen05 GEMM vs torch (cuBLAS) GPU time on the encoder's shapes, measured by CUDA-graph replay of 50 launches
(removes the Python/launch overhead that dominates at these sizes).  Development aid."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.gemm import linear
shapes = [(257, 2304, 768), (257, 768, 768), (257, 3072, 768), (257, 768, 3072), (256, 3072, 1024),
          (514, 3072, 1024), (514, 1024, 1024), (514, 4096, 1024), (514, 1024, 4096), (4112, 3072, 1024),
          (4112, 4096, 1024), (4112, 1024, 4096), (8192, 8192, 8192)]
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
for M, N, K in shapes:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16); r = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    reps = 5 if M >= 8192 else 50
    from styl3r_b200 import _lib
    _lib.lib().s3r_set_tunable(13, 2)
    s1 = gtime(lambda: linear(x, w, b), reps); s2 = gtime(lambda: linear(x, w, b, gelu=True), reps); s3 = gtime(lambda: linear(x, w, b, residual=r), reps)
    _lib.lib().s3r_set_tunable(13, 0)
    o1 = gtime(lambda: linear(x, w, b), reps); t1 = gtime(lambda: torch.nn.functional.linear(x, w, b), reps)
    o2 = gtime(lambda: linear(x, w, b, gelu=True), reps); t2 = gtime(lambda: torch.nn.functional.gelu(torch.nn.functional.linear(x, w, b)), reps)
    o3 = gtime(lambda: linear(x, w, b, residual=r), reps); t3 = gtime(lambda: r + torch.nn.functional.linear(x, w, b), reps)
    fl = 2.0 * M * N * K
    print(f"M={M:5d} N={N:5d} K={K:5d} | bias: ours {o1*1e3:7.1f} us ({fl/o1/1e9:5.0f} TF) torch {t1*1e3:7.1f} | +gelu: {o2*1e3:7.1f} vs {t2*1e3:7.1f} | +res: {o3*1e3:7.1f} vs {t3*1e3:7.1f} | staged epilogue: {s1*1e3:6.1f}/{s2*1e3:6.1f}/{s3*1e3:6.1f}", flush=True)
