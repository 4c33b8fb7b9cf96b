"""Split-K on/off for the narrow-N / long-K GEMMs of the batch-1 encoder (fc2, proj).  Development aid."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.gemm import linear


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


for M, N, K in [(257, 768, 3072), (257, 768, 768), (514, 1024, 4096), (514, 1024, 1024), (257, 1024, 4096), (257, 2304, 768), (514, 768, 1024), (1028, 768, 3072)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = torch.randn(N, K, device="cuda").to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16); r = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    from styl3r_b200 import _lib
    L = _lib.lib()
    res = {}
    outs = {}
    for ks in (1, 0, 2, 4, 8, 9):  # S3R_TUNE_GEMM_KSPLIT: never / auto / forced 64-wide KS / 128-wide KS4 forced / forbidden
        L.s3r_set_tunable(10, ks)
        res[ks] = gtime(lambda: linear(x, w, b, residual=r))
        outs[ks] = linear(x, w, b, residual=r).float()
    L.s3r_set_tunable(10, 0)
    t = gtime(lambda: r + torch.nn.functional.linear(x, w, b))
    ref = r.float() + x.float() @ w.float().t() + b.float()
    err = {k: ((v - ref).abs().max() / ref.abs().max()).item() for k, v in outs.items()}
    print(f"M={M:5d} N={N:5d} K={K:5d} +res: never {res[1]*1e3:6.1f} | auto {res[0]*1e3:6.1f} | KS2 {res[2]*1e3:6.1f} | KS4 {res[4]*1e3:6.1f} | 128-wide KS4 {res[8]*1e3:6.1f} | auto w/o it {res[9]*1e3:6.1f} | torch {t*1e3:6.1f} us | rel err " + " ".join(f"{k}:{e:.4f}" for k, e in err.items()), flush=True)
