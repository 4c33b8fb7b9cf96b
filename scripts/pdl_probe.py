"""Development aid (GPU): per-launch time of a stream-ordered chain of GEMMs inside a CUDA graph with and without
programmatic dependent launch (S3R_TUNE_PDL)."""
import sys
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
    return e0.elapsed_time(e1) / reps * 1e3

for M, N, K in [(257, 768, 768), (514, 1024, 1024), (514, 3072, 1024), (514, 1024, 4096), (4112, 3072, 1024)]:
    x = torch.randn(M, K, device="cuda").to(torch.bfloat16); w = (torch.randn(N, K, device="cuda") / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda").to(torch.bfloat16)
    ref = torch.nn.functional.linear(x.float(), w.float(), b.float())
    line = f"M={M} N={N} K={K}:"
    for pdl in (0, 1):
        _lib.check(L.s3r_set_tunable(5, pdl))
        y = linear(x, w, b); torch.cuda.synchronize()
        err = (y.float() - ref).abs().max().item()
        # a dependent chain: y_{i+1} = f(y_i) keeps true producer -> consumer edges between consecutive launches
        if N == K:
            def chain():
                t = x
                for _ in range(4): t = linear(t, w, b)
                return t
            tc = gtime(chain, 12) / 4
        else:
            tc = float("nan")
        line += f"  pdl={pdl}: {gtime(lambda: linear(x, w, b)):6.2f} us/launch (dependent chain {tc:6.2f}) err {err:.1e} |"
    print(line, flush=True)
_lib.check(L.s3r_set_tunable(5, 0))
# dependent-chain correctness under PDL (graph replay): result must equal the non-PDL chain bit for bit
x = torch.randn(514, 1024, device="cuda").to(torch.bfloat16); w = (torch.randn(1024, 1024, device="cuda") / 32).to(torch.bfloat16)
def chain():
    t = x
    for _ in range(6): t = linear(t, w, None)
    return t
ref = chain().clone()
_lib.check(L.s3r_set_tunable(5, 1))
s = torch.cuda.Stream()
with torch.cuda.stream(s):
    chain()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=s):
        out = chain()
torch.cuda.synchronize()
for _ in range(5): g.replay()
torch.cuda.synchronize()
print("PDL dependent chain equals plain:", torch.equal(out, ref))
_lib.check(L.s3r_set_tunable(5, 0))
