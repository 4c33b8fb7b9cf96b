"""tcgen05 attention vs PyTorch SDPA GPU time on the encoder's shapes (CUDA-graph replay of 30 launches)."""
import sys
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch, torch.nn.functional as F
from styl3r_b200.ops import attention_bf16
def gtime(fn, reps=30):
    fn(); torch.cuda.synchronize()
    s = torch.cuda.Stream()
    with torch.cuda.stream(s):
        fn()
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=s):
            for _ in range(reps): fn()
    torch.cuda.synchronize(); g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1000
for B, H, Nq, Nk in [(2, 16, 257, 257), (1, 16, 256, 256), (1, 12, 257, 257), (1, 12, 257, 514), (1, 12, 514, 514), (1, 12, 514, 256),
                     (16, 16, 257, 257), (4, 12, 1028, 1028), (12, 12, 257, 771)]:
    q = torch.randn(B, Nq, H, 64, device="cuda").to(torch.bfloat16); k = torch.randn(B, Nk, H, 64, device="cuda").to(torch.bfloat16); v = torch.randn_like(k)
    from styl3r_b200 import _lib
    _lib.lib().s3r_set_tunable(12, 2); two = gtime(lambda: attention_bf16(q, k, v, 0.125))
    _lib.lib().s3r_set_tunable(12, 1); one = gtime(lambda: attention_bf16(q, k, v, 0.125))
    _lib.lib().s3r_set_tunable(12, 0)
    ours = gtime(lambda: attention_bf16(q, k, v, 0.125))
    ref = gtime(lambda: F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=0.125))
    fl = 4.0 * B * H * Nq * Nk * 64
    print(f"B={B:2d} H={H} Nq={Nq:4d} Nk={Nk:4d}: two-pass {two:6.1f} one-pass {one:6.1f} auto {ours:6.1f} us ({fl/ours/1e6:6.1f} TF) | SDPA {ref:6.1f} us", flush=True)
