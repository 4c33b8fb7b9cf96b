"""Stress run (GPU): many replays of the graphed cfg3 / cfg2 forward with stream branches, fresh graphs several times -
looks for intermittent hangs / nondeterminism of the cluster / persistent kernels under concurrency."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, GraphedEncoder, get_encoder
from tests.encoder_weights import make_inputs
enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval().to_inference(torch.bfloat16)
for rnd in range(3):
    for (b, v) in ((4, 4), (1, 2), (2, 3)):
        context, style = make_inputs(b, v, 256, seed=rnd, device="cuda")
        fast = GraphedEncoder(enc)
        t0 = time.time()
        ref = None
        for i in range(60):
            g = fast(context, style)
            if i % 20 == 0:
                torch.cuda.synchronize()
                cur = g.means.clone()
                assert torch.isfinite(cur).all()
                if ref is None: ref = cur
                else: assert torch.equal(ref, cur), "replays differ"
        torch.cuda.synchronize()
        print(f"round {rnd} b={b} v={v}: 60 replays {time.time()-t0:.2f} s ok", flush=True)
print("stress ok")
