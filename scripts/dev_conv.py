"""Development aid (GPU): conv2d_nhwc correctness + timing per tile variant vs cuDNN, DPT head errors, encoder time with
cuDNN heads vs tcgen05 heads."""
import sys, time
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parents[1]))
import torch
import torch.nn.functional as F
from styl3r_b200 import _lib
from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight, upsample2x_nhwc

torch.backends.cudnn.allow_tf32 = False
L = _lib.lib()

def t_ms(fn, iters=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters

shapes = [(1, 256, 256, 256, 256, 3), (4, 256, 256, 256, 256, 3), (1, 256, 256, 128, 128, 3), (1, 128, 128, 256, 128, 3),
          (1, 64, 64, 256, 256, 3), (1, 64, 64, 96, 256, 3), (1, 32, 32, 256, 256, 3), (2, 8, 8, 768, 256, 3), (1, 16, 16, 256, 256, 3)]
for (n, h, w, ci, co, k) in shapes:
    x = torch.randn(n, ci, h, w, device="cuda").to(torch.bfloat16)
    wt = (torch.randn(co, ci, k, k, device="cuda") / (ci * k * k) ** 0.5).to(torch.bfloat16)
    ref = F.conv2d(x.float(), wt.float(), None, 1, k // 2)
    xn = x.permute(0, 2, 3, 1).contiguous()
    wp = prep_conv_weight(wt)
    flops = 2.0 * n * h * w * ci * co * k * k
    line = f"conv n={n} {h}x{w} {ci}->{co} k{k}:"
    for var in (0, 1, 2):
        if var and co % 256: continue
        _lib.check(L.s3r_set_tunable(1, var))
        y = conv2d_nhwc(xn, wp, (k, k))
        torch.cuda.synchronize()
        err = (y.float().permute(0, 3, 1, 2) - ref).abs().max().item()
        ms = t_ms(lambda: conv2d_nhwc(xn, wp, (k, k)))
        line += f"  v{var}: err {err:.2e} {ms*1000:.1f} us {flops/ms/1e9:.0f} TF/s |"
    xc = x.contiguous(memory_format=torch.channels_last); wc = wt.contiguous(memory_format=torch.channels_last)
    ms = t_ms(lambda: F.conv2d(xc, wc, None, 1, k // 2))
    line += f"  cudnn bf16 NHWC {ms*1000:.1f} us {flops/ms/1e9:.0f} TF/s"
    torch.backends.cudnn.allow_tf32 = True
    xf, wf = xc.float(), wc.float()
    ms = t_ms(lambda: F.conv2d(xf, wf, None, 1, k // 2))
    torch.backends.cudnn.allow_tf32 = False
    line += f" | cudnn tf32 NHWC {ms*1000:.1f} us"
    print(line, flush=True)
_lib.check(L.s3r_set_tunable(1, 0))

x = torch.randn(1, 128, 128, 256, device="cuda").to(torch.bfloat16)
print(f"upsample2x 128->256 x256ch: {t_ms(lambda: upsample2x_nhwc(x))*1000:.1f} us", flush=True)

# DPT heads
from styl3r_b200.encoder.dpt import PixelwiseDPT
for kind, oc in (("pts3d", 3), ("gs_params", 8), ("gs_sh", 3)):
    torch.manual_seed(3)
    head = PixelwiseDPT(kind, oc).cuda().eval()
    B = 2
    toks = [None] * 13
    for hook, c in zip((0, 6, 9, 12), (1024, 768, 768, 768)):
        toks[hook] = torch.randn(B, 256, c, device="cuda").to(torch.bfloat16)
    img = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
    with torch.no_grad():
        ref = head([None if t is None else t.float() for t in toks], (256, 256), img)
        out = head.forward_nhwc(toks, (256, 256), img)
        got = out.view(B, 256, 256, -1)[..., :oc].permute(0, 3, 1, 2)
        err = (got - ref).abs()
        print(f"head {kind}: mean err {err.mean().item():.3e} max {err.max().item():.3e} ref std {ref.std().item():.3e}", flush=True)
        headc = head.to(memory_format=torch.channels_last)
        torch.backends.cudnn.allow_tf32 = True
        tf = [None if t is None else t.float() for t in toks]
        ms_ref = t_ms(lambda: headc(tf, (256, 256), img), 5)
        torch.backends.cudnn.allow_tf32 = False
        for var in (0, 1, 2):
            _lib.check(L.s3r_set_tunable(1, var))
            ms = t_ms(lambda: head.forward_nhwc(toks, (256, 256), img), 5)
            print(f"   B={B} eager: cudnn tf32 channels_last {ms_ref:.2f} ms | tcgen05 v{var} {ms:.2f} ms", flush=True)
_lib.check(L.s3r_set_tunable(1, 0))

# encoder end to end
from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder, GraphedEncoder
from tests.encoder_weights import make_inputs
torch.backends.cuda.matmul.allow_tf32 = True; torch.backends.cudnn.allow_tf32 = True
enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True)); enc = enc.cuda().eval()
_lib.check(L.s3r_set_tunable(1, -1))
for heads, br in (("cudnn", False), ("tcgen05", False), ("cudnn", True), ("tcgen05", True)):
    enc.to_inference(torch.bfloat16, heads=heads, branches=br)
    for (b, v) in ((1, 2), (4, 4)):
        context, style = make_inputs(b, v, 256, seed=1, device="cuda")
        fast = GraphedEncoder(enc)
        out = fast(context, style); torch.cuda.synchronize()
        ms = t_ms(lambda: fast(context, style), 5)
        flops = {2: 1270.8e9, 4: 2437.2e9}[v] * b
        print(f"GRAPH encoder heads={heads} branches={br} b={b} v={v}: {ms:.2f} ms {flops/ms/1e9:.1f} TFLOP/s  means|mean| {out.means.abs().mean().item():.4f}", flush=True)
