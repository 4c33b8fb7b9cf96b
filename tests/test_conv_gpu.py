"""GPU numerics: tcgen05 implicit-GEMM convolution, bilinear x2 upsampling and the bf16 NHWC DPT head path vs plain
PyTorch fp32 references of the same ops (heads/dpt_block.py, dpt_head.py, dpt_gs_head.py semantics).

Tolerances: operands are rounded to bf16 on both sides (the fp32 reference gets the bf16-rounded values), the kernel
accumulates in fp32, the output is rounded to bf16 once -> |err| <= 2^-8 relative to the output scale + accumulation
order noise; whole heads (~25 chained bf16 layers) are compared on the mean error."""
import pytest

pytestmark = pytest.mark.gpu


def _rt(t):  # bf16 round trip
    import torch
    return t.to(torch.bfloat16).float()


@pytest.mark.parametrize("n,h,w,cin,cout,k", [
    (1, 256, 256, 256, 256, 3),   # gs heads' dominant conv (M = 65536, two 128-pixel tiles per row)
    (2, 128, 128, 256, 128, 3),   # pts3d head.0 (one tile per row)
    (1, 64, 64, 96, 256, 3),      # layer1_rn: Cin = 96 -> second 64-channel block is half out of bounds
    (3, 8, 8, 768, 256, 3),       # layer4_rn at 8x8: tile spans two images, last tile ragged
    (1, 16, 16, 384, 256, 3),     # layer3_rn
    (2, 32, 32, 192, 256, 3),
    (1, 32, 128, 64, 64, 5),      # 5x5, BN = 64 path, non-square
    (1, 16, 16, 128, 256, 1),     # 1x1 through the conv entry
])
@pytest.mark.parametrize("epi", ["plain", "bias_relu", "bias_res", "f32"])
def test_conv2d_matches_fp32_reference(n, h, w, cin, cout, k, epi):
    import torch
    import torch.nn.functional as F
    from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
    g = torch.Generator(device="cuda").manual_seed(n * 1000 + h + cin + k)
    x = _rt(torch.randn(n, cin, h, w, device="cuda", generator=g))
    wt = _rt(torch.randn(cout, cin, k, k, device="cuda", generator=g) / (cin * k * k) ** 0.5)
    b = _rt(torch.randn(cout, device="cuda", generator=g))
    r = _rt(torch.randn(n, cout, h, w, device="cuda", generator=g))
    # float64 reference: cuDNN's fp32 algorithm choice (FFT / Winograd for some filter sizes) is itself only good to
    # ~1e-3, which would mask or fake errors at the fp32-output tolerance
    ref = F.conv2d(x.double(), wt.double(), None, 1, k // 2).float()
    kw = {}
    if epi in ("bias_relu", "bias_res"):
        ref = ref + b.view(1, -1, 1, 1)
        kw["bias"] = b.to(torch.bfloat16)
    if epi == "bias_relu":
        ref = torch.relu(ref)
        kw["relu"] = True
    if epi == "bias_res":
        ref = ref + r
        kw["residual"] = r.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    if epi == "f32":
        kw["out_dtype"] = torch.float32
    xn = x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    y = conv2d_nhwc(xn, prep_conv_weight(wt), (k, k), **kw)
    torch.cuda.synchronize()
    assert y.shape == (n, h, w, cout)
    err = (y.float().permute(0, 3, 1, 2) - ref).abs()
    tol = 2e-4 if epi == "f32" else 1.2e-2  # bf16 output rounding: 2^-8 * |y| (|y| up to ~5 with the residual)
    scale = ref.abs().max().item()
    assert err.max().item() <= tol * max(scale, 1.0), f"max err {err.max().item():.3e} (scale {scale:.2f})"
    assert err.mean().item() <= tol * 0.3


def test_conv2d_rejects_unsupported_shapes():
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
    w = prep_conv_weight(torch.randn(64, 64, 3, 3, device="cuda"))
    with pytest.raises(_lib.S3RError):
        conv2d_nhwc(torch.zeros(1, 24, 24, 64, device="cuda", dtype=torch.bfloat16), w, (3, 3))  # 128 % 24 != 0
    with pytest.raises(_lib.S3RError):
        conv2d_nhwc(torch.zeros(1, 16, 16, 64, dtype=torch.bfloat16), w, (3, 3))  # CPU tensor: no fallback
    y = conv2d_nhwc(torch.zeros(0, 16, 16, 64, device="cuda", dtype=torch.bfloat16), w, (3, 3))  # empty batch
    assert y.shape == (0, 16, 16, 64)


@pytest.mark.parametrize("n,h,w,c,with_add", [(1, 8, 8, 256, False), (2, 64, 64, 256, False), (1, 128, 128, 256, True),
                                              (1, 128, 128, 128, False), (3, 5, 7, 8, True)])
def test_upsample2x_matches_interpolate(n, h, w, c, with_add):
    import torch
    import torch.nn.functional as F
    from styl3r_b200.conv import upsample2x_nhwc
    g = torch.Generator(device="cuda").manual_seed(h * w + c)
    x = _rt(torch.randn(n, c, h, w, device="cuda", generator=g))
    ref = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
    add = None
    if with_add:
        a = _rt(torch.randn(n, c, 2 * h, 2 * w, device="cuda", generator=g))
        ref = _rt(ref) + a
        add = a.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16)
    y = upsample2x_nhwc(x.permute(0, 2, 3, 1).contiguous().to(torch.bfloat16), add)
    torch.cuda.synchronize()
    err = (y.float().permute(0, 3, 1, 2) - ref).abs().max().item()
    assert err <= 2.0 ** -7 * max(1.0, ref.abs().max().item()), err  # one (two with add) bf16 roundings


@pytest.mark.parametrize("kind,out_ch", [("pts3d", 3), ("gs_params", 8), ("gs_sh", 3)])
def test_dpt_head_nhwc_path_matches_fp32_module(kind, out_ch):
    """The whole head: bf16 NHWC tcgen05 path vs the fp32 nn.Module (the reference's op sequence) on the same
    weights and tokens."""
    import torch
    from styl3r_b200.encoder.dpt import PixelwiseDPT
    torch.manual_seed(3)
    torch.backends.cudnn.allow_tf32 = False
    head = PixelwiseDPT(kind, out_ch).cuda().eval()
    B = 2
    toks = [None] * 13
    for hook, c in zip((0, 6, 9, 12), (1024, 768, 768, 768)):
        toks[hook] = torch.randn(B, 256, c, device="cuda")
    img = torch.rand(B, 3, 256, 256, device="cuda") * 2 - 1
    with torch.no_grad():
        ref = head([None if t is None else _rt(t) for t in toks], (256, 256), img)      # [B, C, 256, 256]
        out = head.forward_nhwc([None if t is None else t.to(torch.bfloat16) for t in toks], (256, 256), img)
    torch.cuda.synchronize()
    assert out.dtype == torch.float32 and out.shape == (B * 65536, 8)
    got = out.view(B, 256, 256, 8)[..., :out_ch].permute(0, 3, 1, 2)
    err = (got - ref).abs()
    scale = ref.std().item()
    assert err.mean().item() <= 1.5e-2 * scale and err.max().item() <= 0.25 * scale + 1e-3, \
        f"{kind}: mean {err.mean().item():.3e} max {err.max().item():.3e} scale {scale:.3e}"
    if out_ch < 8:
        assert out.view(-1, 8)[:, out_ch:].abs().max().item() == 0.0  # padded output rows stay zero


def test_gaussian_adapter_nhwc_equals_planar():
    import ctypes as C
    import torch
    from styl3r_b200 import _lib
    torch.manual_seed(0)
    B, HW, G = 2, 4096, 2 * 4096
    L = _lib.lib()
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    st = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    pts, prm, app = (torch.randn(B, c, HW, device="cuda") for c in (3, 8, 3))
    mask = torch.ones(1, device="cuda")
    outs = []
    for nhwc in (False, True):
        means, cov = torch.zeros(B, G, 3, device="cuda"), torch.zeros(B, G, 3, 3, device="cuda")
        harm, opac = torch.zeros(B, G, 3, 1, device="cuda"), torch.zeros(B, G, device="cuda")
        sc, rot = torch.zeros(B, G, 3, device="cuda"), torch.zeros(B, G, 4, device="cuda")
        if nhwc:
            rows = lambda t: torch.nn.functional.pad(t.permute(0, 2, 1), (0, 8 - t.shape[1])).reshape(B * HW, 8).contiguous()
            a, b_, c_ = rows(pts), rows(prm), rows(app)
            _lib.check(L.s3r_gaussian_adapter_nhwc(p(a), p(b_), p(c_), 8, 8, 8, p(mask), B, HW, 1, 1, G, 1.0, p(means),
                                                   p(cov), p(harm), p(opac), p(sc), p(rot), st))
        else:
            _lib.check(L.s3r_gaussian_adapter(p(pts), p(prm), p(app), p(mask), B, HW, 1, 1, G, 1.0, p(means), p(cov),
                                              p(harm), p(opac), p(sc), p(rot), st))
        torch.cuda.synchronize()
        outs.append((means, cov, harm, opac, sc, rot))
    for a, b_ in zip(*outs):
        assert torch.equal(a, b_)
