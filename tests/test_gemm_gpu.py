"""GPU numerics: tcgen05 bf16 GEMM + fused epilogue vs a plain PyTorch fp32 reference of the same op on the same
bf16-rounded operands.  Tolerance: fp32 accumulation on both sides, so the difference is the final bf16 rounding of
the output (2^-8 relative) plus summation order: |err| <= 1e-2 * max|ref| for bf16 output, 2e-3 for fp32 output."""
import pytest

pytestmark = pytest.mark.gpu


def ref(x, w, b, r, gelu):
    import torch
    y = x.float() @ w.float().t()
    if b is not None:
        y = y + b.float()
    if gelu:
        y = torch.nn.functional.gelu(y)
    if r is not None:
        y = y + r.float()
    return y


@pytest.mark.parametrize("M,N,K", [(514, 3072, 1024), (514, 1024, 4096), (128, 128, 64), (257, 768, 768), (1, 64, 72),
                                   (4112, 4096, 1024), (771, 2304, 768), (300, 200, 136)])
@pytest.mark.parametrize("epi", ["none", "bias", "bias_gelu", "bias_res", "bias_gelu_res_f32"])
@pytest.mark.parametrize("direct", [0, 1, 2])
def test_gemm_matches_fp32_reference(M, N, K, epi, direct):
    """direct = 1 / 2 force / forbid the register epilogue of the one-tile kernel (S3R_TUNE_GEMM_DIRECT), 0 = dispatch by
    grid size."""
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.gemm import linear
    _lib.check(_lib.lib().s3r_set_tunable(13, direct))
    torch.manual_seed(M * 7 + N * 3 + K)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16) if "bias" in epi else None
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16) if "res" in epi else None
    out_dtype = torch.float32 if "f32" in epi else torch.bfloat16
    try:
        y = linear(x, w, b, r, gelu="gelu" in epi, out_dtype=out_dtype)
        torch.cuda.synchronize()
    finally:
        _lib.lib().s3r_set_tunable(13, 0)
    expect = ref(x, w, b, r, "gelu" in epi)
    err = (y.float() - expect).abs().max().item()
    tol = (2e-3 if out_dtype == torch.float32 else 1e-2) * expect.abs().max().item()
    assert y.shape == (M, N) and y.dtype == out_dtype
    assert err <= tol, f"max err {err:.4e} > {tol:.4e}"


def test_gemm_batched_leading_dims_and_strided_rows():
    import torch
    from styl3r_b200.gemm import linear
    x = torch.randn(2, 257, 2048, device="cuda").to(torch.bfloat16)[..., :1024]  # row pitch 2048
    w = (torch.randn(768, 1024, device="cuda") / 32).to(torch.bfloat16)
    y = linear(x, w)
    expect = x.float() @ w.float().t()
    assert y.shape == (2, 257, 768)
    assert (y.float() - expect).abs().max() <= 1e-2 * expect.abs().max()


@pytest.mark.parametrize("M,H,pair", [(514, 16, 0), (257, 12, 0), (771, 12, 0), (514, 16, 5), (771, 12, 5), (2570, 16, 0), (514, 16, -1), (257, 12, -2)])
def test_gemm_with_fused_rope_epilogue_matches_gemm_then_rope_oracle(M, H, pair):
    """qkv projection with RoPE-2D on the q and k thirds fused in the epilogue == fp32 GEMM followed by the RoPE
    oracle (oracle/rope_oracle.c restating curope.cpp:11-47).  pair = 5: the persistent CTA-pair kernel's register
    epilogue (rotation pairs inside one thread); M = 2570 takes that kernel through the automatic dispatch."""
    import numpy as np
    import torch
    from oracle import raster_oracle as ro
    from styl3r_b200 import _lib
    from styl3r_b200.gemm import linear
    if pair < 0:   # -1 / -2: one-tile kernel with the register epilogue forced / forbidden
        _lib.check(_lib.lib().s3r_set_tunable(13, -pair))
    else:
        _lib.check(_lib.lib().s3r_set_tunable(11, pair))
    torch.manual_seed(M)
    C_ = H * 64
    x = (torch.randn(M, C_, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(3 * C_, C_, device="cuda") * C_ ** -0.5).to(torch.bfloat16)
    b = (torch.randn(3 * C_, device="cuda") * 0.1).to(torch.bfloat16)
    pos = torch.randint(0, 17, (M, 2), device="cuda")
    try:
        y = linear(x, w, b, rope_pos=pos, rope_cols=2 * C_, rope_base=100.0)
        torch.cuda.synchronize()
    finally:
        _lib.lib().s3r_set_tunable(11, 0)
        _lib.lib().s3r_set_tunable(13, 0)
    ref = (x.float() @ w.float().t() + b.float()).cpu().numpy().reshape(1, M, 3, H, 64)
    expect = ref.copy()
    for part in (0, 1):  # q and k thirds
        expect[:, :, part] = ro.rope2d(np.ascontiguousarray(ref[:, :, part]), pos.cpu().numpy()[None], 100.0, 1.0)
    err = np.abs(y.float().cpu().numpy().reshape(1, M, 3, H, 64) - expect).max()
    assert err <= 1e-2 * np.abs(expect).max(), err
    # the v third is untouched by the rotation
    np.testing.assert_allclose(y.float().cpu().numpy().reshape(1, M, 3, H, 64)[:, :, 2], ref[:, :, 2],
                               atol=1e-2 * np.abs(ref).max())


@pytest.mark.parametrize("M,N,K", [(257, 768, 768), (514, 1024, 4096), (130, 200, 520)])
def test_gemm_split_k_path_is_deterministic_and_correct(M, N, K):
    import torch
    from styl3r_b200.gemm import linear
    torch.manual_seed(K)
    x = (torch.randn(M, K, device="cuda") * 0.5).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * K ** -0.5).to(torch.bfloat16)
    b = (torch.randn(N, device="cuda") * 0.1).to(torch.bfloat16)
    r = torch.randn(M, N, device="cuda").to(torch.bfloat16)
    y1 = linear(x, w, b, r, gelu=True, split_k=True)
    y2 = linear(x, w, b, r, gelu=True, split_k=True)   # counters self-reset; fixed summation order
    y0 = linear(x, w, b, r, gelu=True, split_k=False)
    torch.cuda.synchronize()
    expect = ref(x, w, b, r, True)
    assert torch.equal(y1, y2)
    assert (y1.float() - expect).abs().max() <= 1e-2 * expect.abs().max()
    assert (y0.float() - expect).abs().max() <= 1e-2 * expect.abs().max()


@pytest.mark.parametrize("code", [12, 14, 21, 22, 24])
@pytest.mark.parametrize("M,N,K,big", [(514, 3072, 1024, 0), (257, 768, 768, 0), (4112, 1024, 1024, 0), (4112, 3072, 1024, 1)])
def test_gemm_cluster_multicast_variants_match_reference(code, M, N, K, big):
    """Thread-block clusters (CM x CN) with TMA multicast of the shared operand tile and cluster-wide slot release
    (S3R_TUNE_GEMM_CLUSTER = CM*10 + CN; off by default - measured not to pay off on B200): same
    result as the plain kernel, including ragged M (padded cluster rows) and the 128x256 tile."""
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.gemm import linear
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K + code)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16)
    base = linear(x, w, b, gelu=True)
    try:
        _lib.check(L.s3r_set_tunable(2, code))
        _lib.check(L.s3r_set_tunable(3, big))
        y = linear(x, w, b, gelu=True)
        torch.cuda.synchronize()
    finally:
        L.s3r_set_tunable(2, 0)
        L.s3r_set_tunable(3, 0)
    assert torch.equal(y, base) or (y.float() - base.float()).abs().max().item() <= 2 ** -6


@pytest.mark.parametrize("mode", [1, 3, 4, 5])
@pytest.mark.parametrize("M,N,K", [(514, 3072, 1024), (257, 768, 768), (4112, 1024, 4096), (1300, 512, 200), (129, 256, 64), (300, 264, 72)])
@pytest.mark.parametrize("epi", ["bias", "bias_gelu", "bias_res", "none_f32"])
def test_gemm_cta_pair_variants_match_reference(mode, M, N, K, epi):
    """CTA pairs (tcgen05 cta_group::2, S3R_TUNE_GEMM_PAIR = 1: 256x128 pair tiles, 3: 256x256; 4 / 5: the PERSISTENT pair
    kernel with the double-buffered TMEM accumulator and the register epilogue): one 256-row MMA over the two SMs of a
    TPC, each CTA stages its own A rows and half of the B tile; includes ragged M (odd number of 128-row tiles: the
    padding CTA of the last pair loads zeros and stores nothing), a K tail and an N that is not a multiple of the tile."""
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.gemm import linear
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(M + N + K + mode)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    b = torch.randn(N, device="cuda", generator=g).to(torch.bfloat16) if "bias" in epi else None
    r = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16) if "res" in epi else None
    out_dtype = torch.float32 if "f32" in epi else torch.bfloat16
    try:
        _lib.check(L.s3r_set_tunable(11, mode))
        y = linear(x, w, b, r, gelu="gelu" in epi, out_dtype=out_dtype)
        torch.cuda.synchronize()
    finally:
        L.s3r_set_tunable(11, 0)
    expect = ref(x, w, b, r, "gelu" in epi)
    tol = (2e-3 if out_dtype == torch.float32 else 1e-2) * expect.abs().max().item()
    assert (y.float() - expect).abs().max().item() <= tol


def test_conv_cta_pair_matches_plain():
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(5)
    for (n, h, w, ci, co) in [(1, 128, 128, 256, 256), (3, 8, 8, 128, 128), (1, 256, 256, 64, 256), (2, 32, 32, 256, 256)]:
        x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(torch.bfloat16)
        wp = prep_conv_weight((torch.randn(co, ci, 3, 3, device="cuda", generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16))
        bias = torch.randn(co, device="cuda", generator=g).to(torch.bfloat16)
        try:
            _lib.check(L.s3r_set_tunable(11, 2))
            base = conv2d_nhwc(x, wp, (3, 3), bias=bias, relu=True)
            for mode in (1, 3, 4, 5):
                _lib.check(L.s3r_set_tunable(11, mode))
                y = conv2d_nhwc(x, wp, (3, 3), bias=bias, relu=True)
                torch.cuda.synchronize()
                assert torch.equal(y, base) or (y.float() - base.float()).abs().max().item() <= 2 ** -6
        finally:
            L.s3r_set_tunable(11, 0)


def test_conv_cluster_multicast_matches_plain():
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.conv import conv2d_nhwc, prep_conv_weight
    L = _lib.lib()
    g = torch.Generator(device="cuda").manual_seed(3)
    for (n, h, w, ci, co) in [(1, 128, 128, 256, 256), (3, 8, 8, 128, 128), (1, 256, 256, 64, 256)]:
        x = torch.randn(n, h, w, ci, device="cuda", generator=g).to(torch.bfloat16)
        wp = prep_conv_weight((torch.randn(co, ci, 3, 3, device="cuda", generator=g) / (9 * ci) ** 0.5).to(torch.bfloat16))
        base = conv2d_nhwc(x, wp, (3, 3), relu=True)
        try:
            _lib.check(L.s3r_set_tunable(4, 1))
            y = conv2d_nhwc(x, wp, (3, 3), relu=True)
            torch.cuda.synchronize()
        finally:
            L.s3r_set_tunable(4, 0)
        # (the plain call may take the cluster split-K path on the small shape: another summation order)
        assert torch.equal(y, base) or (y.float() - base.float()).abs().max().item() <= 2 ** -6


@pytest.mark.parametrize("M,C,strided", [(514, 1024, False), (257, 768, False), (1, 256, False), (1028, 768, True), (33, 2048, False)])
def test_layernorm_matches_fp32_reference(M, C, strided):
    """s3r_layernorm_bf16 (one warp per row, fp32 statistics) vs torch fp32 LayerNorm on the same bf16 values:
    |err| <= one bf16 rounding of the output (2^-8 relative)."""
    import torch
    from styl3r_b200.encoder.vit import _ln
    g = torch.Generator(device="cuda").manual_seed(M + C)
    norm = torch.nn.LayerNorm(C, eps=1e-6).cuda()
    with torch.no_grad():
        norm.weight.copy_(1 + 0.1 * torch.randn(C, device="cuda", generator=g))
        norm.bias.copy_(0.1 * torch.randn(C, device="cuda", generator=g))
    norm = norm.to(torch.bfloat16)
    x = (torch.randn(M, 2 * C if strided else C, device="cuda", generator=g) * 3 + 0.5).to(torch.bfloat16)
    if strided:
        x = x[:, :C]          # row pitch 2*C: the kernel takes the pitch, no copy
    with torch.no_grad():
        y = _ln(norm, x)
        ref = torch.nn.functional.layer_norm(x.float(), (C,), norm.weight.float(), norm.bias.float(), 1e-6)
    torch.cuda.synchronize()
    assert y.dtype == torch.bfloat16 and y.shape == x.shape
    assert (y.float() - ref).abs().max().item() <= 2.0 ** -8 * max(1.0, ref.abs().max().item())


@pytest.mark.parametrize("M,N,K", [(514, 1024, 3072), (5140, 768, 768), (257, 3072, 1024), (300, 136, 200), (64, 64, 72)])
def test_backward_gemms_mn_major_operands(M, N, K):
    """dgrad / wgrad of nn.Linear on the tcgen05 GEMM with MN-major operands (no transposed copies) vs fp32 torch.
    Tolerance: bf16 operands, fp32 accumulation -> |err| <= 2e-2 * sqrt(K) * rms(a) * rms(b) (bf16 output rounding
    included), the same bar as the forward GEMM tests."""
    import torch
    from styl3r_b200.gemm import gemm_majors, linear_dgrad, linear_wgrad
    g = torch.Generator(device="cuda").manual_seed(M + N + K)
    dy = torch.randn(M, N, device="cuda", generator=g).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda", generator=g) / K ** 0.5).to(torch.bfloat16)
    x = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    pre = torch.randn(M, K, device="cuda", generator=g).to(torch.bfloat16)
    dx = linear_dgrad(dy, w)
    ref = dy.float() @ w.float()
    assert (dx.float() - ref).abs().max() <= 2e-2 * ref.abs().max() + 1e-3
    dxg = linear_dgrad(dy, w, pre_gelu=pre)
    h = pre.float()
    gp = 0.5 * (1 + torch.erf(h / 2 ** 0.5)) + h * torch.exp(-0.5 * h * h) / (2 * torch.pi) ** 0.5
    assert (dxg.float() - ref * gp).abs().max() <= 2e-2 * (ref * gp).abs().max() + 1e-3
    dw = linear_wgrad(dy, x)
    refw = dy.float().t() @ x.float()
    assert dw.dtype == torch.float32 and (dw - refw).abs().max() <= 2e-3 * refw.abs().max() + 1e-3
    # A MN-major alone: C = A^T-stored . B^T
    at = torch.zeros(K, (M + 7) // 8 * 8, device="cuda", dtype=torch.bfloat16)[:, :M]  # [K, M], 16-byte row pitch
    at.copy_(x.t())
    c = gemm_majors(at, w, M, N, K, True, False, out_dtype=torch.float32)
    refc = x.float() @ w.float().t()
    assert (c - refc).abs().max() <= 2e-3 * refc.abs().max() + 1e-3
