"""GPU numerics: tcgen05 attention vs a plain PyTorch fp32 reference (softmax(q k^T * scale) v on the same
bf16-rounded inputs).  Tolerance: P is rounded to bf16 before P V (2^-8 relative per probability) and the output is
bf16: |err| <= 2e-2 * max|ref|."""
import pytest

pytestmark = pytest.mark.gpu


def ref_attention(q, k, v, scale):
    import torch
    qf, kf, vf = q.float().transpose(1, 2), k.float().transpose(1, 2), v.float().transpose(1, 2)
    p = torch.softmax(qf @ kf.transpose(-1, -2) * scale, dim=-1)
    return (p @ vf).transpose(1, 2)


@pytest.mark.parametrize("onepass", [0, 1, 2])
@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 16, 257, 257), (1, 12, 257, 771), (1, 12, 514, 256), (1, 12, 1028, 1028),
                                       (3, 2, 1, 5), (1, 1, 128, 64), (2, 3, 130, 65), (1, 16, 256, 256), (1, 2, 70, 33),
                                       (1, 2, 40, 97)])
def test_attention_matches_fp32_reference(B, H, Nq, Nk, onepass):
    """onepass = 1 / 2 force the single-pass kernel (two independent softmax streams per row, lazy reference maximum, O
    rows rescaled in TMEM) / the two-pass kernel (exact row maximum first); 0 = the shape-based dispatch."""
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.ops import attention_bf16
    torch.manual_seed(B * 1000 + Nq + Nk)
    q = torch.randn(B, Nq, H, 64, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Nk, H, 64, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Nk, H, 64, device="cuda").to(torch.bfloat16)
    try:
        _lib.check(_lib.lib().s3r_set_tunable(12, onepass))
        o = attention_bf16(q, k, v, 0.125)
        torch.cuda.synchronize()
    finally:
        _lib.lib().s3r_set_tunable(12, 0)
    expect = ref_attention(q, k, v, 0.125)
    err = (o.float() - expect).abs().max().item()
    assert o.shape == (B, Nq, H, 64)
    assert err <= 2e-2 * expect.abs().max().item(), err


def test_attention_on_packed_qkv_views_like_the_vit_block():
    """q, k, v are strided slices of the packed qkv projection [B,N,3,H,64] (croco/blocks.py:97-106)."""
    import torch
    from styl3r_b200.ops import memory_efficient_attention
    torch.manual_seed(0)
    qkv = (torch.randn(2, 257, 3, 16, 64, device="cuda") * 2).to(torch.bfloat16)
    q, k, v = qkv[:, :, 0], qkv[:, :, 1], qkv[:, :, 2]
    with torch.no_grad():
        o = memory_efficient_attention(q, k, v, scale=0.125)
    expect = ref_attention(q, k, v, 0.125)
    assert (o.float() - expect).abs().max() <= 2e-2 * expect.abs().max()
    # large logits: exact row maxima keep exp2 in range
    with torch.no_grad():
        o2 = memory_efficient_attention(q * 8, k * 8, v, scale=0.125)
    e2 = ref_attention(q * 8, k * 8, v, 0.125)
    assert torch.isfinite(o2.float()).all() and (o2.float() - e2).abs().max() <= 3e-2 * e2.abs().max()


@pytest.mark.parametrize("Nk", [257, 1028])
def test_one_pass_attention_rescales_when_the_running_maximum_moves(Nk):
    """Keys are ordered so that the row maximum keeps growing by far more than the lazy threshold (2^6) from tile to tile:
    every tile triggers the TMEM rescale of the O rows; large logits (|S * scale| up to ~100) must stay finite."""
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.ops import attention_bf16
    torch.manual_seed(Nk)
    B, H, Nq = 2, 3, 200
    q = torch.randn(B, Nq, H, 64, device="cuda").abs().to(torch.bfloat16)
    ramp = torch.linspace(0.05, 4.0, Nk, device="cuda").view(1, Nk, 1, 1)
    k = (torch.randn(B, Nk, H, 64, device="cuda").abs() * ramp).to(torch.bfloat16)   # q.k grows with the key index
    v = torch.randn(B, Nk, H, 64, device="cuda").to(torch.bfloat16)
    try:
        _lib.check(_lib.lib().s3r_set_tunable(12, 1))
        o = attention_bf16(q, k, v, 0.5)
        torch.cuda.synchronize()
    finally:
        _lib.lib().s3r_set_tunable(12, 0)
    expect = ref_attention(q, k, v, 0.5)
    assert torch.isfinite(o.float()).all()
    assert (o.float() - expect).abs().max() <= 2e-2 * expect.abs().max()
