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


@pytest.mark.parametrize("B,H,Nq,Nk", [(2, 16, 257, 257), (1, 12, 257, 771), (1, 12, 514, 256), (1, 12, 1028, 1028),
                                       (3, 2, 1, 5), (1, 1, 128, 64), (2, 3, 130, 65), (1, 16, 256, 256)])
def test_attention_matches_fp32_reference(B, H, Nq, Nk):
    import torch
    from styl3r_b200.ops import attention_bf16
    torch.manual_seed(B * 1000 + Nq + Nk)
    q = torch.randn(B, Nq, H, 64, device="cuda").to(torch.bfloat16)
    k = torch.randn(B, Nk, H, 64, device="cuda").to(torch.bfloat16)
    v = torch.randn(B, Nk, H, 64, device="cuda").to(torch.bfloat16)
    o = attention_bf16(q, k, v, 0.125)
    torch.cuda.synchronize()
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
