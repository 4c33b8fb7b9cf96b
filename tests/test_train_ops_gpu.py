"""GPU: the bf16 training layout of the ViT trunks (styl3r_b200/encoder/train_ops.py - forward AND backward on our
kernels) against torch autograd over the reference's fp32 ops on the same modules.

Tolerances: bf16 operands / activations with fp32 accumulation vs fp32 end to end.  Per tensor: relative L2 error
<= 2e-2 (outputs) / 4e-2 (gradients: two chained bf16 GEMMs and bf16 intermediate gradients) and cosine similarity
>= 0.999; the attention backward alone is held to 1.5e-2 relative L2 against SDPA autograd on the same bf16 inputs."""
import pytest

pytestmark = pytest.mark.gpu


def rel(a, b):
    a, b = a.float(), b.float()
    return ((a - b).norm() / (b.norm() + 1e-12)).item()


def cos(a, b):
    import torch
    return torch.nn.functional.cosine_similarity(a.float().flatten(), b.float().flatten(), dim=0).item()


@pytest.fixture(autouse=True)
def _restore_mode():
    from styl3r_b200.encoder import vit
    yield
    vit.TRAIN_BF16 = False


def test_layernorm_backward_matches_torch():
    import torch
    from styl3r_b200.encoder import train_ops as T
    g = torch.Generator(device="cuda").manual_seed(0)
    for M, C_ in ((514, 1024), (1030, 768), (3, 256)):
        ln = torch.nn.LayerNorm(C_, eps=1e-6).cuda()
        with torch.no_grad():
            ln.weight.copy_(1 + 0.1 * torch.randn(C_, device="cuda", generator=g))
            ln.bias.copy_(0.1 * torch.randn(C_, device="cuda", generator=g))
        x = torch.randn(M, C_, device="cuda", generator=g).to(torch.bfloat16)
        dy = torch.randn(M, C_, device="cuda", generator=g).to(torch.bfloat16)
        xa = x.clone().requires_grad_()
        ya = T.layer_norm(xa, ln)
        ya.backward(dy)
        gw, gb = ln.weight.grad.clone(), ln.bias.grad.clone()
        ln.weight.grad = ln.bias.grad = None
        xb = x.float().requires_grad_()
        yb = ln(xb)
        yb.backward(dy.float())
        assert rel(ya, yb) <= 1e-2 and rel(xa.grad, xb.grad) <= 1e-2, (M, C_, rel(ya, yb), rel(xa.grad, xb.grad))
        assert rel(gw, ln.weight.grad) <= 1e-2 and rel(gb, ln.bias.grad) <= 1e-2, (M, C_, rel(gw, ln.weight.grad))
        ln.weight.grad = ln.bias.grad = None


@pytest.mark.parametrize("B,Nq,Nk,H", [(2, 257, 257, 16), (3, 514, 256, 12), (1, 100, 771, 12), (2, 64, 8, 4)])
def test_attention_backward_matches_sdpa_autograd(B, Nq, Nk, H):
    import torch
    from styl3r_b200.attention_bwd import attention_backward
    g = torch.Generator(device="cuda").manual_seed(Nq + Nk)
    mk = lambda n: torch.randn(B, n, H, 64, device="cuda", generator=g).to(torch.bfloat16)
    q, k, v, do = mk(Nq), mk(Nk), mk(Nk), mk(Nq)
    scale = 0.125
    qf, kf, vf = (t.float().requires_grad_() for t in (q, k, v))
    o = torch.nn.functional.scaled_dot_product_attention(qf.transpose(1, 2), kf.transpose(1, 2), vf.transpose(1, 2),
                                                         scale=scale).transpose(1, 2)
    o.backward(do.float())
    dq, dk, dv = attention_backward(q, k, v, o.detach().to(torch.bfloat16), do, scale)
    for name, mine, ref in (("dq", dq, qf.grad), ("dk", dk, kf.grad), ("dv", dv, vf.grad)):
        assert mine.shape == ref.shape
        assert rel(mine, ref) <= 1.5e-2 and cos(mine, ref) >= 0.9995, (name, rel(mine, ref), cos(mine, ref))


def _compare_module(make, inputs, train_call):
    """Run `train_call(module, *inputs)` once on the fp32 torch path and once on the bf16 training layout."""
    import torch
    from styl3r_b200.encoder import vit
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.manual_seed(0)
    mod = make().cuda()
    for p_ in mod.parameters():
        if p_.dim() == 1:
            with torch.no_grad():
                p_.add_(0.05 * torch.randn_like(p_))
    res = {}
    for mode in (False, True):
        vit.TRAIN_BF16 = mode
        ins = [t.clone().requires_grad_() if t.is_floating_point() else t for t in inputs]
        mod.zero_grad(set_to_none=True)
        out = train_call(mod, *ins)
        w = torch.randn(out.shape, device="cuda", generator=torch.Generator(device="cuda").manual_seed(5))
        (out.float() * w).sum().backward()
        res[mode] = (out.detach().float(), [t.grad.detach().float() for t in ins if t.is_floating_point()],
                     {n: p_.grad.detach().float().clone() for n, p_ in mod.named_parameters() if p_.grad is not None})
    vit.TRAIN_BF16 = False
    ref, mine = res[False], res[True]
    assert rel(mine[0], ref[0]) <= 2e-2, ("output", rel(mine[0], ref[0]))
    for a, b in zip(mine[1], ref[1]):
        assert rel(a, b) <= 4e-2 and cos(a, b) >= 0.999, ("input grad", rel(a, b), cos(a, b))
    assert set(mine[2]) == set(ref[2]) and len(ref[2]) > 0
    for n in ref[2]:
        assert mine[2][n].dtype == torch.float32
        assert rel(mine[2][n], ref[2][n]) <= 4e-2 and cos(mine[2][n], ref[2][n]) >= 0.999, (n, rel(mine[2][n], ref[2][n]))


def _pos(B, gh, gw, extra=0):
    import torch
    ys, xs = torch.meshgrid(torch.arange(gh), torch.arange(gw), indexing="ij")
    pos = torch.stack((ys.reshape(-1), xs.reshape(-1)), -1)
    if extra:
        pos = torch.cat((pos, torch.tensor([[gh, 0]])), 0)
    return pos[None].expand(B, -1, -1).contiguous().cuda()


def test_block_forward_backward_on_our_kernels_matches_fp32_autograd():
    import torch
    from styl3r_b200.encoder.vit import Block
    x = torch.randn(2, 257, 1024, device="cuda", generator=torch.Generator(device="cuda").manual_seed(1))
    _compare_module(lambda: Block(1024, 16, 100.0), [x, _pos(2, 16, 16, 1)], lambda m, a, p: m(a, p))


def test_decoder_block_forward_backward_on_our_kernels_matches_fp32_autograd():
    import torch
    from styl3r_b200.encoder.vit import DecoderBlock
    g = torch.Generator(device="cuda").manual_seed(2)
    x = torch.randn(2, 514, 768, device="cuda", generator=g)
    y = torch.randn(2, 256, 768, device="cuda", generator=g)
    xpos = torch.cat((_pos(2, 16, 16, 1), _pos(2, 16, 16, 1)), 1)
    _compare_module(lambda: DecoderBlock(768, 12, 100.0), [x, y, xpos, _pos(2, 16, 16)], lambda m, a, b, p, q: m(a, b, p, q))


def test_patch_embed_training_path():
    import torch
    from styl3r_b200.encoder.vit import PatchEmbed
    img = torch.randn(2, 3, 64, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(3))
    _compare_module(lambda: PatchEmbed(16, 3, 1024), [img], lambda m, a: m(a)[0])


def test_frozen_layers_cost_no_weight_gradients():
    import torch
    from styl3r_b200.encoder import vit
    from styl3r_b200.encoder.vit import Block
    blk = Block(1024, 16, 100.0).cuda()
    for p_ in blk.parameters():
        p_.requires_grad = False
    vit.TRAIN_BF16 = True
    x = torch.randn(1, 257, 1024, device="cuda").requires_grad_()
    blk(x, _pos(1, 16, 16, 1)).float().sum().backward()
    assert x.grad is not None and all(p_.grad is None for p_ in blk.parameters())
