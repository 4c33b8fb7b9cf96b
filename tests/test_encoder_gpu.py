"""GPU parity: our encoder (styl3r_b200.encoder) vs golden vectors produced by the REFERENCE encoder on CPU
(tests/golden/make_encoder_golden.py).  Both models get the same name-derived weights and seeded inputs.
Tolerance (fp32 on both sides, ~60 layers, different GEMM/conv reduction orders; `means` amplify head error through
expm1): max |err| <= 2e-3 * scale and mean |err| <= 2e-4 * scale, scale = the golden tensor's std."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = Path(__file__).parent / "golden"


def sample(t, n=4096):
    import torch
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long().to(f.device)
    return f[idx].float().cpu().numpy()


@pytest.fixture(scope="module")
def encoder():
    import torch
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
    from tests.encoder_weights import fill_named_weights
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False
    enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
    fill_named_weights(enc)
    return enc.cuda().eval()


@pytest.mark.parametrize("tag,b,v", [("b1v2", 1, 2), ("b1v3", 1, 3)])
def test_encoder_matches_reference_golden(encoder, tag, b, v):
    import torch
    from tests.encoder_weights import make_inputs
    g = np.load(GOLD / "encoder_golden.npz")
    context, style = make_inputs(b, v, 256, seed=1234, device="cuda")
    dump = {}
    with torch.no_grad():
        out = encoder(context, style, visualization_dump=dump)
    assert out.means.shape == (b, v * 65536, 3) and out.covariances.shape == (b, v * 65536, 3, 3)
    assert out.harmonics.shape == (b, v * 65536, 3, 1) and out.opacities.shape == (b, v * 65536)
    for name, t in [("means", out.means), ("covariances", out.covariances), ("harmonics", out.harmonics),
                    ("opacities", out.opacities), ("scales", dump["scales"]), ("rotations", dump["rotations"])]:
        ref, scale = g[f"{tag}_{name}"], float(g[f"{tag}_{name}_stats"][2])
        err = np.abs(sample(t) - ref)
        assert err.max() <= 2e-3 * scale and err.mean() <= 2e-4 * scale, \
            f"{tag} {name}: max {err.max():.3e} mean {err.mean():.3e} scale {scale:.3e}"


def test_encoder_feeds_decoder(encoder):
    """Encoder -> DecoderSplattingCUDA end to end (BASELINE cfg2 shapes), strict state-dict round trip."""
    import torch
    from styl3r_b200 import synthetic as syn
    from styl3r_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from tests.encoder_weights import make_inputs
    context, style = make_inputs(1, 2, 256, seed=7, device="cuda")
    with torch.no_grad():
        g = encoder(context, style)
        sc = syn.make_scene(seed=1, v=2, V=1, hw=256)
        t = lambda a: torch.as_tensor(a).cuda()[None]
        dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True)).cuda()
        out = dec(g, t(sc["extrinsics"]), t(sc["intrinsics"]), t(sc["near"]), t(sc["far"]), (256, 256))
    assert out.color.shape == (1, 1, 3, 256, 256) and torch.isfinite(out.color).all()
    sd = encoder.state_dict()
    encoder.load_state_dict(sd, strict=True)


def test_inference_layout_bf16_tcgen05_gemms_close_to_fp32_golden(encoder):
    """to_inference(): bf16 ViT trunks whose Linear layers run on the tcgen05 GEMM (bias/GELU/residual fused), fp32
    channels_last heads, replayed as a CUDA graph.  Tolerance vs the reference's fp32 golden: bf16 operands through
    ~60 layers -> mean |err| <= 3e-2 * scale (measured ~1e-2)."""
    import copy
    import torch
    from styl3r_b200.encoder import GraphedEncoder
    from tests.encoder_weights import make_inputs
    g = np.load(GOLD / "encoder_golden.npz")
    enc = copy.deepcopy(encoder).to_inference(torch.bfloat16)
    context, style = make_inputs(1, 2, 256, seed=1234, device="cuda")
    fast = GraphedEncoder(enc)
    out = fast(context, style)
    out2 = fast(context, style)  # replay
    torch.cuda.synchronize()
    assert torch.equal(out.means, out2.means)
    for name, t in [("means", out.means), ("harmonics", out.harmonics), ("opacities", out.opacities)]:
        ref, scale = g[f"b1v2_{name}"], float(g[f"b1v2_{name}_stats"][2])
        err = np.abs(sample(t) - ref)
        assert err.mean() <= 3e-2 * scale, f"{name}: mean err {err.mean():.3e} scale {scale:.3e}"


def test_stream_branches_do_not_change_results(encoder):
    """to_inference(branches=True) runs independent sub-graphs (content | style ViT, the two decoders, dec_blocks |
    dec_blocks2, K/V projections | self-attention, the DPT pyramids) on concurrent streams: same kernels, same
    operands -> bit-identical Gaussians, eagerly and as a replayed CUDA graph; repeated replays stay identical
    (no cross-stream memory reuse hazard)."""
    import copy
    import torch
    from styl3r_b200.encoder import GraphedEncoder
    from tests.encoder_weights import make_inputs
    context, style = make_inputs(1, 3, 256, seed=99, device="cuda")
    seq = copy.deepcopy(encoder).to_inference(torch.bfloat16, branches=False)
    with torch.no_grad():
        ref = seq(context, style)
    torch.cuda.synchronize()
    par = copy.deepcopy(encoder).to_inference(torch.bfloat16, branches=True)
    with torch.no_grad():
        eager = par(context, style)
    torch.cuda.synchronize()
    fast = GraphedEncoder(par)
    outs = [fast(context, style) for _ in range(3)]
    torch.cuda.synchronize()
    for name in ("means", "covariances", "harmonics", "opacities"):
        assert torch.equal(getattr(eager, name), getattr(ref, name)), f"eager branches changed {name}"
        assert torch.equal(getattr(outs[-1], name), getattr(ref, name)), f"graphed branches changed {name}"
