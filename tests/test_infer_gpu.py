"""GPU integration: the composed inference path (styl3r_b200.infer.infer = infer_model_re10k.py:404-560 on the B200
pieces) from raw 360x640 frames to stylised renders, interpolation video and .ply files, in the bf16 inference layout
under CUDA-graph replay, and its agreement with the fp32 module path."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def test_infer_end_to_end(tmp_path):
    import torch
    from styl3r_b200 import synthetic as syn
    from styl3r_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, GraphedEncoder, get_encoder
    from styl3r_b200.infer import infer
    from styl3r_b200.ply_export import read_ply
    from tests.encoder_weights import fill_named_weights
    torch.manual_seed(0)
    enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
    fill_named_weights(enc)
    enc = enc.cuda().eval()
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True)).cuda()
    g = torch.Generator(device="cuda").manual_seed(3)
    yy, xx = torch.meshgrid(torch.linspace(0, 3, 360, device="cuda"), torch.linspace(0, 5, 640, device="cuda"), indexing="ij")
    frame = lambda k: torch.stack([0.5 + 0.4 * torch.sin(xx * (c + 1) + k) * torch.cos(yy + c) for c in range(3)])
    ctx = torch.stack([frame(0.0), frame(0.7)])
    tgt = torch.stack([frame(0.3)])
    style = torch.rand(3, 300, 400, device="cuda", generator=g)
    sc = syn.make_scene(seed=4, v=2, V=1, hw=256)
    t = lambda a: torch.as_tensor(a).cuda()
    K = torch.tensor([[0.5, 0, 0.5], [0, 0.9, 0.5], [0, 0, 1.0]], device="cuda")
    args = dict(context_images=ctx, context_intrinsics=K.expand(2, 3, 3), context_extrinsics=t(sc["context_extrinsics"]),
                target_images=tgt, target_intrinsics=K.expand(1, 3, 3), target_extrinsics=t(sc["extrinsics"]), style_image=style)
    # fp32 module path
    ref = infer(enc, dec, **args, num_video_frames=0)
    # bf16 inference layout, CUDA-graph replay, with pose alignment, video and .ply export
    fast = GraphedEncoder(enc.to_inference(torch.bfloat16))
    out = infer(fast, dec, **args, pose_align_steps=3, num_video_frames=8, output_dir=tmp_path)
    torch.cuda.synchronize()
    assert out.stylized_color.shape == (1, 1, 3, 256, 256) and torch.isfinite(out.stylized_color).all()
    assert out.video.dtype == torch.uint8 and out.video.shape == (14, 3, 256, 256)
    assert out.extrinsics.shape == (1, 1, 4, 4)
    # intrinsics were adjusted by the crop: fx *= 455/256 (640x360 -> 455x256 -> centre crop 256)
    for name in ("gaussians.ply", "stylized_gaussians.ply"):
        names, rows = read_ply(tmp_path / name)
        assert rows.shape == (2 * 65536, 17) and np.isfinite(rows).all() and names[0] == "x" and names[-1] == "rot_3"
    # bf16 trunks vs fp32: same Gaussians up to mixed-precision noise (geometry is shared by both passes)
    a, b = out.gaussians.means, ref.gaussians.means
    assert (a - b).abs().mean().item() <= 3e-2 * b.std().item()
    assert not torch.equal(out.stylized_gaussians.harmonics, out.gaussians.harmonics)   # the style image changes appearance
    assert torch.equal(out.stylized_gaussians.means, out.gaussians.means) or \
        (out.stylized_gaussians.means - out.gaussians.means).abs().max().item() == 0.0   # ... and nothing else
