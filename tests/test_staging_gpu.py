"""GPU: s3r_rescale_crop (csrc/resize.cu) through styl3r_b200.staging vs the reference-generated golden vectors
(bit-exact: integer arithmetic) and vs the numpy oracle at the RE10K size (640x360 -> 256x256)."""
from pathlib import Path

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
GOLD = np.load(Path(__file__).parent / "golden" / "staging_golden.npz")
dec = lambda codes: (codes / 255).astype(np.float32)


@pytest.mark.parametrize("tag", ["re10k", "portrait", "same_w", "odd"])
def test_rescale_and_crop_bit_exact_vs_reference_golden(tag):
    import torch
    from styl3r_b200.staging import rescale_and_crop
    img = torch.tensor(GOLD[f"{tag}_in"].astype(np.float32)).cuda()
    K = torch.tensor(GOLD[f"{tag}_K"]).cuda()
    out, Kout = rescale_and_crop(img, K, tuple(int(v) for v in GOLD[f"{tag}_shape"]))
    assert np.array_equal(out.cpu().numpy(), dec(GOLD[f"{tag}_out"]))
    assert np.array_equal(Kout.cpu().numpy(), GOLD[f"{tag}_Kout"])
    # leading batch dims like the reference's [b, v, c, h, w] views
    out2, _ = rescale_and_crop(img[None], K[None], tuple(int(v) for v in GOLD[f"{tag}_shape"]))
    assert torch.equal(out2[0], out)


def test_fused_normalise_rescale_style_and_upscale():
    import torch
    from styl3r_b200.staging import apply_style_image_augmentation, rescale, rescale_and_crop
    img = torch.tensor(GOLD["re10k_in"].astype(np.float32)).cuda()
    K = torch.tensor(GOLD["re10k_K"]).cuda()
    out, _ = rescale_and_crop(img, K, (64, 64), normalize=True)
    assert np.array_equal(out.cpu().numpy(), GOLD["re10k_norm"])
    up = rescale(torch.tensor(GOLD["up_in"].astype(np.float32)).cuda(), (50, 96))
    assert np.array_equal(up.cpu().numpy(), dec(GOLD["up_out"]))
    sty = apply_style_image_augmentation(torch.tensor(GOLD["style_in"].astype(np.float32)).cuda(), "val")
    assert np.array_equal(sty.cpu().numpy(), dec(GOLD["style_out"]))


def test_re10k_full_size_equals_oracle_and_shim_surface():
    import torch
    from oracle import resize_oracle as ro
    from styl3r_b200.staging import apply_crop_shim
    rng = np.random.default_rng(1)
    img = rng.random((2, 3, 360, 640), dtype=np.float32)
    K = np.tile(np.array([[0.5, 0, 0.5], [0, 0.9, 0.5], [0, 0, 1]], np.float32), (2, 1, 1))
    ref, Kref = ro.rescale_and_crop(img, K, (256, 256))
    ex = {"context": {"image": torch.tensor(img).cuda(), "intrinsics": torch.tensor(K).cuda(), "near": 1},
          "target": {"image": torch.tensor(img[:1]).cuda(), "intrinsics": torch.tensor(K[:1]).cuda()}, "scene": "s"}
    out = apply_crop_shim(ex, (256, 256))
    assert out["scene"] == "s" and out["context"]["near"] == 1
    assert np.array_equal(out["context"]["image"].cpu().numpy(), ref)
    assert np.array_equal(out["context"]["intrinsics"].cpu().numpy(), Kref)
    assert np.array_equal(out["target"]["image"].cpu().numpy(), ref[:1])


def test_staging_has_no_cpu_fallback():
    import torch
    from styl3r_b200 import _lib
    from styl3r_b200.staging import rescale
    with pytest.raises(_lib.S3RError):
        rescale(torch.zeros(3, 8, 8), (4, 4))
