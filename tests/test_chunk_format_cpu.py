"""RE10K chunk format (host side) against golden vectors produced by the REFERENCE's own
convert_poses / convert_images / camera_normalization (tests/golden/make_chunk_golden.py).  Bit-exact."""
from pathlib import Path

import numpy as np
import torch

GOLD = np.load(Path(__file__).parent / "golden" / "chunk_golden.npz")


def _jpegs():
    out, off = [], 0
    for n in GOLD["n_jpeg"]:
        out.append(torch.tensor(GOLD["jpeg"][off:off + int(n)]))
        off += int(n)
    return out


def test_convert_poses_images_and_normalisation_match_reference():
    from styl3r_b200.chunk_format import camera_normalization, convert_images, convert_poses
    extr, intr = convert_poses(torch.tensor(GOLD["cameras"]))
    assert np.array_equal(extr.numpy(), GOLD["extrinsics"]) and np.array_equal(intr.numpy(), GOLD["intrinsics"])
    imgs = convert_images(_jpegs())
    assert imgs.dtype == torch.float32 and np.array_equal(imgs.numpy(), GOLD["images"])
    norm = camera_normalization(extr[0:1], extr)
    assert np.array_equal(norm.numpy(), GOLD["normalized"])
    assert np.allclose(norm[0].numpy(), np.eye(4), atol=1e-6)


def test_assemble_example_unit_baseline_relative_pose_and_bounds():
    from styl3r_b200.chunk_format import assemble_example, convert_poses
    example = {"key": "scene0", "cameras": torch.tensor(GOLD["cameras"]), "images": _jpegs()}
    ctx, tgt = torch.tensor([0, 2]), torch.tensor([1])
    ex = assemble_example(example, ctx, tgt, torch.rand(3, 8, 8), "s.jpg")
    assert ex["scene"] == "scene0" and ex["style"]["image_name"] == "s.jpg"
    assert ex["context"]["image"].shape == (2, 3, 36, 64) and ex["target"]["image"].shape == (1, 3, 36, 64)
    E = ex["context"]["extrinsics"]
    assert np.allclose(E[0].numpy(), np.eye(4), atol=1e-6)                       # relative to the first context camera
    assert abs(float((E[0, :3, 3] - E[1, :3, 3]).norm()) - 1.0) <= 1e-5          # unit baseline
    extr, _ = convert_poses(example["cameras"])
    scale = float((extr[0, :3, 3] - extr[2, :3, 3]).norm())
    assert np.allclose(ex["context"]["near"].numpy(), 0.1 / scale) and np.allclose(ex["target"]["far"].numpy(), 100.0 / scale)
    assert assemble_example(example, ctx, tgt, torch.rand(3, 8, 8), baseline_min=1e9) is None   # skipped like the reference
