"""RE10K / DL3DV `.torch` chunk format in front of the staging step - host-side mirror of
`DatasetRE10kStyle.convert_poses / convert_images` and the per-example assembly of `__iter__`
(src/dataset/dataset_re10k_style.py:108-246) and of `camera_normalization` (src/misc/cam_utils.py:27-42):

    chunk = torch.load(path)                     # list of {"key", "cameras" [n,18] float32, "images" [n] uint8 JPEG bytes}
    extrinsics, intrinsics = convert_poses(example["cameras"])       # c2w [n,4,4], normalised K [n,3,3]
    images = convert_images([example["images"][i] for i in idx])     # [k,3,H,W] float32 in [0,1]
    ex = assemble_example(example, context_idx, target_idx, style_image)

A camera row is (fx, fy, cx, cy, 0, 0, w2c[3x4] row-major), intrinsics already normalised by the image size.
JPEG decoding stays on the host with PIL exactly like the reference (bit-identical pixels, `ToTensor` scaling); the
decoded frames are uploaded once and everything after them (`staging.rescale_and_crop` ...) runs on the device.  A GPU
JPEG decoder would not be bit-identical to libjpeg and is a separate component."""
from __future__ import annotations

from io import BytesIO
from typing import List, Optional, Sequence, Tuple

import numpy as np
import torch
from torch import Tensor


def convert_poses(poses: Tensor) -> Tuple[Tensor, Tensor]:
    """dataset_re10k_style.py:215-236: [b,18] -> (extrinsics c2w [b,4,4], intrinsics [b,3,3] normalised)."""
    b = poses.shape[0]
    intrinsics = torch.eye(3, dtype=torch.float32).repeat(b, 1, 1)
    fx, fy, cx, cy = poses[:, :4].T
    intrinsics[:, 0, 0] = fx
    intrinsics[:, 1, 1] = fy
    intrinsics[:, 0, 2] = cx
    intrinsics[:, 1, 2] = cy
    w2c = torch.eye(4, dtype=torch.float32).repeat(b, 1, 1)
    w2c[:, :3] = poses[:, 6:].reshape(b, 3, 4)
    return w2c.inverse(), intrinsics


def convert_images(images: Sequence[Tensor]) -> Tensor:
    """dataset_re10k_style.py:238-246: JPEG byte tensors -> [n,3,H,W] float32 in [0,1] (PIL decode + ToTensor)."""
    from PIL import Image
    out = []
    for image in images:
        im = Image.open(BytesIO(image.numpy().tobytes()))
        arr = np.asarray(im.convert("RGB") if im.mode != "RGB" else im, dtype=np.uint8)
        out.append(torch.from_numpy(arr.copy()).permute(2, 0, 1).to(torch.float32).div(255))
    return torch.stack(out)


def camera_normalization(pivotal_pose: Tensor, poses: Tensor) -> Tensor:
    """cam_utils.py:27-42: express all poses [N,4,4] in the frame of `pivotal_pose` [1,4,4]."""
    norm = torch.eye(4, dtype=torch.float32, device=pivotal_pose.device)[None] @ torch.inverse(pivotal_pose)
    return torch.bmm(norm.repeat(poses.shape[0], 1, 1), poses)


def get_bound(bound: str, num_views: int, near: float = 0.1, far: float = 100.0) -> Tensor:
    """infer_model_re10k.py:164-176 / dataset get_bound: constant near / far planes per view."""
    if bound not in ("near", "far"):
        raise ValueError("bound not found!")
    return torch.full((num_views,), near if bound == "near" else far, dtype=torch.float32)


def assemble_example(example: dict, context_indices: Tensor, target_indices: Tensor, style_image: Tensor,
                     style_image_name: str = "", make_baseline_1: bool = True, relative_pose: bool = True,
                     baseline_min: float = 1e-3, baseline_max: float = 1e10, device: Optional[torch.device] = None) -> Optional[dict]:
    """One un-batched example like `__iter__` yields before the crop shim (dataset_re10k_style.py:119-207): poses
    converted, images decoded, world rescaled to a unit context baseline, poses made relative to the first context
    camera, near / far divided by the scale.  Returns None where the reference skips the example (baseline out of
    range).  With `device` the tensors are uploaded (one copy each)."""
    extrinsics, intrinsics = convert_poses(example["cameras"])
    context_images = convert_images([example["images"][int(i)] for i in context_indices])
    target_images = convert_images([example["images"][int(i)] for i in target_indices])
    scale = 1.0
    if make_baseline_1:
        ctx = extrinsics[context_indices]
        a, b = ctx[0, :3, 3], ctx[-1, :3, 3]
        scale = (a - b).norm()
        if scale < baseline_min or scale > baseline_max:
            return None
        extrinsics[:, :3, 3] /= scale
    if relative_pose:
        extrinsics = camera_normalization(extrinsics[context_indices][0:1], extrinsics)
    mv = (lambda t: t.to(device)) if device is not None else (lambda t: t)

    def views(idx, imgs):
        return {"extrinsics": mv(extrinsics[idx]), "intrinsics": mv(intrinsics[idx]), "image": mv(imgs),
                "near": mv(get_bound("near", len(idx)) / scale), "far": mv(get_bound("far", len(idx)) / scale), "index": idx}

    return {"context": views(context_indices, context_images), "target": views(target_indices, target_images),
            "scene": example["key"], "style": {"image": mv(style_image), "image_name": style_image_name}}
