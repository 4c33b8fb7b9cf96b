"""Input staging in front of the encoder - host-side mirror, same names / argument meaning, of

    rescale, center_crop, rescale_and_crop, apply_crop_shim(_to_views)   src/dataset/shims/crop_shim.py:11-100
    normalize_image, apply_normalize_shim                                src/dataset/shims/normalize_shim.py:15-27
    apply_style_image_augmentation                                       src/dataset/shims/augmentation_shim.py:40-62

The reference resizes every image on the CPU through PIL (device -> host -> uint8 -> `Image.resize(LANCZOS)` -> host ->
device).  Here the whole batch is resized, cropped and (optionally) normalised by two CUDA kernels
(`s3r_rescale_crop`, csrc/resize.cu) that reproduce Pillow's 8-bit fixed-point LANCZOS bit for bit; only the tap
tables (a few KB per image size, cached) are built on the host.  CUDA only (no CPU fallback)."""
from __future__ import annotations

import ctypes as C
import math
from functools import lru_cache
from typing import Optional, Tuple

import numpy as np
import torch
from torch import Tensor

from . import _lib

PRECISION_BITS = 22  # Pillow Resample.c: 32 - 8 - 2


@lru_cache(maxsize=64)
def lanczos_taps(in_size: int, out_size: int) -> Tuple[np.ndarray, np.ndarray, int]:
    """Pillow's precompute_coeffs + normalize_coeffs_8bpc for the LANCZOS filter (support 3): per output pixel the
    first input pixel, the tap count and the int32 fixed-point taps.  float64 like the C code."""
    scale = in_size / out_size
    fscale = max(scale, 1.0)
    support = 3.0 * fscale
    ksize = int(math.ceil(support)) * 2 + 1
    center = (np.arange(out_size, dtype=np.float64) + 0.5) * scale
    xmin = np.maximum((center - support + 0.5).astype(np.int64), 0)          # C (int) cast of a positive double
    xmax = np.minimum((center + support + 0.5).astype(np.int64), in_size)
    count = xmax - xmin
    x = np.arange(ksize, dtype=np.float64)[None, :]
    arg = (x + xmin[:, None] - center[:, None] + 0.5) * (1.0 / fscale)
    def sinc(v):
        pv = v * math.pi
        with np.errstate(invalid="ignore", divide="ignore"):
            return np.where(v == 0.0, 1.0, np.sin(pv) / pv)
    w = np.where((arg >= -3.0) & (arg < 3.0), sinc(arg) * sinc(arg / 3.0), 0.0)
    w = np.where(x < count[:, None], w, 0.0)
    ww = np.zeros(out_size, np.float64)
    for j in range(ksize):  # sequential accumulation order of the C loop
        ww = ww + w[:, j]
    k = np.where(ww[:, None] != 0.0, w / np.where(ww == 0.0, 1.0, ww)[:, None], w)
    kk = np.trunc(k * (1 << PRECISION_BITS) + np.where(k < 0, -0.5, 0.5)).astype(np.int32)
    bounds = np.stack((xmin, count), axis=1).astype(np.int32)
    return bounds, kk, ksize


_dev_taps: dict = {}


def _taps_on(device, in_size: int, out_size: int):
    key = (str(device), in_size, out_size)
    if key not in _dev_taps:
        b, k, ks = lanczos_taps(in_size, out_size)
        _dev_taps[key] = (torch.from_numpy(b.copy()).to(device), torch.from_numpy(k.copy()).to(device), ks)
    return _dev_taps[key]


def _resize_window(images: Tensor, scaled: Tuple[int, int], window: Tuple[int, int, int, int],
                   mean: Optional[Tensor] = None, std: Optional[Tensor] = None) -> Tensor:
    """images [..., c, h, w] float in [0,1] -> Pillow-LANCZOS resize to `scaled`, window (row, col, h_out, w_out)."""
    if images.device.type != "cuda":
        raise _lib.S3RError("styl3r_b200.staging needs CUDA tensors (no CPU fallback)")
    *batch, c, h_in, w_in = images.shape
    hs, ws = scaled
    row, col, h_out, w_out = window
    x = images.detach().to(torch.float32).contiguous()
    planes = x.numel() // (h_in * w_in)
    hb, hk, hks = _taps_on(x.device, w_in, ws)
    vb, vk, vks = _taps_on(x.device, h_in, hs)
    scratch = torch.empty(max(planes * h_in * ws, 1), dtype=torch.uint8, device=x.device)
    out = torch.empty((*batch, c, h_out, w_out), dtype=torch.float32, device=x.device)
    p = lambda t: None if t is None else C.c_void_p(t.data_ptr())
    if mean is not None:
        mean = mean.to(x.device, torch.float32).contiguous()
        std = std.to(x.device, torch.float32).contiguous()
    _lib.check(_lib.lib().s3r_rescale_crop(p(x), planes, c, h_in, w_in, hs, ws, p(hb), p(hk), hks, p(vb), p(vk), vks, row,
                                           col, h_out, w_out, p(mean), p(std), p(scratch), p(out),
                                           C.c_void_p(torch.cuda.current_stream(x.device).cuda_stream)),
               "s3r_rescale_crop")
    return out.to(images.dtype)


def rescale(image: Tensor, shape: Tuple[int, int]) -> Tensor:
    """crop_shim.py:11-22: [3, h_in, w_in] -> [3, h, w]."""
    h, w = shape
    return _resize_window(image, (h, w), (0, 0, h, w))


def center_crop(images: Tensor, intrinsics: Tensor, shape: Tuple[int, int]):
    """crop_shim.py:25-51."""
    *_, h_in, w_in = images.shape
    h_out, w_out = shape
    row, col = (h_in - h_out) // 2, (w_in - w_out) // 2
    images = images[..., :, row:row + h_out, col:col + w_out]
    intrinsics = intrinsics.clone()
    intrinsics[..., 0, 0] *= w_in / w_out
    intrinsics[..., 1, 1] *= h_in / h_out
    return images, intrinsics


def _scaled_shape(h_in: int, w_in: int, shape):
    h_out, w_out = shape
    assert h_out <= h_in and w_out <= w_in
    scale_factor = max(h_out / h_in, w_out / w_in)
    h_scaled, w_scaled = round(h_in * scale_factor), round(w_in * scale_factor)
    assert h_scaled == h_out or w_scaled == w_out
    return h_scaled, w_scaled


def rescale_and_crop(images: Tensor, intrinsics: Tensor, shape: Tuple[int, int], normalize: bool = False,
                     mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)):
    """crop_shim.py:54-79: scale (LANCZOS) so that the image covers `shape`, centre-crop, adjust fx / fy.  The crop
    window is cut inside the vertical pass, so the scaled image is never materialised.  `normalize=True` additionally
    fuses `normalize_image` (normalize_shim.py:15-18) - what `apply_normalize_shim` does to the context images."""
    *_, h_in, w_in = images.shape
    h_out, w_out = shape
    hs, ws = _scaled_shape(h_in, w_in, shape)
    row, col = (hs - h_out) // 2, (ws - w_out) // 2
    m = s = None
    if normalize:
        m = torch.as_tensor(mean, dtype=torch.float32, device=images.device)
        s = torch.as_tensor(std, dtype=torch.float32, device=images.device)
    out = _resize_window(images, (hs, ws), (row, col, h_out, w_out), m, s)
    intrinsics = intrinsics.clone()
    intrinsics[..., 0, 0] *= ws / w_out
    intrinsics[..., 1, 1] *= hs / h_out
    return out, intrinsics


def apply_crop_shim_to_views(views: dict, shape: Tuple[int, int]) -> dict:
    images, intrinsics = rescale_and_crop(views["image"], views["intrinsics"], shape)
    return {**views, "image": images, "intrinsics": intrinsics}


def apply_crop_shim(example: dict, shape: Tuple[int, int]) -> dict:
    """crop_shim.py:92-100."""
    return {**example, "context": apply_crop_shim_to_views(example["context"], shape),
            "target": apply_crop_shim_to_views(example["target"], shape)}


def normalize_image(tensor: Tensor, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> Tensor:
    """normalize_shim.py:15-18."""
    mean = torch.as_tensor(mean, dtype=tensor.dtype, device=tensor.device).view(-1, 1, 1)
    std = torch.as_tensor(std, dtype=tensor.dtype, device=tensor.device).view(-1, 1, 1)
    return (tensor - mean) / std


def apply_normalize_shim(batch: dict, mean=(0.5, 0.5, 0.5), std=(0.5, 0.5, 0.5)) -> dict:
    """normalize_shim.py:21-27 (mutates batch["context"]["image"] like the reference)."""
    batch["context"]["image"] = normalize_image(batch["context"]["image"], mean, std)
    return batch


def apply_style_image_augmentation(style_image: Tensor, stage: str = "val", size: int = 256) -> Tensor:
    """augmentation_shim.py:40-62: short side -> 256 (long side int(ratio * 256)), then CenterCrop(256) for every
    stage.  torchvision's CenterCrop offsets are round((H - 256) / 2)."""
    _, H, W = style_image.shape
    if H < W:
        H, W = size, int(W / H * size)
    else:
        H, W = int(H / W * size), size
    top, left = int(round((H - size) / 2.0)), int(round((W - size) / 2.0))
    return _resize_window(style_image, (H, W), (top, left, size, size))
