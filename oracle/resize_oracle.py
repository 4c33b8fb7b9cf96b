"""TEST INFRASTRUCTURE ONLY (never imported by the product path).

CPU restatement (numpy) of the image staging step in front of the encoder:

    rescale          src/dataset/shims/crop_shim.py:11-22   float [3,h,w] in [0,1] -> uint8 -> PIL LANCZOS resize -> /255
    center_crop      src/dataset/shims/crop_shim.py:25-51   crop + fx *= w_in/w_out, fy *= h_in/h_out
    rescale_and_crop src/dataset/shims/crop_shim.py:54-79   scale so that the image covers the target, then centre crop
    normalize_image  src/dataset/shims/normalize_shim.py:15-18

The resize itself lives in a third-party dependency that is not vendored in /root/reference: **Pillow**
(`requirements.txt` does not pin it; the image used here has Pillow 12.2.0), `src/libImaging/Resample.c`:
`precompute_coeffs` (separable filter taps per output pixel, float64, normalised), `normalize_coeffs_8bpc`
(fixed point, PRECISION_BITS = 32 - 8 - 2 = 22), `ImagingResampleHorizontal_8bpc` then
`ImagingResampleVertical_8bpc` (int32 accumulation from 1 << 21, `>> 22`, clip to [0, 255]; the intermediate image is
uint8).  The published algorithm is restated below; it is PINNED against Pillow itself driven through the reference's
own `rescale` / `rescale_and_crop` (tests/golden/make_staging_golden.py -> tests/golden/staging_golden.npz, and live
in tests/test_staging_cpu.py when PIL is importable)."""
from __future__ import annotations

import math

import numpy as np

PRECISION_BITS = 32 - 8 - 2
LANCZOS_SUPPORT = 3.0


def _sinc(x: float) -> float:
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def lanczos_filter(x: float) -> float:
    """Resample.c: lanczos_filter - truncated sinc, a = 3."""
    if -3.0 <= x < 3.0:
        return _sinc(x) * _sinc(x / 3.0)
    return 0.0


def precompute_coeffs(in_size: int, out_size: int):
    """Resample.c: precompute_coeffs (box = the whole image) + normalize_coeffs_8bpc.
    Returns (bounds int32 [out,2] = (xmin, count), kk int32 [out, ksize], ksize)."""
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = LANCZOS_SUPPORT * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [lanczos_filter((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            k = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(k * (1 << PRECISION_BITS) + (-0.5 if k < 0 else 0.5))  # C cast: truncation toward zero
        bounds[xx] = (xmin, xmax)
    return bounds, kk, ksize


def _clip8(acc: np.ndarray) -> np.ndarray:
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def resize_lanczos_u8(img: np.ndarray, h_out: int, w_out: int) -> np.ndarray:
    """uint8 [h, w, c] -> uint8 [h_out, w_out, c]: horizontal pass, uint8 intermediate, vertical pass (Pillow skips a pass
    whose size does not change)."""
    h, w, _ = img.shape
    cur = img
    if w_out != w:
        bounds, kk, _ = precompute_coeffs(w, w_out)
        out = np.empty((h, w_out, cur.shape[2]), np.uint8)
        for xx in range(w_out):
            x0, n = bounds[xx]
            acc = (cur[:, x0:x0 + n, :].astype(np.int64) * kk[xx, :n].astype(np.int64)[None, :, None]).sum(1)
            out[:, xx, :] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        cur = out
    if h_out != h:
        bounds, kk, _ = precompute_coeffs(h, h_out)
        out = np.empty((h_out, cur.shape[1], cur.shape[2]), np.uint8)
        for yy in range(h_out):
            y0, n = bounds[yy]
            acc = (cur[y0:y0 + n].astype(np.int64) * kk[yy, :n].astype(np.int64)[:, None, None]).sum(0)
            out[yy] = _clip8(acc + (1 << (PRECISION_BITS - 1)))
        cur = out
    return cur


def rescale(image: np.ndarray, shape) -> np.ndarray:
    """crop_shim.py:11-22 on a float32 [3,h,w] array."""
    h, w = shape
    u8 = np.clip(image.astype(np.float32) * np.float32(255), 0, 255).astype(np.uint8)  # torch .type(uint8): truncation
    out = resize_lanczos_u8(np.transpose(u8, (1, 2, 0)), h, w)
    return np.transpose((out / 255).astype(np.float32), (2, 0, 1))


def scaled_shape(h_in: int, w_in: int, shape):
    h_out, w_out = shape
    assert h_out <= h_in and w_out <= w_in
    scale_factor = max(h_out / h_in, w_out / w_in)
    h_scaled, w_scaled = round(h_in * scale_factor), round(w_in * scale_factor)
    assert h_scaled == h_out or w_scaled == w_out
    return h_scaled, w_scaled


def rescale_and_crop(images: np.ndarray, intrinsics: np.ndarray, shape):
    """crop_shim.py:54-79 + 25-51 on float32 [n,3,h,w] / [n,3,3]."""
    n, _, h_in, w_in = images.shape
    h_out, w_out = shape
    hs, ws = scaled_shape(h_in, w_in, shape)
    scaled = np.stack([rescale(im, (hs, ws)) for im in images])
    row, col = (hs - h_out) // 2, (ws - w_out) // 2
    K = intrinsics.astype(np.float32).copy()
    K[:, 0, 0] *= np.float32(ws / w_out)
    K[:, 1, 1] *= np.float32(hs / h_out)
    return scaled[:, :, row:row + h_out, col:col + w_out], K


def style_shape(h: int, w: int, size: int = 256):
    """apply_style_image_augmentation (augmentation_shim.py:40-62): short side -> 256, long side int(ratio*256)."""
    if h < w:
        return size, int(w / h * size)
    return int(h / w * size), size
