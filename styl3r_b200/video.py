"""Video trajectory rendering - mirror of `render_video_generic` /
`render_video_interpolation` (infer_model_re10k.py:179-255): frame times with cosine ease, extrinsics / intrinsics
interpolated between the first and the last context camera, every frame rendered by the decoder, clip -> uint8, and
the optional ping-pong loop.

The reference renders the stylised *and* the plain Gaussians (and discards the latter), frame by frame inside
`render_cuda`'s per-view Python loop; here the `num_frames` views of a trajectory are ONE batched launch chain of the
rasterizer (decoder_splatting_cuda.DecoderSplattingCUDA.forward) and the trajectory itself is built on the device
(styl3r_b200.trajectory).  Returned: uint8 frames [T', 3, h, w] (what the reference hands to wandb / moviepy; MP4
encoding is outside the path)."""
from __future__ import annotations

from typing import Callable, Tuple

import torch
from torch import Tensor

from .trajectory import (generate_wobble, generate_wobble_transformation, interpolate_extrinsics, interpolate_intrinsics,
                         smooth_time)

TrajectoryFn = Callable[[Tensor], Tuple[Tensor, Tensor]]


def render_video_generic(gaussians, decoder, batch: dict, trajectory_fn: TrajectoryFn, num_frames: int = 60,
                         smooth: bool = True, loop_reverse: bool = True) -> Tensor:
    """infer_model_re10k.py:179-232 for the stylised Gaussians.  `batch["context"]` needs image [b,v,3,h,w], near/far
    [b,v].  Returns uint8 [T', 3, h, w] on the device (T' = 2*num_frames - 2 with loop_reverse)."""
    ctx = batch["context"]
    device = ctx["image"].device
    t = smooth_time(num_frames, device, smooth)
    extrinsics, intrinsics = trajectory_fn(t)
    h, w = ctx["image"].shape[-2:]
    near = ctx["near"][:, :1].expand(-1, num_frames)
    far = ctx["far"][:, :1].expand(-1, num_frames)
    out = decoder.forward(gaussians, extrinsics, intrinsics, near, far, (h, w), "depth")
    video = (out.color[0].clip(min=0, max=1) * 255).type(torch.uint8)
    if loop_reverse:  # pack([video, video[::-1][1:-1]])
        video = torch.cat((video, video.flip(0)[1:-1]), dim=0)
    return video


def render_video_interpolation(gaussians, decoder, batch: dict, num_frames: int = 60, smooth: bool = True,
                               loop_reverse: bool = True) -> Tensor:
    """infer_model_re10k.py:235-255: interpolate between the first and last context camera of scene 0."""
    ctx = batch["context"]

    def trajectory_fn(t):
        extrinsics = interpolate_extrinsics(ctx["extrinsics"][0, 0], ctx["extrinsics"][0, -1], t)
        intrinsics = interpolate_intrinsics(ctx["intrinsics"][0, 0], ctx["intrinsics"][0, -1], t)
        return extrinsics[None], intrinsics[None]

    return render_video_generic(gaussians, decoder, batch, trajectory_fn, num_frames, smooth, loop_reverse)


def render_video_wobble(gaussians, decoder, batch: dict, num_frames: int = 60, smooth: bool = True,
                        loop_reverse: bool = True):
    """model_wrapper_style.py:632-654: wobble around the first context camera with a quarter of the context baseline as
    radius (needs exactly two context views, else None like the reference)."""
    ctx = batch["context"]
    if ctx["extrinsics"].shape[1] != 2:
        return None

    def trajectory_fn(t):
        delta = (ctx["extrinsics"][:, 0, :3, 3] - ctx["extrinsics"][:, 1, :3, 3]).norm(dim=-1)
        extrinsics = generate_wobble(ctx["extrinsics"][:, 0], delta * 0.25, t)
        return extrinsics, ctx["intrinsics"][:, 0, None].expand(-1, t.shape[0], -1, -1)

    return render_video_generic(gaussians, decoder, batch, trajectory_fn, num_frames, smooth, loop_reverse)


def render_video_interpolation_exaggerated(gaussians, decoder, batch: dict, num_frames: int = 300, smooth: bool = False,
                                           loop_reverse: bool = False):
    """model_wrapper_style.py:684-728: interpolation extrapolated to t in [-2, 3] with a 5-turn wobble of half the
    baseline on top (two context views only)."""
    ctx = batch["context"]
    if ctx["extrinsics"].shape[1] != 2:
        return None

    def trajectory_fn(t):
        delta = (ctx["extrinsics"][:, 0, :3, 3] - ctx["extrinsics"][:, 1, :3, 3]).norm(dim=-1)
        tf = generate_wobble_transformation(delta * 0.5, t, 5, scale_radius_with_t=False)
        extrinsics = interpolate_extrinsics(ctx["extrinsics"][0, 0], ctx["extrinsics"][0, 1], t * 5 - 2)
        intrinsics = interpolate_intrinsics(ctx["intrinsics"][0, 0], ctx["intrinsics"][0, 1], t * 5 - 2)
        return extrinsics @ tf, intrinsics[None]

    return render_video_generic(gaussians, decoder, batch, trajectory_fn, num_frames, smooth, loop_reverse)
