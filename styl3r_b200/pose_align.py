"""Test-time pose alignment on device — SURVEY §8 row f1.

Mirrors `test_step_align` (infer_model_re10k.py:79-161; src/model/model_wrapper_style.py:391-461): for
`pose_align_steps` iterations render the target views with zero camera deltas, back-propagate an image loss to the
deltas (rho = cam_trans_delta, theta = cam_rot_delta), take an Adam step, fold the step into the pose with
`w2c <- SE3_exp([rho, theta]) @ w2c` (src/misc/cam_utils.py:103-137) and reset the deltas.

B200 version: the world->camera matrices stay on the device for the whole loop (no inverse / per-view Python loop /
`.item()` per step), only dL/dtau is requested from the rasterizer backward (no Gaussian gradients are written), and
one iteration = camera kernel + raster forward + loss gradient + raster backward + Adam + SE3 update is captured in
ONE CUDA graph that is replayed `steps` times.

Loss: the reference sums EVERY configured loss (`for loss_fn in losses: total_loss += loss_fn.forward(...)`,
infer_model_re10k.py:121-124; the shipped config is `[mse, lpips]`).  Pass the same list as `losses` - callables
`loss(color [B,3,h,w], target [B,3,h,w]) -> scalar` whose sum is differentiated with torch autograd inside the captured
iteration - or an analytic `loss_grad(color, target) -> dL/dcolor`.  With neither, the MSE term alone is optimised
(`loss_mse` with weight 1): that equals the reference only for an MSE-only loss list, and refined poses differ from the
reference's whenever other terms (LPIPS needs the `lpips` package's VGG weights, which this repository does not ship)
are configured.
"""
from __future__ import annotations

from math import isqrt
from typing import Callable, Optional, Sequence

import torch
from torch import Tensor

from . import rasterizer as _rz
from .decoder.cuda_splatting import _sh_layout, camera_setup
from .pose import se3_update_w2c


def _mse_grad(color: Tensor, target: Tensor) -> Tensor:
    return (2.0 / color.numel()) * (color - target)


@torch.no_grad()
def pose_align(gaussians, extrinsics: Tensor, intrinsics: Tensor, near: Tensor, far: Tensor, image_shape,
               target_image: Tensor, steps: int = 50, rot_lr: float = 0.005, trans_lr: float = 0.005,
               background: Optional[Tensor] = None, scale_invariant: bool = True,
               loss_grad: Optional[Callable[[Tensor, Tensor], Tensor]] = None, use_graph: bool = True,
               betas=(0.9, 0.999), eps: float = 1e-8, losses: Optional[Sequence[Callable[[Tensor, Tensor], Tensor]]] = None):
    """gaussians: object with means[b,G,3], covariances[b,G,3,3], harmonics[b,G,3,d_sh], opacities[b,G];
    extrinsics [b,v,4,4] camera-to-world; intrinsics [b,v,3,3]; near/far [b,v]; target_image [b,v,3,h,w].
    Returns (refined extrinsics [b,v,4,4] camera-to-world, loss history [steps] on device)."""
    b, v = extrinsics.shape[:2]
    B = b * v
    h, w = image_shape
    dev = extrinsics.device
    if loss_grad is not None and losses:
        raise ValueError("pass either `losses` or `loss_grad`")
    if losses:
        loss_fns = list(losses)

        def loss_grad(color, tgt):  # autograd over the caller's loss list (the sum the reference back-propagates)
            with torch.enable_grad():
                c = color.detach().requires_grad_()
                total = sum(fn(c, tgt) for fn in loss_fns)
                return torch.autograd.grad(total, c)[0]
    loss_grad = loss_grad or _mse_grad
    K = intrinsics.reshape(B, 3, 3).float().contiguous()
    nr, fr = near.reshape(B).float().contiguous(), far.reshape(B).float().contiguous()
    target = target_image.reshape(B, 3, h, w).float().contiguous()
    bg = (background if background is not None else torch.zeros(3, device=dev)).float().expand(B, 3).contiguous()
    view_set = torch.arange(b, device=dev, dtype=torch.int32).repeat_interleave(v)
    means, cov, opac = gaussians.means.float().contiguous(), gaussians.covariances.float().contiguous(), \
        gaussians.opacities.float().contiguous()
    shs = _sh_layout(gaussians.harmonics.float())
    degree = min(isqrt(gaussians.harmonics.shape[-1]) - 1, 3)
    S, P, M = means.shape[0], means.shape[1], shs.shape[2]

    w2c = extrinsics.reshape(B, 4, 4).float().inverse().contiguous()
    m = torch.zeros(B, 6, device=dev)
    vv = torch.zeros(B, 6, device=dev)
    t = torch.zeros((), device=dev)
    lr = torch.tensor([trans_lr] * 3 + [rot_lr] * 3, device=dev)  # tau = (rho, theta)
    loss_hist = torch.zeros(steps, device=dev)
    it = torch.zeros((), dtype=torch.long, device=dev)
    overflowed = torch.zeros(1, dtype=torch.int64, device=dev)  # sticky across graph replays (any iteration, not the last)

    # capacity from one synchronous probe
    cam = camera_setup(w2c, K, nr, fr, scale_invariant, input_is_w2c=True)
    tensors = lambda c: (means, cov, opac, shs, None, c[0], c[1], c[2], c[3], c[4], c[5] if scale_invariant else None, bg,
                         view_set)
    probe = _rz.RasterPlan(tensors(cam), S, P, B, w, h, M, degree, 9, max(4 * P * B, 1 << 16))
    probe.launch()
    st = probe.ctx.status()
    cap = max(int(st["num_instances"] * 1.5) + 4096, 1 << 16)

    def iteration():
        c = camera_setup(w2c, K, nr, fr, scale_invariant, input_is_w2c=True)
        plan = _rz.RasterPlan(tensors(c), S, P, B, w, h, M, degree, 9, cap)
        plan.launch()
        overflowed.bitwise_or_(plan.ctx.view("status")[1:2])
        g_color = loss_grad(plan.color, target)
        loss_hist.index_put_((it,), ((plan.color - target) ** 2).mean())
        g = _rz.backward_raw(plan.ctx, g_color, None, need_pose=True, only_pose=True)["tau"]
        # Adam on parameters that are reset to zero every step (torch.optim.Adam semantics)
        t.add_(1.0)
        m.mul_(betas[0]).add_(g, alpha=1 - betas[0])
        vv.mul_(betas[1]).addcmul_(g, g, value=1 - betas[1])
        bc1 = 1 - betas[0] ** t
        bc2 = 1 - betas[1] ** t
        delta = -(lr / bc1) * m / ((vv.sqrt() / bc2.sqrt()) + eps)
        w2c.copy_(se3_update_w2c(w2c, delta[:, :3].contiguous(), delta[:, 3:].contiguous()))
        it.add_(1)
        return plan

    if use_graph and steps > 1:
        side = torch.cuda.Stream(device=dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):
            iteration()  # warm-up (also step 1 of the optimisation)
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            plan = iteration()
        for _ in range(steps - 1):  # capture records but does not execute
            graph.replay()
    else:
        for _ in range(steps):
            iteration()
    if int(overflowed.item()):  # an iteration that dropped instances corrupts every later Adam step
        raise _rz._lib.S3RError("pose_align: the instance capacity overflowed during an iteration (poses moved the "
                                "scene into more tiles than the probe saw); re-run - the capacity hint has grown")
    return w2c.inverse().reshape(b, v, 4, 4), loss_hist
