"""Differentiable float64 PyTorch mirror of the rasterizer — the *gradient* oracle.

TEST INFRASTRUCTURE ONLY (see raster_oracle.c).  Given the discrete structure produced by the
C oracle's forward pass (sorted per-tile lists), this re-evaluates projection, EWA covariance,
colour and the front-to-back blend as a smooth torch graph so that autograd yields the exact
derivative of the piecewise-smooth function the rasterizer computes, including the derivative
w.r.t. a left-multiplied camera perturbation  w2c' = SE3_exp(tau) @ w2c  at tau = 0
(the convention of src/misc/cam_utils.py:103-137 consumed via cuda_splatting.py:127-128).

Conventions kept from the upstream backward pass:
  * the min(0.99, .) clamp on alpha is treated as pass-through for gradients,
  * threshold decisions (power > 0, alpha < 1/255, T < 1e-4) are constants.
"""
from __future__ import annotations

import torch

SH_C0 = 0.28209479177387814
SH_C1 = 0.4886025119029199
SH_C2 = [1.0925484305920792, -1.0925484305920792, 0.31539156525252005, -1.0925484305920792, 0.5462742152960396]
SH_C3 = [-0.5900435899266435, 2.890611442640554, -0.4570457994644658, 0.3731763325901154,
         -0.4570457994644658, 1.445305721320277, -0.5900435899266435]


def hat(v):
    z = torch.zeros((), dtype=v.dtype)
    return torch.stack([torch.stack([z, -v[2], v[1]]), torch.stack([v[2], z, -v[0]]), torch.stack([-v[1], v[0], z])])


def perturb_w2c(w2c, tau):
    """First-order SE3_exp(tau) @ w2c — exact value and derivative at tau = 0."""
    T = torch.eye(4, dtype=w2c.dtype)
    T = T + torch.zeros_like(T)
    top = torch.cat([hat(tau[3:]), tau[:3, None]], dim=1)
    T = T + torch.cat([top, torch.zeros(1, 4, dtype=w2c.dtype)], dim=0)
    return T @ w2c


def sh_color(deg, sh, dirs):
    x, y, z = dirs[:, 0:1], dirs[:, 1:2], dirs[:, 2:3]
    r = SH_C0 * sh[:, 0]
    if deg > 0:
        r = r - SH_C1 * y * sh[:, 1] + SH_C1 * z * sh[:, 2] - SH_C1 * x * sh[:, 3]
    if deg > 1:
        xx, yy, zz, xy, yz, xz = x * x, y * y, z * z, x * y, y * z, x * z
        r = (r + SH_C2[0] * xy * sh[:, 4] + SH_C2[1] * yz * sh[:, 5] + SH_C2[2] * (2 * zz - xx - yy) * sh[:, 6]
             + SH_C2[3] * xz * sh[:, 7] + SH_C2[4] * (xx - yy) * sh[:, 8])
    if deg > 2:
        r = (r + SH_C3[0] * y * (3 * xx - yy) * sh[:, 9] + SH_C3[1] * xy * z * sh[:, 10]
             + SH_C3[2] * y * (4 * zz - xx - yy) * sh[:, 11] + SH_C3[3] * z * (2 * zz - 3 * xx - 3 * yy) * sh[:, 12]
             + SH_C3[4] * x * (4 * zz - xx - yy) * sh[:, 13] + SH_C3[5] * z * (xx - yy) * sh[:, 14]
             + SH_C3[6] * x * (xx - 3 * yy) * sh[:, 15])
    return torch.clamp_min(r + 0.5, 0.0)


def render(means, cov6, opac, w2c, proj_raw, tanx, tany, W, H, bg, point_list, ranges, shs=None, colors=None,
           deg=0, tau=None):
    """All tensor inputs float64. w2c, proj_raw: mathematical 4x4 (row-major). Returns (color[3,H,W], depth[H,W])."""
    dt = torch.float64
    if tau is not None:
        w2c = perturb_w2c(w2c, tau)
    R, tr = w2c[:3, :3], w2c[:3, 3]
    t = means @ R.T + tr
    fx, fy = W / (2.0 * tanx), H / (2.0 * tany)
    limx, limy = 1.3 * tanx, 1.3 * tany
    tz = t[:, 2]
    tcx = torch.clamp(t[:, 0] / tz, -limx, limx) * tz
    tcy = torch.clamp(t[:, 1] / tz, -limy, limy) * tz
    P = means.shape[0]
    J = torch.zeros(P, 2, 3, dtype=dt)
    J[:, 0, 0] = fx / tz
    J[:, 0, 2] = -fx * tcx / (tz * tz)
    J[:, 1, 1] = fy / tz
    J[:, 1, 2] = -fy * tcy / (tz * tz)
    S = torch.stack([cov6[:, 0], cov6[:, 1], cov6[:, 2], cov6[:, 1], cov6[:, 3], cov6[:, 4], cov6[:, 2], cov6[:, 4],
                     cov6[:, 5]], dim=1).reshape(P, 3, 3)
    M = J @ R
    cov2 = M @ S @ M.transpose(1, 2)
    a, b, c = cov2[:, 0, 0] + 0.3, cov2[:, 0, 1], cov2[:, 1, 1] + 0.3
    det = a * c - b * b
    cA, cB, cC = c / det, -b / det, a / det
    th = torch.cat([t, torch.ones(P, 1, dtype=dt)], dim=1) @ proj_raw.T
    w = 1.0 / (th[:, 3] + 1e-7)
    px = ((th[:, 0] * w + 1.0) * W - 1.0) * 0.5
    py = ((th[:, 1] * w + 1.0) * H - 1.0) * 0.5
    if colors is not None:
        rgb = colors
    else:
        campos = -(R.T @ tr)
        d = means - campos
        rgb = sh_color(deg, shs, d / d.norm(dim=1, keepdim=True))
    depth_g = tz

    gx, gy = (W + 15) // 16, (H + 15) // 16
    color = torch.zeros(3, H, W, dtype=dt)
    depth = torch.zeros(H, W, dtype=dt)
    bg = torch.as_tensor(bg, dtype=dt)
    for tile in range(gx * gy):
        r0, r1 = int(ranges[tile, 0]), int(ranges[tile, 1])
        x0, y0 = (tile % gx) * 16, (tile // gx) * 16
        x1, y1 = min(x0 + 16, W), min(y0 + 16, H)
        ys, xs = torch.meshgrid(torch.arange(y0, y1, dtype=dt), torch.arange(x0, x1, dtype=dt), indexing="ij")
        npx = ys.numel()
        if r1 > r0:
            ids = torch.as_tensor(point_list[r0:r1].astype("int64"))
            dx = px[ids][None, :] - xs.reshape(-1, 1)
            dy = py[ids][None, :] - ys.reshape(-1, 1)
            power = -0.5 * (cA[ids] * dx * dx + cC[ids] * dy * dy) - cB[ids] * dx * dy
            G = torch.exp(power)
            araw = opac[ids][None, :] * G
            alpha = araw + (torch.clamp(araw, max=0.99) - araw).detach()
            valid = ((power <= 0) & (alpha >= 1.0 / 255.0)).detach()
            aeff = torch.where(valid, alpha, torch.zeros_like(alpha))
            one_m = 1.0 - aeff
            Texcl = torch.cumprod(torch.cat([torch.ones(npx, 1, dtype=dt), one_m[:, :-1]], dim=1), dim=1)
            done = (valid & ((Texcl * one_m) < 1e-4)).detach()
            live = (torch.cumsum(done.to(torch.int64), dim=1) == 0)
            wgt = aeff * Texcl * live
            Tfin = torch.prod(torch.where(live, one_m, torch.ones_like(one_m)), dim=1)
            col = wgt @ rgb[ids]
            dep = wgt @ depth_g[ids]
        else:
            Tfin = torch.ones(npx, dtype=dt)
            col = torch.zeros(npx, 3, dtype=dt)
            dep = torch.zeros(npx, dtype=dt)
        col = col + Tfin[:, None] * bg[None, :]
        color[:, y0:y1, x0:x1] = col.T.reshape(3, y1 - y0, x1 - x0)
        depth[y0:y1, x0:x1] = dep.reshape(y1 - y0, x1 - x0)
    return color, depth
