"""Shared test plumbing: run a synthetic scene through the CPU oracle and through the CUDA C-ABI path with
bit-identical camera matrices."""
from __future__ import annotations

import numpy as np

from oracle import raster_oracle as ro


def oracle_scene(scene, scale_invariant=True, deg=0, bg=(0.0, 0.0, 0.0), use_sh=True, render=True):
    """Per-view oracle forward; returns (list of per-view dicts, list of camera dicts)."""
    H, W = scene["image_shape"]
    outs, cams = [], []
    bg = np.asarray(bg, np.float32)
    for v in range(scene["extrinsics"].shape[0]):
        cam = ro.camera_setup(scene["extrinsics"][v], scene["intrinsics"][v], scene["near"][v], scene["far"][v],
                              scale_invariant)
        m, c = ro.scale_gaussians(scene["means"], scene["covariances"], cam["scale"])
        shs = np.ascontiguousarray(scene["harmonics"].transpose(0, 2, 1))  # [P, M, 3]
        kw = dict(shs=shs, deg=deg) if use_sh else dict(colors=shs[:, 0, :])
        outs.append(ro.forward(m, ro.cov3x3_to_6(c), scene["opacities"], cam["view16"], cam["proj16"], cam["campos"],
                               W, H, cam["tanx"], cam["tany"], bg, render=render, **kw))
        cams.append(cam)
    return outs, cams


def gpu_scene(scene, cams, deg=0, bg=(0.0, 0.0, 0.0), use_sh=True, want_n_touched=True, cov_packed=False,
              capacity=None, check="sync", device="cuda"):
    """Same scene through styl3r_b200.rasterizer.forward_raw: one Gaussian set shared by all views."""
    import torch

    from styl3r_b200 import rasterizer as rz

    H, W = scene["image_shape"]
    V = len(cams)
    t = lambda a, dt=torch.float32: torch.as_tensor(np.ascontiguousarray(a), dtype=dt, device=device)
    means = t(scene["means"])[None]
    cov = t(ro.cov3x3_to_6(scene["covariances"]) if cov_packed else scene["covariances"])[None]
    opac = t(scene["opacities"])[None]
    shs = t(np.ascontiguousarray(scene["harmonics"].transpose(0, 2, 1)))[None]
    kw = dict(shs=shs, sh_degree=deg) if use_sh else dict(colors_precomp=shs[:, :, 0, :].contiguous())
    view = t(np.stack([c["view16"] for c in cams])).reshape(V, 4, 4)
    proj = t(np.stack([c["proj16"] for c in cams])).reshape(V, 4, 4)
    praw = t(np.stack([c["projraw16"] for c in cams])).reshape(V, 4, 4)
    campos = t(np.stack([c["campos"] for c in cams]))
    tanfov = t(np.stack([[c["tanx"], c["tany"]] for c in cams]))
    scales = t(np.array([c["scale"] for c in cams], np.float32))
    bgt = t(np.repeat(np.asarray(bg, np.float32)[None], V, 0))
    view_set = torch.zeros(V, dtype=torch.int32, device=device)
    return rz.forward_raw(means, cov, opac, view, proj, tanfov, bgt, W, H, campos=campos, projmatrix_raw=praw,
                          scales=scales, view_set=view_set, want_n_touched=want_n_touched, capacity=capacity,
                          check=check, **kw)
