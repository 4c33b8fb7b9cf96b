"""Synthetic "trained-like" pixel-aligned Gaussian scenes (SURVEY.md §8d recipe).

Used by tests/ and bench.py.  Pure numpy; deterministic in (seed, v, V, hw).  Shapes follow the
reference's `Gaussians` dataclass (src/model/types.py:7-12): means[G,3], covariances[G,3,3],
harmonics[G,3,d_sh], opacities[G] with G = v*h*w, plus OpenCV camera-to-world extrinsics and
normalised intrinsics as `DecoderSplattingCUDA.forward` receives them
(src/model/decoder/decoder_splatting_cuda.py:37-49).
"""
from __future__ import annotations

import numpy as np

SH_C0 = 0.28209479177387814


def look_at_c2w(pos, pivot):
    """OpenCV convention (+x right, +y down, +z forward) camera-to-world looking at `pivot`."""
    f = pivot - pos
    f = f / np.linalg.norm(f)
    down = np.array([0.0, 1.0, 0.0])
    r = np.cross(down, f)
    r = r / np.linalg.norm(r)
    d = np.cross(f, r)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = r, d, f, pos
    return m


def quat_xyzw_to_matrix(q):
    """Same convention as the reference adapter (src/model/encoder/common/gaussians.py:8-30)."""
    i, j, k, r = q[..., 0], q[..., 1], q[..., 2], q[..., 3]
    two_s = 2.0 / ((q * q).sum(-1) + 1e-8)
    o = np.stack([
        1 - two_s * (j * j + k * k), two_s * (i * j - k * r), two_s * (i * k + j * r),
        two_s * (i * j + k * r), 1 - two_s * (i * i + k * k), two_s * (j * k - i * r),
        two_s * (i * k - j * r), two_s * (j * k + i * r), 1 - two_s * (i * i + j * j)], -1)
    return o.reshape(q.shape[:-1] + (3, 3))


def make_scene(seed: int = 1234, v: int = 2, V: int = 1, hw: int = 256, d_sh: int = 1):
    """Returns dict of float32 arrays for one scene: G = v*hw*hw Gaussians, V target cameras."""
    rng = np.random.default_rng(seed)
    a, b = rng.integers(1, 4), rng.integers(1, 4)
    K = np.array([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]])
    pivot = np.array([0.0, 0.0, 4.0])
    ctx = [look_at_c2w(np.array([i / max(v - 1, 1), 0.0, 0.0]), pivot) for i in range(v)]
    u = (np.arange(hw) + 0.5) / hw
    xs, ys = np.meshgrid(u, u, indexing="xy")  # row-major pixels: y outer, x inner
    rays = np.stack([(xs - 0.5) / 0.8, (ys - 0.5) / 0.8, np.ones_like(xs)], -1).reshape(-1, 3)
    means, covs, opac, sh = [], [], [], []
    for c2w in ctx:
        n = hw * hw
        d = (4 + 3 * np.sin(2 * np.pi * xs * a) * np.cos(2 * np.pi * ys * b)).reshape(-1) + rng.uniform(0, 0.5, n)
        pc = rays * d[:, None]
        means.append(pc @ c2w[:3, :3].T + c2w[:3, 3])
        s = np.minimum((d / (0.8 * hw))[:, None] * np.exp(rng.normal(0, 0.5, (n, 3))), 0.3)
        q = rng.normal(0, 1, (n, 4))
        q = q / np.linalg.norm(q, axis=-1, keepdims=True)
        Rm = quat_xyzw_to_matrix(q)
        RS = Rm * s[:, None, :]
        covs.append(RS @ RS.transpose(0, 2, 1))
        opac.append(1 / (1 + np.exp(-rng.normal(1, 1.5, n))))
        rgb = rng.uniform(0, 1, (n, 3))
        h = np.zeros((n, 3, d_sh))
        h[:, :, 0] = (rgb - 0.5) / SH_C0
        if d_sh > 1:
            h[:, :, 1:] = rng.normal(0, 0.1, (n, 3, d_sh - 1))
        sh.append(h)
    tgt = []
    for k in range(V):
        t = (k + 1) / (V + 1)
        tgt.append(look_at_c2w(np.array([t * (1.0 if v > 1 else 0.3), 0.0, 0.0]), pivot))
    f32 = np.float32
    return dict(
        means=np.concatenate(means).astype(f32), covariances=np.concatenate(covs).astype(f32),
        harmonics=np.concatenate(sh).astype(f32), opacities=np.concatenate(opac).astype(f32),
        extrinsics=np.stack(tgt).astype(f32), intrinsics=np.repeat(K[None], V, 0).astype(f32),
        near=np.full(V, 0.1, f32), far=np.full(V, 100.0, f32), context_extrinsics=np.stack(ctx).astype(f32),
        image_shape=(hw, hw),
    )


def make_small_scene(seed: int = 0, P: int = 600, W: int = 64, H: int = 48, V: int = 2, d_sh: int = 1,
                     big_frac: float = 0.05, behind_frac: float = 0.05):
    """Unstructured random Gaussians for parity tests: includes large splats, off-screen and behind-camera
    points, ragged image sizes (W, H need not be multiples of 16)."""
    rng = np.random.default_rng(seed)
    pts = np.stack([rng.uniform(-2.5, 2.5, P), rng.uniform(-2.0, 2.0, P), rng.uniform(1.5, 8.0, P)], -1)
    nb = int(P * behind_frac)
    if nb:
        pts[:nb, 2] = rng.uniform(-3.0, 0.3, nb)
    s = np.exp(rng.normal(np.log(0.04), 0.6, (P, 3)))
    nbig = int(P * big_frac)
    if nbig:
        s[nb:nb + nbig] *= rng.uniform(5, 30, (nbig, 1))
    q = rng.normal(0, 1, (P, 4))
    q /= np.linalg.norm(q, axis=-1, keepdims=True)
    Rm = quat_xyzw_to_matrix(q)
    RS = Rm * s[:, None, :]
    cov = RS @ RS.transpose(0, 2, 1)
    opac = 1 / (1 + np.exp(-rng.normal(0.5, 2.0, P)))
    h = np.zeros((P, 3, d_sh))
    h[:, :, 0] = (rng.uniform(-0.2, 1.2, (P, 3)) - 0.5) / SH_C0
    if d_sh > 1:
        h[:, :, 1:] = rng.normal(0, 0.3, (P, 3, d_sh - 1))
    K = np.array([[0.9, 0, 0.5], [0, 0.9 * W / H, 0.5], [0, 0, 1.0]])
    pivot = np.array([0.0, 0.0, 4.0])
    cams = [look_at_c2w(np.array([rng.uniform(-0.5, 0.5), rng.uniform(-0.3, 0.3), rng.uniform(-0.5, 0.2)]), pivot)
            for _ in range(V)]
    f32 = np.float32
    return dict(
        means=pts.astype(f32), covariances=cov.astype(f32), harmonics=h.astype(f32), opacities=opac.astype(f32),
        extrinsics=np.stack(cams).astype(f32), intrinsics=np.repeat(K[None], V, 0).astype(f32),
        near=np.full(V, 0.5, f32), far=np.full(V, 50.0, f32), image_shape=(H, W),
    )
