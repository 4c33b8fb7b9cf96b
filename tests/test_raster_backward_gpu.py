"""GPU parity: CUDA rasterizer backward (C-ABI) vs the CPU oracle's analytic backward, which is itself checked
against torch autograd in tests/test_oracle_cpu.py.  Tolerance: 2e-3 of the gradient's max magnitude (fp32 atomics
in a different summation order; exp via MUFU)."""
import numpy as np
import pytest

from oracle import raster_oracle as ro
from styl3r_b200 import synthetic as syn
from tests.helpers import gpu_scene, oracle_scene

pytestmark = pytest.mark.gpu
REL = 2e-3


def close(name, mine, ref, rel=REL):
    mine, ref = np.asarray(mine, np.float64), np.asarray(ref, np.float64)
    scale = np.abs(ref).max() + 1e-12
    err = np.abs(mine - ref).max()
    assert err <= rel * scale + 1e-7, f"{name}: max err {err:.3e} vs scale {scale:.3e}"


def run(scene, deg=0, bg=(0.1, 0.2, 0.3), use_sh=True, cov_packed=False, depth_grad=True, seed=0):
    import torch

    from styl3r_b200 import rasterizer as rz

    outs, cams = oracle_scene(scene, deg=deg, bg=bg, use_sh=use_sh)
    color, depth, opacity, radii, n_touched, ctx = gpu_scene(scene, cams, deg=deg, bg=bg, use_sh=use_sh,
                                                             cov_packed=cov_packed, want_n_touched=False)
    H, W = scene["image_shape"]
    V, P = len(cams), scene["means"].shape[0]
    rng = np.random.default_rng(seed)
    gc = rng.normal(size=(V, 3, H, W)).astype(np.float32)
    gd = (0.1 * rng.normal(size=(V, H, W))).astype(np.float32) if depth_grad else None
    g = rz.backward_raw(ctx, torch.as_tensor(gc).cuda(), None if gd is None else torch.as_tensor(gd).cuda())
    torch.cuda.synchronize()
    M = scene["harmonics"].shape[2]
    ref_means = np.zeros((P, 3)); ref_cov6 = np.zeros((P, 6)); ref_op = np.zeros(P)
    ref_sh = np.zeros((P, M, 3)); ref_col = np.zeros((P, 3)); ref_tau = np.zeros((V, 6)); ref_m2d = np.zeros((V, P, 2))
    for v, (o, cam) in enumerate(zip(outs, cams)):
        gb = ro.backward(o, gc[v], None if gd is None else gd[v], cam["projraw16"])
        s = float(cam["scale"])
        ref_means += s * gb["dL_dmeans"]
        ref_cov6 += s * s * gb["dL_dcov6"]
        ref_op += gb["dL_dopacity"]
        if use_sh:
            ref_sh += gb["dL_dsh"]
        else:
            ref_col += gb["dL_dcolor"]
        ref_tau[v] = gb["dL_dtau"]
        ref_m2d[v] = gb["dL_dmean2D"]
    close("means", g["means"][0].cpu().numpy(), ref_means)
    cov = g["cov"][0].cpu().numpy()
    if not cov_packed:
        c9 = cov.reshape(P, 3, 3)
        assert np.all(c9[:, [1, 2, 2], [0, 0, 1]] == 0), "lower triangle must stay zero (reference gathers the upper)"
        cov = c9[:, [0, 0, 0, 1, 1, 2], [0, 1, 2, 1, 2, 2]]
    close("cov", cov, ref_cov6)
    close("opacities", g["opacities"][0].cpu().numpy(), ref_op)
    if use_sh:
        close("shs", g["shs"][0].cpu().numpy(), ref_sh)
    else:
        close("colors", g["colors"][0].cpu().numpy(), ref_col)
    close("tau", g["tau"].cpu().numpy(), ref_tau)
    close("means2D", g["means2D"].cpu().numpy()[..., :2], ref_m2d)


@pytest.mark.parametrize("seed,P,W,H,V", [(0, 600, 64, 48, 2), (1, 3000, 96, 80, 3), (3, 1, 40, 24, 1)])
def test_backward_degree0(seed, P, W, H, V):
    run(syn.make_small_scene(seed=seed, P=P, W=W, H=H, V=V), seed=seed)


@pytest.mark.parametrize("deg,d_sh", [(1, 4), (3, 16)])
def test_backward_sh(deg, d_sh):
    run(syn.make_small_scene(seed=20 + deg, P=700, W=64, H=64, V=2, d_sh=d_sh), deg=deg)


def test_backward_colors_precomp_packed_cov_no_depth_grad():
    run(syn.make_small_scene(seed=31, P=700, W=64, H=48, V=2), use_sh=False, cov_packed=True, depth_grad=False)


def test_backward_pixel_aligned_scene():
    run(syn.make_scene(seed=5, v=2, V=2, hw=64), bg=(0.0, 0.0, 0.0))


def test_autograd_through_render_cuda():
    """Public API: gradients reach means / covariances / SH / opacities and the camera deltas."""
    import torch

    from styl3r_b200.decoder import render_cuda

    sc = syn.make_scene(seed=9, v=2, V=2, hw=64)
    t = lambda a: torch.as_tensor(a).cuda()
    means = t(sc["means"])[None].requires_grad_()
    cov = t(sc["covariances"])[None].requires_grad_()
    sh = t(sc["harmonics"])[None].requires_grad_()
    op = t(sc["opacities"])[None].requires_grad_()
    rot = torch.zeros(2, 3, device="cuda", requires_grad=True)
    trans = torch.zeros(2, 3, device="cuda", requires_grad=True)
    color, depth = render_cuda(t(sc["extrinsics"]), t(sc["intrinsics"]), t(sc["near"]), t(sc["far"]), (64, 64),
                               torch.zeros(2, 3, device="cuda"), means, cov, sh, op, cam_rot_delta=rot,
                               cam_trans_delta=trans, view_set=torch.zeros(2, dtype=torch.int32, device="cuda"))
    target = torch.rand_like(color)
    loss = ((color - target) ** 2).mean() + 0.01 * depth.mean()
    loss.backward()
    for name, x in [("means", means), ("cov", cov), ("sh", sh), ("op", op), ("rot", rot), ("trans", trans)]:
        assert x.grad is not None and torch.isfinite(x.grad).all(), name
        assert x.grad.abs().max() > 0, name
    # one Adam-like descent step on the pose must reduce the loss (sanity of the sign convention)
    with torch.no_grad():
        from styl3r_b200.pose import update_pose
        step = 1e-3
        new_c2w = update_pose(-step * trans.grad / (trans.grad.norm() + 1e-12), -step * rot.grad / (rot.grad.norm() + 1e-12),
                              t(sc["extrinsics"]))
        c2, d2 = render_cuda(new_c2w, t(sc["intrinsics"]), t(sc["near"]), t(sc["far"]), (64, 64),
                             torch.zeros(2, 3, device="cuda"), means, cov, sh, op,
                             view_set=torch.zeros(2, dtype=torch.int32, device="cuda"))
        loss2 = ((c2 - target) ** 2).mean() + 0.01 * d2.mean()
    assert loss2 < loss, (float(loss), float(loss2))


def test_backward_full_size_cfg2():
    """BASELINE cfg2 size: 131 072 Gaussians, one 256x256 view - every gradient (incl. dL/dtau) against the oracle."""
    run(syn.make_scene(seed=1234, v=2, V=1, hw=256), bg=(0.0, 0.0, 0.0), seed=3)


def test_only_requested_gradients_are_computed():
    """needs_input_grad is honoured: with only the SH coefficients requiring a gradient (stage-2
    training: frozen structure heads) the other gradient buffers are neither allocated nor accumulated, and dL/dSH
    equals the one of the full backward."""
    import torch
    from styl3r_b200.decoder import render_cuda
    sc = syn.make_scene(seed=9, v=2, V=2, hw=64)
    t = lambda a: torch.as_tensor(a).cuda()
    vs = torch.zeros(2, dtype=torch.int32, device="cuda")
    target = torch.rand(2, 3, 64, 64, device="cuda", generator=torch.Generator(device="cuda").manual_seed(0))

    def run(req):
        means, cov, sh, op = (t(sc[k])[None].requires_grad_(k in req) for k in ("means", "covariances", "harmonics", "opacities"))
        color, _ = render_cuda(t(sc["extrinsics"]), t(sc["intrinsics"]), t(sc["near"]), t(sc["far"]), (64, 64),
                               torch.zeros(2, 3, device="cuda"), means, cov, sh, op, view_set=vs)
        ((color - target) ** 2).mean().backward()
        return means.grad, cov.grad, sh.grad, op.grad

    full = run({"means", "covariances", "harmonics", "opacities"})
    only = run({"harmonics"})
    assert only[0] is None and only[1] is None and only[3] is None
    assert torch.allclose(only[2], full[2], rtol=1e-4, atol=1e-9) and only[2].abs().max() > 0


def test_backward_kernel_variants_agree():
    """The blend-backward kernel is specialised on the requested outputs (geometry only = pose-align loop, colour only =
    stage-2 training, all): every variant must reproduce the 10-value kernel (S3R_TUNE_BWD_ALL) on what it returns."""
    import torch

    from styl3r_b200 import _lib
    from styl3r_b200 import rasterizer as rz

    scene = syn.make_scene(seed=7, v=2, V=2, hw=64)
    outs, cams = oracle_scene(scene, render=False)
    *_, ctx = gpu_scene(scene, cams, want_n_touched=False)
    rng = np.random.default_rng(1)
    gc = torch.as_tensor(rng.normal(size=(2, 3, 64, 64)).astype(np.float32)).cuda()
    gd = torch.as_tensor((0.1 * rng.normal(size=(2, 64, 64))).astype(np.float32)).cuda()
    only_sh = dict(means=False, cov=False, opacities=False, shs=True, colors=False, means2D=False)

    def three():
        return (rz.backward_raw(ctx, gc, gd), rz.backward_raw(ctx, gc, gd, only_pose=True),
                rz.backward_raw(ctx, gc, gd, need_pose=False, needs=only_sh))

    try:
        _lib.lib().s3r_set_tunable(14, 1)
        ref_full, ref_pose, ref_sh = three()
    finally:
        _lib.lib().s3r_set_tunable(14, 0)
    full, pose, sh = three()
    torch.cuda.synchronize()
    close("tau (geometry-only kernel)", pose["tau"].cpu().numpy(), ref_pose["tau"].cpu().numpy(), rel=1e-4)
    close("tau (geometry-only vs full)", pose["tau"].cpu().numpy(), full["tau"].cpu().numpy(), rel=1e-4)
    close("shs (colour-only kernel)", sh["shs"].cpu().numpy(), ref_sh["shs"].cpu().numpy(), rel=1e-4)
    close("shs (colour-only vs full)", sh["shs"].cpu().numpy(), full["shs"].cpu().numpy(), rel=1e-4)
    for k in ("means", "cov", "opacities", "shs", "tau"):
        close(k, full[k].cpu().numpy(), ref_full[k].cpu().numpy(), rel=1e-5)
    assert pose["means"] is None and sh["means"] is None and sh["opacities"] is None


def test_tile_and_warp_granular_backward_agree():
    """S3R_TUNE_BLEND_KERNEL selects the (forward, backward) kernel pair: tile-granular (TMA ring + cull) or warp-granular
    over the per-block survivor lists (default).  Same gradients up to the summation order of the atomics."""
    import torch

    from styl3r_b200 import _lib
    from styl3r_b200 import rasterizer as rz

    scene = syn.make_scene(seed=11, v=2, V=2, hw=64)
    outs, cams = oracle_scene(scene, render=False)
    rng = np.random.default_rng(2)
    gc = torch.as_tensor(rng.normal(size=(2, 3, 64, 64)).astype(np.float32)).cuda()
    gd = torch.as_tensor((0.1 * rng.normal(size=(2, 64, 64))).astype(np.float32)).cuda()

    def run():
        *_, ctx = gpu_scene(scene, cams, want_n_touched=False)
        return rz.backward_raw(ctx, gc, gd)

    warp = run()
    try:
        _lib.lib().s3r_set_tunable(15, 1)
        tile = run()
    finally:
        _lib.lib().s3r_set_tunable(15, 0)
    torch.cuda.synchronize()
    for k in ("means", "cov", "opacities", "shs", "tau", "means2D"):
        close(k, warp[k].cpu().numpy(), tile[k].cpu().numpy(), rel=1e-4)
