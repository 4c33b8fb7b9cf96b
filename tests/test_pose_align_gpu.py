"""GPU: on-device pose alignment (row f1) vs an eager restatement of the reference loop
(infer_model_re10k.py:79-161) built from the public pieces: render_cuda + torch.optim.Adam + update_pose."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _setup():
    import torch
    from types import SimpleNamespace
    from styl3r_b200 import synthetic as syn
    from styl3r_b200.decoder import render_cuda
    sc = syn.make_scene(seed=17, v=2, V=2, hw=64)
    t = lambda a: torch.as_tensor(a).cuda()
    g = SimpleNamespace(means=t(sc["means"])[None], covariances=t(sc["covariances"])[None],
                        harmonics=t(sc["harmonics"])[None], opacities=t(sc["opacities"])[None])
    extr, intr, near, far = t(sc["extrinsics"])[None], t(sc["intrinsics"])[None], t(sc["near"])[None], t(sc["far"])[None]
    vs = torch.zeros(2, dtype=torch.int32, device="cuda")
    with torch.no_grad():
        target, _ = render_cuda(extr[0], intr[0], near[0], far[0], (64, 64), torch.zeros(2, 3, device="cuda"), g.means,
                                g.covariances, g.harmonics, g.opacities, view_set=vs)
    # perturbed start pose
    pert = extr.clone()
    pert[0, :, 0, 3] += 0.03
    pert[0, :, 1, 3] -= 0.02
    return g, extr, pert, intr, near, far, target[None], vs


def test_pose_align_matches_eager_reference_loop_and_converges():
    import torch
    from styl3r_b200.decoder import render_cuda
    from styl3r_b200.pose import update_pose
    from styl3r_b200.pose_align import pose_align
    g, extr_gt, pert, intr, near, far, target, vs = _setup()
    steps, lr = 30, 0.003
    # ---- eager restatement of the reference loop
    rot = torch.nn.Parameter(torch.zeros(1, 2, 3, device="cuda"))
    trans = torch.nn.Parameter(torch.zeros(1, 2, 3, device="cuda"))
    opt = torch.optim.Adam([{"params": [rot], "lr": lr}, {"params": [trans], "lr": lr}])
    extr = pert.clone()
    hist = []
    for _ in range(steps):
        opt.zero_grad()
        color, _ = render_cuda(extr[0], intr[0], near[0], far[0], (64, 64), torch.zeros(2, 3, device="cuda"), g.means,
                               g.covariances, g.harmonics, g.opacities, cam_rot_delta=rot[0], cam_trans_delta=trans[0],
                               view_set=vs)
        loss = ((color - target[0]) ** 2).mean()
        hist.append(float(loss))
        loss.backward()
        with torch.no_grad():
            opt.step()
            extr = update_pose(cam_rot_delta=rot[0], cam_trans_delta=trans[0], extrinsics=extr[0])[None]
            rot.data.fill_(0)
            trans.data.fill_(0)
    # ---- device loop (CUDA graph)
    refined, losses = pose_align(g, pert, intr, near, far, (64, 64), target, steps=steps, rot_lr=lr, trans_lr=lr)
    losses = losses.cpu().numpy()
    assert losses[-1] < 0.5 * losses[0], losses  # converges
    np.testing.assert_allclose(losses[:5], hist[:5], rtol=2e-3)
    np.testing.assert_allclose(refined.cpu().numpy(), extr.cpu().numpy(), atol=2e-3)
    # pose moved toward the ground truth
    e0 = (pert - extr_gt)[0, :, :3, 3].norm()
    e1 = (refined - extr_gt)[0, :, :3, 3].norm()
    assert e1 < e0


def test_pose_align_graph_equals_eager_iterations():
    import torch
    from styl3r_b200.pose_align import pose_align
    g, _, pert, intr, near, far, target, _ = _setup()
    a, la = pose_align(g, pert, intr, near, far, (64, 64), target, steps=8, use_graph=True)
    b, lb = pose_align(g, pert, intr, near, far, (64, 64), target, steps=8, use_graph=False)
    np.testing.assert_allclose(la.cpu().numpy(), lb.cpu().numpy(), rtol=1e-3)
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), atol=1e-4)


def test_pose_align_takes_the_configured_loss_list():
    """the reference sums every configured loss.  A list [mse] must equal the default; [mse, 0.5 * l1] must
    run through autograd inside the captured iteration and still converge."""
    import torch
    from styl3r_b200.pose_align import pose_align
    g, _, pert, intr, near, far, target, _ = _setup()
    mse = lambda c, t: ((c - t) ** 2).mean()
    l1 = lambda c, t: 0.5 * (c - t).abs().mean()
    a, la = pose_align(g, pert, intr, near, far, (64, 64), target, steps=6)
    b, lb = pose_align(g, pert, intr, near, far, (64, 64), target, steps=6, losses=[mse])
    np.testing.assert_allclose(a.cpu().numpy(), b.cpu().numpy(), atol=1e-5)
    c, lc = pose_align(g, pert, intr, near, far, (64, 64), target, steps=25, rot_lr=0.003, trans_lr=0.003, losses=[mse, l1])
    lc = lc.cpu().numpy()
    assert lc[-1] < 0.6 * lc[0]
    with pytest.raises(ValueError):
        pose_align(g, pert, intr, near, far, (64, 64), target, steps=2, losses=[mse], loss_grad=lambda c, t: c - t)
