"""Golden vectors for the camera set-up, the `render_cuda` call site and the pose-delta path, produced by the
REFERENCE's own Python, run on CPU in the build container:

  * `get_fov` (src/geometry/projection.py:247-261), `get_projection_matrix` (src/model/decoder/cuda_splatting.py:16-43);
  * `render_cuda` / `DecoderSplattingCUDA.forward` (cuda_splatting.py:46-133, decoder_splatting_cuda.py:37-68), imported
    UNMODIFIED, with a stand-in module named `diff_gaussian_rasterization` (the un-vendored third-party rasterizer)
    that (i) records the `GaussianRasterizationSettings` and tensors the reference hands over per view and (ii)
    renders them with the CPU oracle (oracle/raster_oracle.c, forward and backward) inside a torch.autograd.Function.
    Everything on the reference's side of that boundary - scale-invariant rescale, SH layout, fov / view / projection
    matrices, triu covariance gather, autograd through those torch ops, theta / rho plumbing - is therefore the
    reference's own code; only the rasterizer proper is the (parity-unpinned) oracle;
  * `SE3_exp`, `SO3_exp`, `V`, `update_pose` (src/misc/cam_utils.py:67-137);
  * the `test_step_align` loop (infer_model_re10k.py:79-161), executed from the reference source with the reference's
    `LossMse` (src/loss/loss_mse.py) on top of the same stand-in.

Writes tests/golden/camera_pose_golden.npz.   python tests/golden/make_camera_pose_golden.py
"""
import ast
import sys
import textwrap
import types
from pathlib import Path
from typing import NamedTuple

import numpy as np
import torch
from torch import nn

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from oracle import raster_oracle as ro  # noqa: E402
from styl3r_b200 import synthetic as syn  # noqa: E402
from tests.golden.make_encoder_golden import install_stubs  # noqa: E402

REF = Path("/root/reference")
CAPTURED = []  # one dict per GaussianRasterizer call


# ----------------------------------------------------------------------------- stand-in third-party module


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    projmatrix_raw: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


class _OracleRaster(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D, shs, colors_precomp, opacities, cov3D_precomp, theta, rho, s):
        n = lambda t: None if t is None else t.detach().cpu().numpy().astype(np.float32)
        fwd = ro.forward(n(means3D), n(cov3D_precomp), n(opacities).reshape(-1), n(s.viewmatrix).reshape(16),
                         n(s.projmatrix).reshape(16), n(s.campos), int(s.image_width), int(s.image_height),
                         np.float32(s.tanfovx), np.float32(s.tanfovy), n(s.bg), shs=n(shs), colors=n(colors_precomp),
                         deg=int(s.sh_degree))
        ctx.fwd, ctx.praw = fwd, n(s.projmatrix_raw).reshape(16)
        ctx.has = (shs is not None, theta is not None, rho is not None)
        ctx.opac_shape = opacities.shape
        CAPTURED.append(dict(
            tanfov=np.array([s.tanfovx, s.tanfovy], np.float64), viewmatrix=n(s.viewmatrix), projmatrix=n(s.projmatrix),
            projmatrix_raw=n(s.projmatrix_raw), campos=n(s.campos), bg=n(s.bg), sh_degree=int(s.sh_degree),
            hw=(int(s.image_height), int(s.image_width)), means3D=n(means3D), cov6=n(cov3D_precomp),
            opacities=n(opacities), shs=n(shs), scale_modifier=float(s.scale_modifier),
            prefiltered=bool(s.prefiltered), debug=bool(s.debug), R=int(fwd["R"])))
        t = torch.from_numpy
        radii, nt = t(fwd["radii"].copy()), t(fwd["n_touched"].copy())
        color, depth, opacity = t(fwd["color"].copy()), t(fwd["depth"].copy())[None], t(fwd["opacity"].copy())[None]
        ctx.mark_non_differentiable(radii, opacity, nt)
        return color, radii, depth, opacity, nt

    @staticmethod
    def backward(ctx, g_color, _gr, g_depth, _go, _gn):
        g = ro.backward(ctx.fwd, g_color.numpy(), None if g_depth is None else g_depth[0].numpy(), proj_raw=ctx.praw)
        has_sh, has_theta, has_rho = ctx.has
        t = torch.from_numpy
        P = ctx.fwd["inputs"]["means"].shape[0]
        m2d = torch.zeros(P, 3)
        m2d[:, :2] = t(g["dL_dmean2D"])
        tau = t(g["dL_dtau"])
        return (t(g["dL_dmeans"]), m2d, t(g["dL_dsh"]) if has_sh else None,
                None if has_sh else t(g["dL_dcolor"]), t(g["dL_dopacity"]).reshape(ctx.opac_shape), t(g["dL_dcov6"]),
                tau[3:].clone() if has_theta else None, tau[:3].clone() if has_rho else None, None)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings):
        super().__init__()
        self.raster_settings = raster_settings

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None, rotations=None,
                cov3D_precomp=None, theta=None, rho=None):
        assert scales is None and rotations is None and cov3D_precomp is not None
        return _OracleRaster.apply(means3D, means2D, shs, colors_precomp, opacities, cov3D_precomp, theta, rho,
                                   self.raster_settings)


def install_reference():
    install_stubs()
    mod = types.ModuleType("diff_gaussian_rasterization")
    mod.GaussianRasterizationSettings, mod.GaussianRasterizer = GaussianRasterizationSettings, GaussianRasterizer
    sys.modules["diff_gaussian_rasterization"] = mod
    if str(REF) not in sys.path:
        sys.path.insert(0, str(REF))


def ref_test_step_align():
    """The reference's test_step_align, compiled from its own source (the module around it needs hydra/wandb)."""
    src = (REF / "infer_model_re10k.py").read_text()
    from einops import rearrange
    from src.misc.cam_utils import update_pose
    ns = dict(torch=torch, nn=nn, rearrange=rearrange, update_pose=update_pose, tqdm=lambda it, **k: it)
    for node in ast.parse(src).body:
        if isinstance(node, ast.FunctionDef) and node.name == "test_step_align":
            exec(textwrap.dedent(ast.unparse(node)), ns)
    return ns["test_step_align"]


def scene_tensors(sc):
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a))
    from src.model.types import Gaussians
    g = Gaussians(t(sc["means"])[None].clone().requires_grad_(), t(sc["covariances"])[None].clone().requires_grad_(),
                  t(sc["harmonics"])[None].clone().requires_grad_(), t(sc["opacities"])[None].clone().requires_grad_())
    cams = (t(sc["extrinsics"])[None], t(sc["intrinsics"])[None], t(sc["near"])[None], t(sc["far"])[None])
    return g, cams


def main():
    install_reference()
    from src.geometry.projection import get_fov
    from src.loss.loss_mse import LossMse, LossMseCfg, LossMseCfgWrapper
    from src.misc.cam_utils import SE3_exp, update_pose
    from src.model.decoder.cuda_splatting import get_projection_matrix
    from src.model.decoder.decoder_splatting_cuda import DecoderSplattingCUDA, DecoderSplattingCUDACfg

    out = {}
    rng = np.random.default_rng(0)

    # ---- get_fov / get_projection_matrix
    K = np.tile(np.array([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1.0]], np.float32), (6, 1, 1))
    K[:, 0, 0] = [0.8, 0.52, 1.3, 0.9, 0.7, 2.0]
    K[:, 1, 1] = [0.8, 0.93, 1.1, 1.2, 0.4, 2.5]
    K[:, 0, 2] = [0.5, 0.5, 0.48, 0.55, 0.5, 0.5]
    K[:, 1, 2] = [0.5, 0.5, 0.52, 0.45, 0.5, 0.5]
    near = np.array([0.1, 0.5, 1.0, 0.01, 2.0, 1.0], np.float32)
    far = np.array([100.0, 50.0, 1000.0, 10.0, 2.5, 200.0], np.float32)
    fov = get_fov(torch.tensor(K))
    out["fov_K"], out["fov_out"] = K, fov.numpy()
    out["proj_near"], out["proj_far"] = near, far
    out["proj_out"] = get_projection_matrix(torch.tensor(near), torch.tensor(far), fov[:, 0], fov[:, 1]).numpy()

    # ---- render_cuda call site: reference decoder over the oracle-backed stand-in
    cases = {
        "si": (syn.make_scene(seed=21, v=2, V=3, hw=64), True, [0.0, 0.0, 0.0]),
        "raw": (syn.make_small_scene(seed=4, P=1500, W=80, H=48, V=2), False, [0.2, 0.5, 0.7]),
    }
    cases["si"][0]["near"][:] = [0.1, 0.25, 0.5]
    for tag, (sc, si, bg) in cases.items():
        g, (extr, intr, nr, fr) = scene_tensors(sc)
        b, v = extr.shape[:2]
        rot = torch.zeros(b, v, 3, requires_grad=True)
        trans = torch.zeros(b, v, 3, requires_grad=True)
        dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", bg, si))
        CAPTURED.clear()
        o = dec.forward(g, extr, intr, nr, fr, sc["image_shape"], cam_rot_delta=rot, cam_trans_delta=trans)
        wgen = torch.Generator().manual_seed(3)
        wc = torch.rand(o.color.shape, generator=wgen)
        wd = torch.rand(o.depth.shape, generator=wgen) * 0.1
        ((o.color * wc).sum() + (o.depth * wd).sum()).backward()
        for k in ("means", "covariances", "harmonics", "opacities", "extrinsics", "intrinsics", "near", "far"):
            out[f"{tag}_{k}"] = sc[k]
        out[f"{tag}_hw"] = np.array(sc["image_shape"])
        out[f"{tag}_bg"] = np.array(bg, np.float32)
        out[f"{tag}_color"], out[f"{tag}_depth"] = o.color.detach().numpy(), o.depth.detach().numpy()
        out[f"{tag}_wc"], out[f"{tag}_wd"] = wc.numpy(), wd.numpy()
        out[f"{tag}_g_means"], out[f"{tag}_g_cov"] = g.means.grad.numpy(), g.covariances.grad.numpy()
        out[f"{tag}_g_sh"], out[f"{tag}_g_opac"] = g.harmonics.grad.numpy(), g.opacities.grad.numpy()
        out[f"{tag}_g_rot"], out[f"{tag}_g_trans"] = rot.grad.numpy(), trans.grad.numpy()
        for k in ("tanfov", "viewmatrix", "projmatrix", "projmatrix_raw", "campos"):
            out[f"{tag}_cam_{k}"] = np.stack([c[k] for c in CAPTURED])
        out[f"{tag}_R"] = np.array([c["R"] for c in CAPTURED])
        # what the reference hands to the rasterizer for view 0 (scaled means / packed cov / sh layout), sampled
        out[f"{tag}_v0_means3D"] = CAPTURED[0]["means3D"][::37]
        out[f"{tag}_v0_cov6"] = CAPTURED[0]["cov6"][::37]
        out[f"{tag}_v0_shs"] = CAPTURED[0]["shs"][::37]
        assert CAPTURED[0]["scale_modifier"] == 1.0 and not CAPTURED[0]["prefiltered"] and not CAPTURED[0]["debug"]

    # ---- SE3_exp / update_pose
    tau = rng.normal(0, 0.2, (8, 6)).astype(np.float32)
    tau[1, 3:] = 1e-7          # small-angle branch of SO3_exp / V
    tau[2] = 0                 # identity
    tau[3, 3:] = [3e-6, -4e-6, 5e-6]
    tau[4, 3:] *= 10           # large rotation
    tau[5, :3] = 0
    sc = syn.make_small_scene(seed=1, P=4, V=8)
    c2w = sc["extrinsics"]
    out["pose_tau"], out["pose_c2w"] = tau, c2w
    out["pose_se3"] = np.stack([SE3_exp(torch.tensor(t_)).numpy() for t_ in tau])
    out["pose_new_c2w"] = update_pose(torch.tensor(tau[:, :3]), torch.tensor(tau[:, 3:]), torch.tensor(c2w)).numpy()

    # ---- test_step_align (reference loop) with LossMse on the oracle-backed decoder
    sc = syn.make_scene(seed=33, v=2, V=2, hw=64)
    g, (extr, intr, nr, fr) = scene_tensors(sc)
    dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True))
    with torch.no_grad():
        target = dec.forward(g, extr, intr, nr, fr, (64, 64)).color
    pert = extr.clone()
    pert[0, :, 0, 3] += torch.tensor([0.03, -0.02])
    pert[0, :, 1, 3] += torch.tensor([-0.01, 0.02])
    batch = {"target": {"image": target, "extrinsics": pert, "intrinsics": intr, "near": nr, "far": fr}}
    steps = 6
    cfg = types.SimpleNamespace(rot_opt_lr=0.005, trans_opt_lr=0.005, pose_align_steps=steps)
    align = ref_test_step_align()
    enc = nn.Linear(1, 1)
    gd = type(g)(g.means.detach(), g.covariances.detach(), g.harmonics.detach(), g.opacities.detach())
    history = []
    real_forward = dec.forward

    def spy(gaussians, extrinsics, *a, **k):
        history.append(extrinsics.detach().clone().numpy())
        return real_forward(gaussians, extrinsics, *a, **k)

    dec.forward = spy
    o, _ = align(cfg, enc, dec, [LossMse(LossMseCfgWrapper(LossMseCfg(1.0)))], batch, gd, gd, "cpu")
    for k in ("means", "covariances", "harmonics", "opacities", "intrinsics", "near", "far"):
        out[f"align_{k}"] = sc[k]
    out["align_target"], out["align_start"] = target.numpy(), pert.numpy()
    out["align_extrinsics_per_step"] = np.stack(history[:steps + 1])  # extrinsics used by forward i (i = steps: final)
    out["align_final_color"] = o.color.detach().numpy()
    out["align_steps"] = np.array(steps)
    out["align_lr"] = np.array([0.005, 0.005])

    path = ROOT / "tests" / "golden" / "camera_pose_golden.npz"
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({path.stat().st_size / 1e3:.0f} kB, {len(out)} arrays)")


if __name__ == "__main__":
    main()
