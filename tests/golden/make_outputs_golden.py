"""Generates tests/golden/outputs_golden.npz from the REFERENCE's own output-side code:

  * camera trajectories: interpolate_extrinsics / interpolate_intrinsics
    (/root/reference/src/visualization/camera_trajectory/interpolation.py:8-258, scipy Euler round trips included);
  * .ply export: export_ply (/root/reference/src/model/ply_export.py:26-74).  `plyfile` is not installed in this
    image, so a stub module captures the structured vertex array export_ply hands to PlyElement.describe - the golden
    is that array (the 17 float32 attributes per Gaussian), for shift_and_scale False and True.

Run in the build container (needs /root/reference):  python tests/golden/make_outputs_golden.py
"""
import importlib.util
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference/src")


def load(path: Path, name: str):
    spec = importlib.util.spec_from_file_location(name, path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


def pose(rx, ry, rz, t):
    from scipy.spatial.transform import Rotation as R
    m = np.eye(4)
    m[:3, :3] = R.from_euler("xyz", [rx, ry, rz]).as_matrix()
    m[:3, 3] = t
    return torch.tensor(m, dtype=torch.float32)


def main():
    out = {}
    interp = load(REF / "visualization/camera_trajectory/interpolation.py", "ref_interpolation")
    t = torch.linspace(0, 1, 7)
    t = (torch.cos(torch.pi * (t + 1)) + 1) / 2
    pairs = {
        "converging": (pose(0.02, -0.25, 0.01, [-0.5, 0.0, 0.0]), pose(-0.03, 0.3, 0.02, [0.6, 0.05, 0.1])),
        "parallel": (pose(0, 0, 0, [0.0, 0.0, 0.0]), pose(0, 0, 0, [1.0, 0.0, 0.0])),       # identical look vectors
        "twisted": (pose(0.1, 0.4, 0.3, [0.0, 0.2, -0.1]), pose(-0.2, -0.5, -0.4, [0.3, -0.1, 0.4])),
        "wrap": (pose(0.0, 3.0, 0.0, [0.0, 0.0, 1.0]), pose(0.0, -3.0, 0.0, [0.2, 0.0, -1.0])),  # angles straddle +-pi
    }
    for name, (a, b) in pairs.items():
        out[f"traj_{name}_initial"] = a.numpy()
        out[f"traj_{name}_final"] = b.numpy()
        out[f"traj_{name}_extrinsics"] = interp.interpolate_extrinsics(a, b, t).numpy()
    out["traj_t"] = t.numpy()
    Ka = torch.tensor([[0.8, 0, 0.5], [0, 0.8, 0.5], [0, 0, 1]])
    Kb = torch.tensor([[0.9, 0, 0.48], [0, 0.85, 0.52], [0, 0, 1]])
    out["intr_initial"], out["intr_final"] = Ka.numpy(), Kb.numpy()
    out["intr_out"] = interp.interpolate_intrinsics(Ka, Kb, t).numpy()

    # ---- export_ply with a capturing plyfile stub
    captured = {}
    ply = types.ModuleType("plyfile")

    class PlyElement:
        @staticmethod
        def describe(elements, name):
            captured["elements"], captured["name"] = elements.copy(), name
            return ("element", name)

    class PlyData:
        def __init__(self, elements):
            self.elements = elements

        def write(self, path):
            captured["path"] = str(path)

    ply.PlyElement, ply.PlyData = PlyElement, PlyData
    sys.modules["plyfile"] = ply
    pe = load(REF / "model/ply_export.py", "ref_ply_export")
    g = torch.Generator().manual_seed(5)
    n = 257
    means = torch.randn(n, 3, generator=g) * 2 + torch.tensor([0.3, -0.2, 4.0])
    scales = torch.rand(n, 3, generator=g) * 0.05 + 1e-4
    rot = torch.randn(n, 4, generator=g)
    rot = rot / rot.norm(dim=-1, keepdim=True)          # xyzw, unit (the adapter's output, gaussian_adapter.py:139)
    rot[0] = torch.tensor([0.0, 0.0, 0.0, 1.0])
    rot[1] = torch.tensor([0.0, 0.0, 0.0, -1.0])         # w < 0: the scipy round trip flips the sign
    rot[2] = torch.tensor([1.0, 0.0, 0.0, 0.0])
    rot[3] = torch.tensor([0.5, 0.5, 0.5, -0.5])
    harm = torch.randn(n, 3, 4, generator=g)
    opac = torch.rand(n, generator=g)
    for tag, sas in (("plain", False), ("shift", True)):
        pe.export_ply(means, scales, rot, harm, opac, Path("/tmp/_s3r_golden/x.ply"), shift_and_scale=sas,
                      save_sh_dc_only=True)
        el = captured["elements"]
        out[f"ply_{tag}"] = np.stack([el[name] for name in el.dtype.names], axis=-1).astype(np.float32)
        out[f"ply_{tag}_names"] = np.array(el.dtype.names)
    pe.export_ply(means, scales, rot, harm, opac, Path("/tmp/_s3r_golden/x.ply"), save_sh_dc_only=False)
    el = captured["elements"]
    out["ply_rest"] = np.stack([el[name] for name in el.dtype.names], axis=-1).astype(np.float32)
    out["ply_rest_names"] = np.array(el.dtype.names)
    out.update(ply_means=means.numpy(), ply_scales=scales.numpy(), ply_rot=rot.numpy(), ply_harm=harm.numpy(),
               ply_opac=opac.numpy())
    dst = Path(__file__).parent / "outputs_golden.npz"
    np.savez_compressed(dst, **out)
    print(dst, {k: v.shape for k, v in out.items() if k.startswith(("traj_conv", "ply_plain", "ply_rest"))})


if __name__ == "__main__":
    main()
