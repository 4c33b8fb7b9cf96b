"""Host logic of the output side against golden vectors produced by the REFERENCE's own code
(tests/golden/make_outputs_golden.py: src/visualization/camera_trajectory/interpolation.py, src/model/ply_export.py)."""
from pathlib import Path

import numpy as np
import torch

GOLD = np.load(Path(__file__).parent / "golden" / "outputs_golden.npz")


def test_interpolate_extrinsics_matches_reference():
    from styl3r_b200.trajectory import interpolate_extrinsics
    t = torch.tensor(GOLD["traj_t"])
    for name in ("converging", "parallel", "twisted", "wrap"):
        a, b = torch.tensor(GOLD[f"traj_{name}_initial"]), torch.tensor(GOLD[f"traj_{name}_final"])
        out = interpolate_extrinsics(a, b, t)
        assert out.dtype == torch.float32 and out.shape == (t.numel(), 4, 4)
        assert np.abs(out.numpy() - GOLD[f"traj_{name}_extrinsics"]).max() <= 1e-6, name
        # end points reproduce the inputs; rotations stay orthonormal
        assert np.abs(out[0].numpy() - a.numpy()).max() <= 1e-5 and np.abs(out[-1].numpy() - b.numpy()).max() <= 1e-5
        R = out[:, :3, :3].double()
        assert (R @ R.transpose(1, 2) - torch.eye(3, dtype=torch.float64)).abs().max() <= 1e-6


def test_interpolate_extrinsics_batched_equals_unbatched():
    from styl3r_b200.trajectory import interpolate_extrinsics
    t = torch.tensor(GOLD["traj_t"])
    names = ("converging", "parallel", "twisted", "wrap")
    a = torch.stack([torch.tensor(GOLD[f"traj_{n}_initial"]) for n in names])
    b = torch.stack([torch.tensor(GOLD[f"traj_{n}_final"]) for n in names])
    out = interpolate_extrinsics(a, b, t)
    assert out.shape == (4, t.numel(), 4, 4)
    for i, n in enumerate(names):
        assert np.abs(out[i].numpy() - GOLD[f"traj_{n}_extrinsics"]).max() <= 1e-6, n


def test_interpolate_intrinsics_and_frame_times():
    from styl3r_b200.trajectory import interpolate_intrinsics, smooth_time
    t = torch.tensor(GOLD["traj_t"])
    out = interpolate_intrinsics(torch.tensor(GOLD["intr_initial"]), torch.tensor(GOLD["intr_final"]), t)
    assert np.array_equal(out.numpy(), GOLD["intr_out"])
    assert np.allclose(smooth_time(7).numpy(), GOLD["traj_t"], atol=1e-7)
    assert torch.equal(smooth_time(5, smooth=False), torch.linspace(0, 1, 5))


def test_ply_attribute_names_match_reference():
    from styl3r_b200.ply_export import construct_list_of_attributes
    assert construct_list_of_attributes(0) == list(GOLD["ply_plain_names"])
    assert construct_list_of_attributes(9) == list(GOLD["ply_rest_names"])


def test_wobble_trajectories_match_reference_module():
    """generate_wobble / generate_wobble_transformation vs the reference file executed in place when available
    (src/visualization/camera_trajectory/wobble.py), else vs their closed form."""
    import importlib.util
    from styl3r_b200.trajectory import generate_wobble, generate_wobble_transformation
    t = torch.linspace(0, 1, 9)
    radius = torch.tensor([0.25, 0.5])
    extr = torch.eye(4).repeat(2, 1, 1)
    extr[:, :3, 3] = torch.tensor([[0.1, 0.2, 0.3], [-1.0, 0.5, 2.0]])
    tf = generate_wobble_transformation(radius, t, 5, scale_radius_with_t=False)
    assert tf.shape == (2, 9, 4, 4)
    assert torch.allclose(tf[1, :, 0, 3], torch.sin(2 * torch.pi * 5 * t) * 0.5) and torch.allclose(tf[0, :, 1, 3], -torch.cos(2 * torch.pi * 5 * t) * 0.25)
    wob = generate_wobble(extr, radius, t)
    assert torch.allclose(wob[:, 0, :3, 3], extr[:, :3, 3])           # radius scales with t: starts at the camera
    ref_path = Path("/root/reference/src/visualization/camera_trajectory/wobble.py")
    if ref_path.exists():                                              # build container: compare with the reference itself
        spec = importlib.util.spec_from_file_location("ref_wobble", ref_path)
        ref = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(ref)
        assert torch.equal(ref.generate_wobble(extr, radius, t), wob)
        assert torch.equal(ref.generate_wobble_transformation(radius, t, 5, scale_radius_with_t=False), tf)
