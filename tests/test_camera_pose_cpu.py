"""CPU: the oracle's camera / pose restatements are PINNED on golden vectors produced by the reference's own Python
(tests/golden/make_camera_pose_golden.py: get_fov, get_projection_matrix, the render_cuda set-up block, SE3_exp,
update_pose — SURVEY.md §8 rows a12, a15)."""
from pathlib import Path

import numpy as np

from oracle import pose_oracle as po
from oracle import raster_oracle as ro

G = np.load(Path(__file__).parent / "golden" / "camera_pose_golden.npz")


def test_oracle_get_fov_and_projection_match_reference():
    for i in range(len(G["fov_K"])):
        fx, fy = ro.get_fov(G["fov_K"][i])
        np.testing.assert_allclose([fx, fy], G["fov_out"][i], rtol=3e-7, atol=0)
        np.testing.assert_allclose(ro.projection_matrix(G["proj_near"][i], G["proj_far"][i], fx, fy), G["proj_out"][i],
                                   rtol=3e-7, atol=1e-9)


def test_oracle_camera_setup_matches_what_reference_render_cuda_hands_to_the_rasterizer():
    for tag, si in (("si", True), ("raw", False)):
        V = G[f"{tag}_extrinsics"].shape[0]
        for v in range(V):
            cam = ro.camera_setup(G[f"{tag}_extrinsics"][v], G[f"{tag}_intrinsics"][v], G[f"{tag}_near"][v],
                                  G[f"{tag}_far"][v], si)
            # torch's batched LU inverse vs numpy's fp64 inverse rounded once: a few fp32 ulps on O(1) entries
            np.testing.assert_allclose(cam["view16"], G[f"{tag}_cam_viewmatrix"][v].reshape(16), rtol=2e-6, atol=2e-6)
            np.testing.assert_allclose(cam["projraw16"], G[f"{tag}_cam_projmatrix_raw"][v].reshape(16), rtol=1e-6, atol=1e-8)
            np.testing.assert_allclose(cam["proj16"], G[f"{tag}_cam_projmatrix"][v].reshape(16), rtol=4e-6, atol=4e-6)
            np.testing.assert_allclose(cam["campos"], G[f"{tag}_cam_campos"][v], rtol=1e-7, atol=0)
            np.testing.assert_allclose([cam["tanx"], cam["tany"]], G[f"{tag}_cam_tanfov"][v], rtol=3e-7)
        # scaled means / packed covariance / SH layout handed to the rasterizer for view 0
        cam = ro.camera_setup(G[f"{tag}_extrinsics"][0], G[f"{tag}_intrinsics"][0], G[f"{tag}_near"][0],
                              G[f"{tag}_far"][0], si)
        m, c = ro.scale_gaussians(G[f"{tag}_means"], G[f"{tag}_covariances"], cam["scale"])
        np.testing.assert_array_equal(m[::37], G[f"{tag}_v0_means3D"])
        np.testing.assert_array_equal(ro.cov3x3_to_6(c)[::37], G[f"{tag}_v0_cov6"])
        np.testing.assert_array_equal(G[f"{tag}_harmonics"].transpose(0, 2, 1)[::37], G[f"{tag}_v0_shs"])


def test_pose_oracle_matches_reference_se3_exp_and_update_pose():
    tau, c2w = G["pose_tau"], G["pose_c2w"]
    for i in range(len(tau)):
        np.testing.assert_allclose(po.se3_exp(tau[i]), G["pose_se3"][i], rtol=0, atol=3e-7)
    np.testing.assert_allclose(po.update_pose(tau[:, :3], tau[:, 3:], c2w), G["pose_new_c2w"], rtol=0, atol=3e-6)


def test_reference_pose_align_golden_is_a_descent():
    """Sanity of the golden itself: the reference loop moved the perturbed cameras towards the true ones."""
    e = G["align_extrinsics_per_step"]
    assert e.shape[0] == int(G["align_steps"]) + 1
    assert np.abs(e[-1] - e[0]).max() > 1e-3
