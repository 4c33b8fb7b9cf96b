"""CPU, build container only (needs /root/reference): the UNMODIFIED reference file
src/model/decoder/cuda_splatting.py (`render_cuda`, :46-133) and decoder_splatting_cuda.py import and run on top of OUR
drop-in module `styl3r_b200.compat.diff_gaussian_rasterization` registered by `compat.install()`.

There is no GPU where the reference sources exist and no reference where the GPU is, so the shim's compute call
(`styl3r_b200.rasterizer.rasterize`, CUDA only) is replaced here by the CPU oracle — as the CHECKER of what the shim
forwards: every keyword, tensor layout and settings field the reference hands to `GaussianRasterizer` must arrive at
the rasterizer entry unchanged, and the image must equal the golden produced by the same reference file over the
oracle stand-in (tests/golden/make_camera_pose_golden.py).  The GPU side of the same contract is
tests/test_compat_gpu.py (shim == batched path) and tests/test_camera_pose_gpu.py (batched path == these goldens)."""
import sys
from pathlib import Path

import numpy as np
import pytest

REF = Path("/root/reference")
pytestmark = pytest.mark.skipif(not (REF / "src/model/decoder/cuda_splatting.py").exists(),
                                reason="reference sources are only present in the build container")
G = np.load(Path(__file__).parent / "golden" / "camera_pose_golden.npz")


def test_unmodified_reference_render_cuda_runs_on_the_shim(monkeypatch):
    import torch
    from oracle import raster_oracle as ro
    from tests.golden.make_encoder_golden import install_stubs
    install_stubs()
    for m in [k for k in sys.modules if k == "diff_gaussian_rasterization" or k.startswith("src.model.decoder")]:
        monkeypatch.delitem(sys.modules, m)
    from styl3r_b200 import compat, rasterizer as rz
    from styl3r_b200.compat import diff_gaussian_rasterization as shim
    monkeypatch.setitem(sys.modules, "diff_gaussian_rasterization", shim)
    compat.install()
    assert sys.modules["diff_gaussian_rasterization"] is shim
    calls = []

    def oracle_rasterize(means, cov, opacities, *, shs=None, colors_precomp=None, rho=None, theta=None, means2D=None,
                         **cfg):
        """Stands where the CUDA entry is: same signature as styl3r_b200.rasterizer.rasterize, one view."""
        assert means.shape[0] == 1 and cfg["viewmatrix"].shape == (1, 4, 4)
        n = lambda x: None if x is None else x.detach().numpy().astype(np.float32)
        o = ro.forward(n(means[0]), n(cov[0]), n(opacities[0]), n(cfg["viewmatrix"][0]).reshape(16),
                       n(cfg["projmatrix"][0]).reshape(16), n(cfg["campos"][0]), cfg["W"], cfg["H"],
                       float(cfg["tanfov"][0, 0]), float(cfg["tanfov"][0, 1]), n(cfg["background"][0]),
                       shs=None if shs is None else n(shs[0]), colors=None if colors_precomp is None else n(colors_precomp[0]),
                       deg=cfg["sh_degree"])
        calls.append(dict(cfg=cfg, rho=rho, theta=theta, means2D=means2D, R=o["R"]))
        f = torch.from_numpy
        return (f(o["color"])[None], f(o["depth"])[None], f(o["opacity"])[None], f(o["radii"])[None],
                f(o["n_touched"])[None])

    monkeypatch.setattr(rz, "rasterize", oracle_rasterize)
    sys.path.insert(0, str(REF))
    try:
        from src.model.decoder.decoder_splatting_cuda import DecoderSplattingCUDA, DecoderSplattingCUDACfg
        from src.model.types import Gaussians
    finally:
        sys.path.remove(str(REF))
    import src.model.decoder.cuda_splatting as ref_cs
    assert ref_cs.GaussianRasterizer is shim.GaussianRasterizer  # the reference file bound OUR classes
    for tag, si in (("si", True), ("raw", False)):
        calls.clear()
        t = lambda k: torch.as_tensor(G[f"{tag}_{k}"])[None]
        g = Gaussians(t("means"), t("covariances"), t("harmonics"), t("opacities"))
        V = G[f"{tag}_extrinsics"].shape[0]
        rot, trans = torch.zeros(1, V, 3), torch.zeros(1, V, 3)
        dec = DecoderSplattingCUDA(DecoderSplattingCUDACfg("splatting_cuda", G[f"{tag}_bg"].tolist(), si))
        h, w = (int(x) for x in G[f"{tag}_hw"])
        with torch.no_grad():
            out = dec.forward(g, t("extrinsics"), t("intrinsics"), t("near"), t("far"), (h, w), cam_rot_delta=rot,
                              cam_trans_delta=trans)
        np.testing.assert_array_equal(out.color.numpy(), G[f"{tag}_color"])
        np.testing.assert_array_equal(out.depth.numpy(), G[f"{tag}_depth"])
        assert len(calls) == V and [c["R"] for c in calls] == G[f"{tag}_R"].tolist()
        for v, c in enumerate(calls):
            np.testing.assert_array_equal(c["cfg"]["projmatrix_raw"][0].numpy(), G[f"{tag}_cam_projmatrix_raw"][v])
            assert c["rho"].shape == (1, 3) and c["theta"].shape == (1, 3) and c["means2D"].shape == (g.means.shape[1], 3)
            assert c["cfg"]["want_n_touched"] is True
