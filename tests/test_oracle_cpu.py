"""CPU tests (no GPU): the oracle against its fixtures and against torch autograd; RoPE oracle against the
reference's own golden vectors."""
import hashlib
from pathlib import Path

import numpy as np
import pytest
import torch

from oracle import raster_oracle as ro
from oracle import torch_mirror as tm
from styl3r_b200 import synthetic as syn
from tests.helpers import oracle_scene

GOLD = Path(__file__).parent / "golden"


def test_raster_oracle_matches_regression_fixture():
    g = np.load(GOLD / "raster_small.npz")
    sc = syn.make_small_scene(seed=42, P=400, W=48, H=32, V=2, d_sh=4)
    outs, _ = oracle_scene(sc, deg=1, bg=(0.1, 0.2, 0.3))
    for v, o in enumerate(outs):
        np.testing.assert_array_equal(o["radii"], g[f"v{v}_radii"])
        np.testing.assert_array_equal(o["ranges"], g[f"v{v}_ranges"])
        np.testing.assert_array_equal(o["point_list"], g[f"v{v}_point_list"])
        assert hashlib.sha256(o["keys"].tobytes()).hexdigest() == str(g[f"v{v}_keys_sha256"])
        np.testing.assert_array_equal(o["n_contrib"], g[f"v{v}_n_contrib"])
        np.testing.assert_allclose(o["color"], g[f"v{v}_color"], atol=1e-6)
        np.testing.assert_allclose(o["depth"], g[f"v{v}_depth"], atol=1e-5)


def test_raster_oracle_structure_invariants():
    """Size-independent properties: keys sorted, ranges partition the list, every instance lies in its rect."""
    sc = syn.make_scene(seed=3, v=2, V=1, hw=64)
    (o,), _ = oracle_scene(sc)
    keys = o["keys"]
    assert np.all(keys[1:] >= keys[:-1])
    assert o["R"] == int(o["tiles_touched"].sum()) == len(o["point_list"])
    tiles = (keys >> np.uint64(32)).astype(np.int64)
    gx = 64 // 16
    r = o["rects"][o["point_list"]]
    tx, ty = tiles % gx, tiles // gx
    assert np.all((tx >= r[:, 0]) & (tx < r[:, 2]) & (ty >= r[:, 1]) & (ty < r[:, 3]))
    nonempty = o["ranges"][:, 1] > o["ranges"][:, 0]
    assert o["ranges"][nonempty, 1].max() == o["R"]
    # stable sort: equal keys keep ascending Gaussian index
    same = keys[1:] == keys[:-1]
    assert np.all(o["point_list"][1:][same] > o["point_list"][:-1][same])
    # compositing: opacity = 1 - final_T, colour bounded
    np.testing.assert_allclose(o["opacity"], 1 - o["final_T"], atol=1e-6)
    assert o["color"].min() >= 0 and o["color"].max() <= 1.5


@pytest.mark.parametrize("deg,d_sh", [(0, 1), (2, 9)])
def test_oracle_backward_matches_autograd(deg, d_sh):
    """The analytic backward (incl. the se(3) pose gradient) equals torch autograd of the float64 mirror."""
    W, H = 48, 32
    sc = syn.make_small_scene(seed=3 + deg, P=160, W=W, H=H, V=1, d_sh=d_sh)
    cam = ro.camera_setup(sc["extrinsics"][0], sc["intrinsics"][0], sc["near"][0], sc["far"][0], True)
    m, c = ro.scale_gaussians(sc["means"], sc["covariances"], cam["scale"])
    c6 = ro.cov3x3_to_6(c)
    shs = np.ascontiguousarray(sc["harmonics"].transpose(0, 2, 1))
    bg = np.array([0.1, 0.2, 0.3], np.float32)
    f = ro.forward(m, c6, sc["opacities"], cam["view16"], cam["proj16"], cam["campos"], W, H, cam["tanx"], cam["tany"],
                   bg, shs=shs, deg=deg)
    rng = np.random.default_rng(0)
    gc = rng.normal(size=(3, H, W)).astype(np.float32)
    gd = (0.1 * rng.normal(size=(H, W))).astype(np.float32)
    g = ro.backward(f, gc, gd, cam["projraw16"])
    dt = torch.float64
    tmeans = torch.tensor(m, dtype=dt, requires_grad=True)
    tc6 = torch.tensor(c6, dtype=dt, requires_grad=True)
    tsh = torch.tensor(shs, dtype=dt, requires_grad=True)
    top = torch.tensor(sc["opacities"], dtype=dt, requires_grad=True)
    tau = torch.zeros(6, dtype=dt, requires_grad=True)
    w2c = torch.tensor(cam["view16"].reshape(4, 4).T.copy(), dtype=dt)
    praw = torch.tensor(cam["projraw16"].reshape(4, 4).T.copy(), dtype=dt)
    col, dep = tm.render(tmeans, tc6, top, w2c, praw, float(cam["tanx"]), float(cam["tany"]), W, H, bg,
                         f["point_list"], f["ranges"], shs=tsh, deg=deg, tau=tau)
    assert np.abs(col.detach().numpy() - f["color"]).max() < 1e-5
    ((col * torch.tensor(gc, dtype=dt)).sum() + (dep * torch.tensor(gd, dtype=dt)).sum()).backward()
    for name, mine, ref in [("means", g["dL_dmeans"], tmeans.grad), ("cov6", g["dL_dcov6"], tc6.grad),
                            ("sh", g["dL_dsh"], tsh.grad), ("opac", g["dL_dopacity"], top.grad),
                            ("tau", g["dL_dtau"], tau.grad)]:
        ref = ref.numpy()
        assert np.abs(mine - ref).max() <= 2e-4 * (np.abs(ref).max() + 1e-9), name


def test_rope_oracle_matches_reference_golden():
    g = np.load(GOLD / "rope2d_golden.npz")
    tok = np.ascontiguousarray(g["tokens_bhnd"].transpose(0, 2, 1, 3))  # [B,N,H,D]
    out = ro.rope2d(tok, g["positions"], base=float(g["base"]), fwd=1.0)
    # reference PyTorch RoPE2D (pos_embed.py:142-159)
    np.testing.assert_allclose(out, g["out_bhnd"].transpose(0, 2, 1, 3), atol=1e-5)
    # reference compiled rope_2d_cpu (curope.cpp:11-47)
    if "out_ref_cpu_bnhd" in g:
        np.testing.assert_allclose(out, g["out_ref_cpu_bnhd"], atol=1e-5)
    back = ro.rope2d(out, g["positions"], base=float(g["base"]), fwd=-1.0)
    np.testing.assert_allclose(back, tok, atol=2e-6)


def test_rope_oracle_against_compiled_reference_if_present():
    from oracle import build_ref
    mod = build_ref.load()
    if mod is None:
        pytest.skip("oracle/_ref not built in this checkout")
    torch.manual_seed(1)
    tok = torch.randn(2, 9, 4, 32)
    pos = torch.randint(0, 40, (2, 9, 2))
    ref = tok.clone()
    mod.rope_2d(ref, pos, 100.0, 1.0)
    np.testing.assert_allclose(ro.rope2d(tok.numpy(), pos.numpy(), 100.0, 1.0), ref.numpy(), atol=1e-5)


def test_camera_setup_matches_torch_restating_reference_ops():
    """oracle.camera_setup vs the product's torch camera set-up (same math as cuda_splatting.py:65-88)."""
    from styl3r_b200.decoder import cuda_splatting as cs
    sc = syn.make_small_scene(seed=2, P=10, V=3)
    e, k = torch.tensor(sc["extrinsics"]), torch.tensor(sc["intrinsics"])
    near, far = torch.tensor(sc["near"]), torch.tensor(sc["far"])
    scale = 1 / near
    e2 = e.clone()
    e2[:, :3, 3] *= scale[:, None]
    fov = cs.get_fov(k)
    proj = cs.get_projection_matrix(near * scale, far * scale, fov[:, 0], fov[:, 1]).transpose(1, 2)
    view = e2.inverse().transpose(1, 2)
    full = view @ proj
    for v in range(3):
        cam = ro.camera_setup(sc["extrinsics"][v], sc["intrinsics"][v], sc["near"][v], sc["far"][v], True)
        np.testing.assert_allclose(view[v].reshape(16).numpy(), cam["view16"], atol=2e-6)
        np.testing.assert_allclose(full[v].reshape(16).numpy(), cam["proj16"], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose((0.5 * fov[v]).tan().numpy(), [cam["tanx"], cam["tany"]], rtol=2e-6)


def test_raster_oracle_domain_properties():
    """Size-independent properties of the restated algorithm (the same ones the GPU path is checked for at full size):
    colour linearity, invariance under a permutation of the Gaussians (up to equal-depth ties, which keep input order),
    empty / all-culled input."""
    sc = syn.make_scene(seed=5, v=2, V=1, hw=64)
    (o,), _ = oracle_scene(sc, use_sh=False)
    # linearity: precomputed colours scaled by k (no clamp, zero background) scale the image by k, leave T / depth alone
    sc2 = dict(sc, harmonics=sc["harmonics"] * 0.5)
    (o2,), _ = oracle_scene(sc2, use_sh=False)
    np.testing.assert_allclose(o2["color"], 0.5 * o["color"], atol=2e-6)
    np.testing.assert_array_equal(o2["final_T"], o["final_T"])
    np.testing.assert_array_equal(o2["depth"], o["depth"])
    # permutation of the input order: same image (ties in depth are measure-zero for this scene), same per-tile counts
    perm = np.random.default_rng(0).permutation(sc["means"].shape[0])
    sc3 = dict(sc, means=sc["means"][perm], covariances=sc["covariances"][perm], harmonics=sc["harmonics"][perm],
               opacities=sc["opacities"][perm])
    (o3,), _ = oracle_scene(sc3, use_sh=False)
    np.testing.assert_array_equal(o3["ranges"], o["ranges"])
    mapped = perm[o3["point_list"]]
    tie = np.zeros(len(mapped), bool)                      # equal (tile, depth) keys keep input order: permutation-dependent
    tie[1:] |= o["keys"][1:] == o["keys"][:-1]
    tie[:-1] |= o["keys"][1:] == o["keys"][:-1]
    np.testing.assert_array_equal(mapped[~tie], o["point_list"][~tie])
    assert tie.mean() < 0.01
    for a, b in o["ranges"][::7]:                           # per tile the same set of Gaussians
        np.testing.assert_array_equal(np.sort(mapped[a:b]), np.sort(o["point_list"][a:b]))
    assert np.abs(o3["color"] - o["color"]).max() <= 2e-2 and np.abs(o3["color"] - o["color"]).mean() <= 1e-5
    # everything behind the camera: nothing is binned, the image is the background
    sc4 = dict(sc, means=sc["means"] * np.array([1, 1, -1], np.float32))
    (o4,), _ = oracle_scene(sc4, use_sh=False, bg=(0.2, 0.4, 0.6))
    assert o4["R"] == 0 and int(o4["radii"].max()) == 0 and int(o4["ranges"].max()) == 0
    np.testing.assert_allclose(o4["color"], np.broadcast_to(np.array([0.2, 0.4, 0.6], np.float32)[:, None, None], o4["color"].shape))
    assert float(o4["opacity"].max()) == 0.0
