"""Golden vectors for the encoder from the REFERENCE implementation, run on CPU in the build container.

Imports /root/reference/src unmodified under stubs for the 7 Python modules absent from this image:
xformers.ops (memory_efficient_attention -> F.scaled_dot_product_attention), dacite, lightning, skvideo, matplotlib,
omegaconf, e3nn.  curope is not built there, so the reference's own PyTorch RoPE2D is used.  Weights are a
deterministic function of the parameter NAME (tests/encoder_weights.py), so the GPU test can rebuild exactly the same
model without a 4 GB checkpoint.  Writes
  tests/golden/encoder_state_manifest.json   name -> shape of the reference state_dict (the checkpoint contract)
  tests/golden/encoder_golden.npz            inputs' seeds + strided samples of every output tensor
python tests/golden/make_encoder_golden.py
"""
import json
import sys
import types
from pathlib import Path

import numpy as np
import torch
import torch.nn.functional as F

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.encoder_weights import fill_named_weights, make_inputs  # noqa: E402

REF = Path("/root/reference")


class _Dummy(types.ModuleType):
    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        d = _Dummy(f"{self.__name__}.{name}")
        setattr(self, name, d)
        return d

    def __call__(self, *a, **k):
        return _Dummy("call")

    def __mro_entries__(self, bases):
        return (object,)

    def __getitem__(self, k):
        return self


def install_stubs():
    def mea(q, k, v, attn_bias=None, p=0.0, scale=None):
        return F.scaled_dot_product_attention(q.transpose(1, 2), k.transpose(1, 2), v.transpose(1, 2), scale=scale).transpose(1, 2)

    xf, xfo = types.ModuleType("xformers"), types.ModuleType("xformers.ops")
    xfo.memory_efficient_attention = mea
    xf.ops = xfo
    sys.modules.update({"xformers": xf, "xformers.ops": xfo})
    for name in ["dacite", "lightning", "lightning.pytorch", "lightning.pytorch.utilities", "lightning.pytorch.loggers",
                 "lightning.pytorch.loggers.wandb", "skvideo", "skvideo.io", "matplotlib", "matplotlib.figure",
                 "matplotlib.pyplot", "omegaconf", "e3nn", "e3nn.o3", "lpips", "plyfile", "moviepy", "moviepy.editor",
                 "colorspacious", "svg"]:
        if name not in sys.modules:
            try:
                __import__(name)
            except Exception:
                sys.modules[name] = _Dummy(name)


def build_reference_encoder(sh_degree=0):
    install_stubs()
    sys.path.insert(0, str(REF))
    from src.model.encoder.backbone.backbone_croco import BackboneCrocoCfg
    from src.model.encoder.common.gaussian_adapter import GaussianAdapterCfg
    from src.model.encoder.encoder_noposplat_multi_token_style import EncoderNoPoSplatMultiTokenStyle, OpacityMappingCfg
    from src.model.encoder.encoder_noposplat_token_style import EncoderNoPoSplatTokenStyleCfg
    from src.model.encoder.token_stylizer.token_stylizer import TokenStylizerCfg
    cfg = EncoderNoPoSplatTokenStyleCfg(
        name="noposplat_multi_token_style", d_feature=128, num_monocular_samples=32,
        backbone=BackboneCrocoCfg(name="croco_multi", model="ViTLarge_BaseDecoder", intrinsics_embed_loc="encoder",
                                  intrinsics_embed_degree=4, intrinsics_embed_type="token"),
        token_stylizer=TokenStylizerCfg(model="ViTLarge_BaseDecoder"), structure_builder=None, visualizer=None,
        gaussian_adapter=GaussianAdapterCfg(0.5, 15.0, sh_degree), apply_bounds_shim=True,
        opacity_mapping=OpacityMappingCfg(0.0, 0.0, 1), gaussians_per_pixel=1, num_surfaces=1,
        gs_params_head_type="dpt_gs", gs_sh_head_type="dpt", pose_free=True, stylized=True)
    return EncoderNoPoSplatMultiTokenStyle(cfg)


def sample(t, n=4096):
    f = t.detach().reshape(-1).double()
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].float().numpy(), np.array([float(f.abs().mean()), float(f.mean()), float(f.std())], np.float64)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(8)
    enc = build_reference_encoder(0).eval()
    sd = enc.state_dict()
    manifest = {k: list(v.shape) for k, v in sd.items()}
    (ROOT / "tests/golden/encoder_state_manifest.json").write_text(json.dumps(manifest, indent=0))
    fill_named_weights(enc)
    out = {}
    for tag, (b, v) in {"b1v2": (1, 2), "b1v3": (1, 3)}.items():
        context, style = make_inputs(b, v, 256, seed=1234)
        dump = {}
        with torch.no_grad():
            g = enc(context, style, visualization_dump=dump)
        for name, t in [("means", g.means), ("covariances", g.covariances), ("harmonics", g.harmonics),
                        ("opacities", g.opacities), ("scales", dump["scales"]), ("rotations", dump["rotations"])]:
            s, st = sample(t)
            out[f"{tag}_{name}"], out[f"{tag}_{name}_stats"] = s, st
            print(tag, name, tuple(t.shape), st)
    np.savez_compressed(ROOT / "tests/golden/encoder_golden.npz", **out)


if __name__ == "__main__":
    main()
