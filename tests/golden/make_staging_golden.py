"""Generates tests/golden/staging_golden.npz by running the REFERENCE's own staging code (which calls Pillow):
src/dataset/shims/crop_shim.py rescale / rescale_and_crop, src/dataset/shims/normalize_shim.py normalize_image and
src/dataset/shims/augmentation_shim.py apply_style_image_augmentation, on small seeded images.
Run in the build container (needs /root/reference and PIL):  python tests/golden/make_staging_golden.py"""
import sys
import types
from pathlib import Path

import numpy as np
import torch

REF = Path("/root/reference")


def load_shims():
    # import the three shim files as a tiny package so that their relative `..types` import resolves to a stub
    pkg = types.ModuleType("refds"); pkg.__path__ = []
    shims = types.ModuleType("refds.shims"); shims.__path__ = []
    tmod = types.ModuleType("refds.types")
    tmod.AnyExample = tmod.AnyViews = tmod.BatchedExample = dict
    sys.modules.update({"refds": pkg, "refds.shims": shims, "refds.types": tmod})
    out = {}
    for name in ("crop_shim", "normalize_shim", "augmentation_shim"):
        src = (REF / "src/dataset/shims" / f"{name}.py").read_text()
        mod = types.ModuleType(f"refds.shims.{name}")
        mod.__package__ = "refds.shims"
        sys.modules[mod.__name__] = mod
        exec(compile(src, name, "exec"), mod.__dict__)
        out[name] = mod
    return out


def main():
    m = load_shims()
    crop, norm, aug = m["crop_shim"], m["normalize_shim"], m["augmentation_shim"]
    g = torch.Generator().manual_seed(7)
    out = {}

    def image(n, h, w):  # smooth-ish content + noise so that the filter taps matter, values slightly outside [0,1]
        yy, xx = torch.meshgrid(torch.linspace(0, 6.3, h), torch.linspace(0, 9.1, w), indexing="ij")
        base = torch.stack([0.5 + 0.5 * torch.sin(yy + c) * torch.cos(xx * (c + 1)) for c in range(3)])
        return (base[None] + 0.15 * torch.randn(n, 3, h, w, generator=g)).half().float()  # fp16-exact: stored as fp16

    def u8(t):  # outputs are exactly k/255 in float32: store the codes (image = (codes / 255).astype(float32))
        codes = torch.round(t * 255).to(torch.uint8)
        assert torch.equal((codes.double() / 255).float(), t.float())
        return codes.numpy()

    cases = {"re10k": (2, 90, 160, (64, 64)), "portrait": (1, 150, 100, (48, 80)), "same_w": (1, 100, 64, (64, 64)),
             "odd": (2, 77, 123, (50, 70))}
    K = torch.tensor([[0.6, 0, 0.5], [0, 0.9, 0.5], [0, 0, 1]])
    for tag, (n, h, w, shape) in cases.items():
        img = image(n, h, w)
        Ks = K[None].repeat(n, 1, 1)
        o_img, o_K = crop.rescale_and_crop(img, Ks, shape)
        out[f"{tag}_in"], out[f"{tag}_K"], out[f"{tag}_shape"] = img.half().numpy(), Ks.numpy(), np.array(shape)
        out[f"{tag}_out"], out[f"{tag}_Kout"] = u8(o_img), o_K.numpy()
        if tag == "re10k":
            out[f"{tag}_norm"] = norm.normalize_image(o_img).numpy()
    up = image(1, 40, 60)[0]
    out["up_in"], out["up_out"] = up.half().numpy(), u8(crop.rescale(up, (50, 96)))        # enlarging
    sty = image(1, 120, 200)[0].clip(0, 1)
    out["style_in"] = sty.half().numpy()
    out["style_out"] = u8(aug.apply_style_image_augmentation(sty, "val"))
    dst = Path(__file__).parent / "staging_golden.npz"
    np.savez_compressed(dst, **out)
    print(dst, {k: v.shape for k, v in out.items() if k.endswith("_out")})


if __name__ == "__main__":
    main()
