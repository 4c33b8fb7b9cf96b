"""Golden values for the style / identity losses from the REFERENCE's own code (src/loss/loss_style.py,
src/loss/loss_identity.py, src/test/vgg_model.py), run on CPU in the build container.  `vgg19(pretrained=True)` would
download ImageNet weights (no network): it is patched to build the same architecture with weights that are a
deterministic function of the parameter name (tests/encoder_weights.py), which the test re-creates.
python tests/golden/make_loss_golden.py"""
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
from tests.golden.make_encoder_golden import REF, install_stubs  # noqa: E402
from tests.encoder_weights import fill_vgg_named  # noqa: E402


def make_batch(seed=0, b=2, v=2, hw=32):
    g = torch.Generator().manual_seed(seed)
    pred = torch.rand(b, v, 3, hw, hw, generator=g)
    batch = {"target": {"image": torch.rand(b, v, 3, hw, hw, generator=g)}, "style": {"image": torch.rand(b, 3, hw, hw, generator=g)}}
    return pred, batch


def main():
    install_stubs()
    import types
    dgr = types.ModuleType("diff_gaussian_rasterization")   # imported by src.model.decoder at module load; unused here
    dgr.GaussianRasterizationSettings = dgr.GaussianRasterizer = object
    sys.modules["diff_gaussian_rasterization"] = dgr
    sys.path.insert(0, str(REF))
    import torchvision.models as tvm
    real_vgg19 = tvm.vgg19
    tvm.vgg19 = lambda pretrained=False, **kw: real_vgg19(weights=None)
    import src.test.vgg_model as vm
    vm.vgg19 = tvm.vgg19
    from src.loss.loss_identity import IdentityLoss
    from src.loss.loss_style import LossStyle, LossStyleCfg, LossStyleCfgWrapper
    from src.model.decoder.decoder import DecoderOutput

    def named(loss):  # the reference converts the VGG parameters to non-persistent buffers: fill them by name
        fill_vgg_named(loss.vgg.named_buffers())
        return loss

    out = {}
    style = named(LossStyle(LossStyleCfgWrapper(LossStyleCfg(10.0))))
    ident = named(IdentityLoss())
    pred, batch = make_batch()
    pred.requires_grad_(True)
    ls = style(DecoderOutput(pred, None), batch, None, 0)
    ls.backward()
    out["style_loss"], out["style_grad"] = ls.detach().numpy(), pred.grad.numpy().copy()
    pred.grad = None
    li = ident(DecoderOutput(pred, None), batch, None, 0)
    li.backward()
    out["identity_loss"], out["identity_grad"] = li.detach().numpy(), pred.grad.numpy().copy()
    out["vgg_buffer_names"] = np.array([n for n, _ in style.vgg.named_buffers()])
    dst = Path(__file__).parent / "loss_golden.npz"
    np.savez_compressed(dst, **out)
    print(dst, float(ls), float(li), list(out["vgg_buffer_names"])[:4])


if __name__ == "__main__":
    main()
