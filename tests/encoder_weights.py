"""Deterministic model weights / inputs shared by the golden generator (reference model, CPU) and the GPU parity
test (our model): every parameter is a function of its NAME only."""
import zlib

import torch


def named_tensor(name: str, shape, std: float, mean: float = 0.0) -> torch.Tensor:
    g = torch.Generator().manual_seed(zlib.crc32(name.encode()) & 0x7FFFFFFF)
    return torch.randn(tuple(shape), generator=g, dtype=torch.float32) * std + mean


def fill_named_weights(module: torch.nn.Module) -> None:
    """LayerNorm weights ~ 1 +- 0.02, biases ~ 0.02, matrices / conv kernels ~ N(0, 1/fan_in) (unit-gain)."""
    seen = {}
    with torch.no_grad():
        for name, p in module.state_dict().items():
            if p.data_ptr() in seen:  # aliased parameters (scratch.layer_rn.N == scratch.layerN_rn) keep one value
                continue
            seen[p.data_ptr()] = name
            if p.dim() >= 2:
                fan_in = p[0].numel() if p.dim() > 1 else p.numel()
                if "act_postprocess" in name and p.dim() == 4 and name.endswith("1.weight") and p.shape[0] == p.shape[1] and p.shape[-1] in (2, 4):
                    fan_in = p.shape[0]  # ConvTranspose2d [in, out, k, k] with k == stride: one tap per output pixel
                gain = 0.2 if name.endswith("dpt.head.4.weight") else 1.0  # keep expm1(|xyz|) well conditioned
                p.copy_(named_tensor(name, p.shape, gain * fan_in ** -0.5).to(p.device))
            elif "norm" in name and name.endswith("weight"):
                p.copy_(named_tensor(name, p.shape, 0.02, 1.0).to(p.device))
            else:
                p.copy_(named_tensor(name, p.shape, 0.02).to(p.device))


def make_inputs(b: int, v: int, hw: int, seed: int, device="cpu"):
    g = torch.Generator().manual_seed(seed)
    img = torch.rand(b, v, 3, hw, hw, generator=g) * 2 - 1
    sty = torch.rand(b, 3, hw, hw, generator=g) * 2 - 1
    K = torch.tensor([[0.8, 0.0, 0.5], [0.0, 0.8, 0.5], [0.0, 0.0, 1.0]]).expand(b, v, 3, 3).contiguous()
    return {"image": img.to(device), "intrinsics": K.to(device)}, {"image": sty.to(device)}


def fill_vgg_named(named_tensors) -> None:
    """VGG-19 slices (src/test/vgg_model.py VGGEncoder): He-scaled kernels / small biases as a function of the name
    (`slice2.5.weight`, ...).  Accepts named_parameters() or named_buffers() (the reference's losses convert the VGG
    parameters to buffers)."""
    with torch.no_grad():
        for name, t in named_tensors:
            if t.dim() == 4:
                t.copy_(named_tensor("vgg." + name, t.shape, (2.0 / t[0].numel()) ** 0.5).to(t.device))
            else:
                t.copy_(named_tensor("vgg." + name, t.shape, 0.02).to(t.device))
