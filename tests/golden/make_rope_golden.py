"""Generates tests/golden/rope2d_golden.npz from the REFERENCE's own PyTorch RoPE2D
(/root/reference/src/model/encoder/backbone/croco/pos_embed.py:112-159; the class the reference falls back to when
curope is not compiled) and, if oracle/_ref was built, from the reference's compiled rope_2d_cpu
(curope.cpp:11-47).  Run in the build container (needs /root/reference):  python tests/golden/make_rope_golden.py
"""
import importlib.util
import sys
from pathlib import Path

import numpy as np
import torch

ROOT = Path(__file__).resolve().parents[2]
sys.path.insert(0, str(ROOT))
REF = Path("/root/reference/src/model/encoder/backbone/croco/pos_embed.py")


def load_ref_rope():
    src = REF.read_text()
    # the module does a relative import of the compiled extension first; it is absent, so the PyTorch class is used
    spec = importlib.util.spec_from_loader("ref_pos_embed", loader=None)
    mod = importlib.util.module_from_spec(spec)
    mod.__package__ = "ref_pos_embed_pkg"
    exec(compile(src, str(REF), "exec"), mod.__dict__)
    return mod.RoPE2D


def main():
    RoPE2D = load_ref_rope()
    torch.manual_seed(0)
    B, H, D = 1, 2, 64
    ys, xs = torch.meshgrid(torch.arange(16), torch.arange(16), indexing="ij")
    pos = torch.stack([ys.flatten(), xs.flatten()], -1)
    pos = torch.cat([pos, torch.tensor([[16, 0]])], 0)  # + intrinsics token at (16, 0) like backbone_croco_multiview
    pos = pos[None].repeat(B, 1, 1).contiguous()
    N = pos.shape[1]
    tokens = torch.randn(B, H, N, D)
    rope = RoPE2D(freq=100.0, F0=1.0)
    out = rope(tokens.clone(), pos)
    d = dict(tokens_bhnd=tokens.numpy(), positions=pos.numpy(), out_bhnd=out.numpy(), base=np.float32(100.0))
    from oracle import build_ref
    mod = build_ref.load()
    if mod is not None:
        t = tokens.transpose(1, 2).contiguous().clone()  # [B,N,H,D]
        mod.rope_2d(t, pos, 100.0, 1.0)
        d["out_ref_cpu_bnhd"] = t.numpy()
    np.savez_compressed(ROOT / "tests/golden/rope2d_golden.npz", **d)
    print({k: getattr(v, "shape", v) for k, v in d.items()})


if __name__ == "__main__":
    main()
