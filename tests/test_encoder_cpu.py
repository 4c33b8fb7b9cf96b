"""CPU: the encoder oracle (oracle/encoder_oracle.py = our module tree evaluated with plain torch ops on the host)
against golden vectors produced by the REFERENCE encoder (tests/golden/make_encoder_golden.py).  This pins, without a
GPU, (i) the parameter registry / state-dict contract, (ii) the op sequence of backbone, token stylizer, DPT heads and
adapter, (iii) the RoPE-2D restatement.  Tolerance: fp32 on both sides, different reduction orders -> max |err| <=
2e-3 * std, mean |err| <= 2e-4 * std (the tolerance of the GPU parity test)."""
from pathlib import Path

import numpy as np
import torch

GOLD = Path(__file__).parent / "golden"


def _sample(t, n=4096):
    f = t.detach().reshape(-1)
    idx = torch.linspace(0, f.numel() - 1, min(n, f.numel())).long()
    return f[idx].float().numpy()


def test_encoder_oracle_matches_reference_golden():
    from oracle.encoder_oracle import encoder_forward
    from styl3r_b200.encoder import EncoderNoPoSplatTokenStyleCfg, get_encoder
    from tests.encoder_weights import fill_named_weights, make_inputs
    torch.set_num_threads(max(1, torch.get_num_threads()))
    enc, _ = get_encoder(EncoderNoPoSplatTokenStyleCfg(stylized=True))
    fill_named_weights(enc)
    enc = enc.eval()
    g = np.load(GOLD / "encoder_golden.npz")
    context, style = make_inputs(1, 2, 256, seed=1234, device="cpu")
    means, cov, harm, opac, scales, rots = encoder_forward(enc, context, style)
    assert means.shape == (1, 2 * 65536, 3) and cov.shape == (1, 2 * 65536, 3, 3)
    for name, t in [("means", means), ("covariances", cov), ("harmonics", harm), ("opacities", opac), ("scales", scales),
                    ("rotations", rots)]:
        ref, scale = g[f"b1v2_{name}"], float(g[f"b1v2_{name}_stats"][2])
        err = np.abs(_sample(t) - ref)
        assert err.max() <= 2e-3 * scale and err.mean() <= 2e-4 * scale, \
            f"{name}: max {err.max():.3e} mean {err.mean():.3e} scale {scale:.3e}"


def test_product_encoder_refuses_cpu_tensors():
    """The product path has no CPU fallback: the oracle above is test infrastructure, not a code path of the encoder."""
    import pytest
    from styl3r_b200 import _lib
    from styl3r_b200.encoder.encoder import EncoderNoPoSplatMultiTokenStyle
    enc = EncoderNoPoSplatMultiTokenStyle.__new__(EncoderNoPoSplatMultiTokenStyle)
    with pytest.raises(_lib.S3RError):
        EncoderNoPoSplatMultiTokenStyle.forward(enc, {"image": torch.zeros(1, 2, 3, 256, 256)}, {"image": torch.zeros(1, 3, 256, 256)})
