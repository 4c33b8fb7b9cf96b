"""Row f2 (input staging): the numpy oracle of Pillow's 8-bit LANCZOS resize is pinned against golden vectors produced
by the REFERENCE's own rescale / rescale_and_crop / apply_style_image_augmentation running on Pillow
(tests/golden/make_staging_golden.py), and - when PIL is importable - against Pillow live.  Bit-exact."""
from pathlib import Path

import numpy as np
import pytest

from oracle import resize_oracle as ro

GOLD = np.load(Path(__file__).parent / "golden" / "staging_golden.npz")
dec = lambda codes: (codes / 255).astype(np.float32)
CASES = ("re10k", "portrait", "same_w", "odd")


@pytest.mark.parametrize("tag", CASES)
def test_oracle_rescale_and_crop_equals_reference_golden(tag):
    img, K = GOLD[f"{tag}_in"].astype(np.float32), GOLD[f"{tag}_K"]
    out, Kout = ro.rescale_and_crop(img, K, tuple(GOLD[f"{tag}_shape"]))
    assert np.array_equal(out, dec(GOLD[f"{tag}_out"]))
    assert np.array_equal(Kout, GOLD[f"{tag}_Kout"])


def test_oracle_upscale_style_and_normalise_equal_reference_golden():
    assert np.array_equal(ro.rescale(GOLD["up_in"].astype(np.float32), (50, 96)), dec(GOLD["up_out"]))
    sty = GOLD["style_in"].astype(np.float32)
    h, w = ro.style_shape(*sty.shape[1:])
    s = ro.rescale(sty, (h, w))
    r, c = round((h - 256) / 2.0), round((w - 256) / 2.0)
    assert np.array_equal(s[:, r:r + 256, c:c + 256], dec(GOLD["style_out"]))
    norm = (dec(GOLD["re10k_out"]) - np.float32(0.5)) / np.float32(0.5)
    assert np.array_equal(norm, GOLD["re10k_norm"])


def test_oracle_equals_pillow_live():
    Image = pytest.importorskip("PIL.Image")
    rng = np.random.default_rng(0)
    for (h, w, ho, wo) in [(37, 53, 20, 31), (64, 48, 64, 30), (50, 50, 81, 67), (9, 200, 5, 64)]:
        u8 = rng.integers(0, 256, (h, w, 3), dtype=np.uint8)
        ref = np.array(Image.fromarray(u8).resize((wo, ho), Image.LANCZOS))
        assert np.array_equal(ro.resize_lanczos_u8(u8, ho, wo), ref), (h, w, ho, wo)


def test_host_tap_tables_equal_oracle():
    from styl3r_b200.staging import lanczos_taps
    for (a, b) in [(160, 114), (90, 64), (64, 64), (40, 50), (200, 426), (77, 50), (640, 455), (360, 256), (3, 7)]:
        ob, ok, oks = ro.precompute_coeffs(a, b)
        hb, hk, hks = lanczos_taps(a, b)
        assert oks == hks and np.array_equal(ob, hb) and np.array_equal(ok, hk), (a, b)
    # fixed-point taps of every output pixel sum to 2^22 within rounding
    _, hk, _ = lanczos_taps(640, 455)
    assert np.abs(hk.sum(1) - (1 << 22)).max() <= 8
