"""Host-side restatement of Pillow's 8-bit LANCZOS resampler (the coefficient tables the device resize kernel
consumes) against Pillow itself, bit for bit.  CPU only."""
import numpy as np
import pytest
from PIL import Image

from g2vlm_b200.host_prep import lanczos_tables


def _resample_axis(img, bounds, coef, axis):
    a = np.moveaxis(img, axis, 0).astype(np.int64)
    out = np.empty((bounds.shape[0],) + a.shape[1:], np.uint8)
    for i in range(bounds.shape[0]):
        lo, n = bounds[i]
        acc = (1 << 21) + np.tensordot(coef[i, :n].astype(np.int64), a[lo:lo + n], axes=(0, 0))
        out[i] = np.clip(acc >> 22, 0, 255).astype(np.uint8)
    return np.moveaxis(out, 0, axis)


@pytest.mark.parametrize("H,W,th,tw", [(270, 480, 154, 266), (120, 90, 196, 140), (97, 518, 70, 518), (64, 64, 64, 28)])
def test_tables_reproduce_pillow_lanczos(H, W, th, tw):
    rng = np.random.default_rng(H * 1000 + W)
    img = rng.integers(0, 256, (H, W, 3), dtype=np.uint8)
    img[: H // 2] = (np.linspace(0, 255, W)[None, :, None] + rng.integers(0, 3, (H // 2, W, 3))).clip(0, 255).astype(np.uint8)
    ref = np.asarray(Image.fromarray(img).resize((tw, th), Image.Resampling.LANCZOS))
    got = img
    if tw != W:
        got = _resample_axis(got, *lanczos_tables(W, tw), axis=1)     # Pillow: horizontal pass first, 8-bit intermediate
    if th != H:
        got = _resample_axis(got, *lanczos_tables(H, th), axis=0)
    assert np.array_equal(got, ref)


def test_tables_shape_and_normalisation():
    b, c = lanczos_tables(1920, 518)
    assert b.shape == (518, 2) and c.shape[0] == 518 and c.dtype == np.int32
    assert (b[:, 0] >= 0).all() and (b[:, 0] + b[:, 1] <= 1920).all()
    assert np.abs(c.sum(1) - (1 << 22)).max() <= c.shape[1]          # weights sum to 1 in fixed point (rounding)
