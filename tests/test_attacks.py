"""attacks.py against the reference's own attack modules (tests/golden/attacks.npz from tests/golden/make_golden_attacks.py:
Gaussian_blur and the DiffJPEG-based Jpeg at test-time settings).  JPEG is torch-only and is checked on CPU; the blur goes
through the native FIR op and needs the GPU.  JPEG quantisation rounds: a DCT coefficient that sits on a .5 boundary can land
on the other side under a different summation order, which changes one 8x8 block by one quantisation step, so the JPEG
bar is 'all but a handful of blocks agree to 1e-4'."""
import os

import numpy as np
import pytest
import torch

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "attacks.npz")


def gold():
    with np.load(GOLD) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("q", [50, 90, 20])
def test_jpeg_matches_reference_diffjpeg(q):
    import attacks
    g = gold()
    out = attacks.Jpeg(q)(torch.from_numpy(g["img"])).numpy()
    ref = g[f"jpeg_{q}"]
    bad = np.abs(out - ref) > 1e-4
    # per 8x8 block: how many blocks differ at all
    blocks = bad.reshape(2, 3, 8, 8, 8, 8).any(axis=(1, 3, 5)).mean()
    assert blocks <= 0.02, blocks
    assert np.abs(out - ref).max() <= 0.5
    assert attacks.quality_to_factor(50) == 1.0 and abs(attacks.quality_to_factor(20) - 2.5) < 1e-12


def test_noise_and_initializer_semantics():
    import attacks
    g = torch.Generator().manual_seed(3)
    img = torch.from_numpy(gold()["img"])
    out = attacks.GaussianNoise(0.1, g)(img)
    assert out.dtype == torch.float32 and float(out.abs().max()) <= 1.0
    d = out - img
    assert 0.08 < float(d[img.abs() < 0.8].std()) < 0.12       # N(0, 0.1) away from the clamp
    with pytest.raises(ValueError, match="Not available Attacks"):
        attacks.attack_initializer("None")
    assert isinstance(attacks.attack_initializer("Combination"), attacks.Combination)


@pytest.mark.gpu
@pytest.mark.parametrize("sigma", [0.5, 2.0])
def test_blur_matches_reference_gaussian_blur(sigma):
    import attacks
    g = gold()
    out = attacks.GaussianBlur(sigma)(torch.from_numpy(g["img"]).cuda()).cpu().numpy()
    np.testing.assert_allclose(out, g[f"blur_{sigma}"], rtol=1e-5, atol=2e-6)


@pytest.mark.gpu
def test_combination_runs_on_gpu_and_stays_in_range():
    import attacks
    img = torch.from_numpy(gold()["img"]).cuda()
    out = attacks.attack_initializer("Combination")(img)
    assert out.shape == img.shape and torch.isfinite(out).all() and float(out.abs().max()) <= 1.0 + 1e-6
