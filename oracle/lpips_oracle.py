"""CPU oracle of the reference's perceptual loss: LPIPS v0.1 with the VGG16 backbone and linear heads
(src/custom_lpips/networks_basic.py:27-91 ``PNetLin``; backbone slices src/custom_lpips/pretrained_networks.py:97-135;
scaling layer networks_basic.py:93-100; ``normalize_tensor`` custom_lpips/__init__.py:42-44; call site src/utils.py:16,44-50,
``percept(target, est)`` -> ``DistModel.forward(in0=target... )`` dist_model.py:107-116).

TEST INFRASTRUCTURE ONLY (see oracle/__init__.py).  Plain torch-CPU ops; every function cites what it restates.

Pinning: tests/golden/make_golden_lpips.py imports the reference's own ``networks_basic.PNetLin`` (with ``pnet_rand=True``:
torchvision's VGG16 architecture, seeded random weights - the ImageNet weights cannot be downloaded here - and the
reference's shipped linear heads weights/v0.1/vgg.pth) and records its outputs and image gradients in
tests/golden/lpips.npz; tests/test_oracle_golden.py checks this restatement against them.  The BACKBONE WEIGHTS are
therefore unpinned (random), the arithmetic is pinned.
"""
from __future__ import annotations

from typing import Dict, List, Sequence

import torch
import torch.nn.functional as F

# torchvision vgg16().features indices of the 13 convolutions, grouped by the reference's five slices
# (pretrained_networks.py:108-117: slice1 = features[0:4], slice2 = [4:9], slice3 = [9:16], slice4 = [16:23], slice5 = [23:30]);
# a slice after the first starts with the 2x2 max-pool (features[4], [9], [16], [23])
VGG_SLICES = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21), (24, 26, 28))
VGG_CHANNELS = (64, 128, 256, 512, 512)
SHIFT = (-.030, -.088, -.188)     # networks_basic.py:96
SCALE = (.458, .448, .450)        # networks_basic.py:97


def vgg_param_names() -> List[str]:
    """state_dict names of PNetLin's backbone + heads (what the native plan consumes)."""
    names = []
    for si, convs in enumerate(VGG_SLICES):
        for idx in convs:
            names += [f"net.slice{si + 1}.{idx}.weight", f"net.slice{si + 1}.{idx}.bias"]
    names += [f"lin{k}.model.1.weight" for k in range(5)]
    return names


def scaling_layer(x):
    """(inp - shift) / scale per RGB channel (networks_basic.py:93-100)."""
    shift = torch.tensor(SHIFT, dtype=x.dtype)[None, :, None, None]
    scale = torch.tensor(SCALE, dtype=x.dtype)[None, :, None, None]
    return (x - shift) / scale


def vgg_features(params: Dict[str, torch.Tensor], x) -> List[torch.Tensor]:
    """relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 (pretrained_networks.py:119-133)."""
    outs = []
    h = x
    for si, convs in enumerate(VGG_SLICES):
        if si > 0:
            h = F.max_pool2d(h, 2, 2)
        for idx in convs:
            h = F.relu(F.conv2d(h, params[f"net.slice{si + 1}.{idx}.weight"], params[f"net.slice{si + 1}.{idx}.bias"], padding=1))
        outs.append(h)
    return outs


def normalize_tensor(f, eps: float = 1e-10):
    """custom_lpips/__init__.py:42-44."""
    return f / (torch.sqrt(torch.sum(f * f, dim=1, keepdim=True)) + eps)


def lpips_from_features(params, feats0: Sequence[torch.Tensor], feats1: Sequence[torch.Tensor]):
    """Sum over the five taps of spatial_average(lin_k((n0 - n1)^2)) (networks_basic.py:68-88; Dropout is the identity in
    eval mode, dist_model.py:96).  Returns [N, 1, 1, 1]."""
    val = None
    for k in range(5):
        d = (normalize_tensor(feats0[k]) - normalize_tensor(feats1[k])) ** 2
        r = F.conv2d(d, params[f"lin{k}.model.1.weight"]).mean([2, 3], keepdim=True)
        val = r if val is None else val + r
    return val


def lpips(params, in0, in1):
    """PNetLin.forward(in0, in1), version 0.1 (networks_basic.py:63-91)."""
    return lpips_from_features(params, vgg_features(params, scaling_layer(in0)), vgg_features(params, scaling_layer(in1)))


def perceptual_loss(params, target, est):
    """``get_loss(target, est, 'perceptual')`` (src/utils.py:44-50): ``percept(img1, img2)`` -> PerceptualLoss.forward(pred,
    target) calls ``self.model.forward(target, pred)`` (custom_lpips/__init__.py:27-40), i.e. in0 = est... the metric is
    symmetric in its arguments, so the order does not matter numerically."""
    return lpips(params, target, est)


def make_vgg_params(seed: int = 0, lin_weights: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Seeded random PNetLin parameters: the shared fixture builder lives in tests/fixtures.py (bench.py uses it too)."""
    import fixtures
    return fixtures.make_vgg_params(seed, lin_weights)
