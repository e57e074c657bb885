"""Seeded fixture builders shared by the tests and by tests/golden/make_golden.py.

All randomness comes from ``numpy.random.RandomState`` (bit-stable across numpy and
torch versions), so the golden vectors generated in the build container can be
re-derived from nothing but a seed on the GPU box.
"""
from __future__ import annotations

import math
from typing import Dict, List

import numpy as np
import torch

CHANNELS = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64, 512: 32, 1024: 16}


def channels_for(res: int, cm: int = 2) -> int:
    return CHANNELS[res] if res <= 32 else CHANNELS[res] * cm


def _randn(rs, *shape, scale=1.0):
    return torch.from_numpy((rs.standard_normal(shape) * scale).astype(np.float32))


def make_params(size: int, seed: int = 0, cm: int = 2, style_dim: int = 512, n_mlp: int = 8,
                lr_mlp: float = 0.01, perturb: bool = True) -> Dict[str, torch.Tensor]:
    """A ``Generator(size, style_dim, n_mlp, cm)`` state dict under the reference's names.

    Distributions follow the reference initialisers (src/model.py:98-100, :138-141, :215-219,
    :323, :308, :377); ``perturb`` overwrites the zero-initialised noise weights and biases
    with seeded non-zero values so those code paths are exercised (SURVEY.md 8c caveat).
    """
    rs = np.random.RandomState(seed)
    p: Dict[str, torch.Tensor] = {}
    z = 0.1 if perturb else 0.0
    for i in range(1, n_mlp + 1):
        p[f"style.{i}.weight"] = _randn(rs, style_dim, style_dim) / lr_mlp
        p[f"style.{i}.bias"] = _randn(rs, style_dim, scale=z * 10)
    p["input.input"] = _randn(rs, 1, channels_for(4, cm), 4, 4)

    def conv(prefix, cin, cout, k):
        p[f"{prefix}.conv.weight"] = _randn(rs, 1, cout, cin, k, k)
        p[f"{prefix}.conv.modulation.weight"] = _randn(rs, cin, style_dim)
        p[f"{prefix}.conv.modulation.bias"] = torch.ones(cin) + _randn(rs, cin, scale=z)

    def styled(prefix, cin, cout):
        conv(prefix, cin, cout, 3)
        p[f"{prefix}.noise.weight"] = _randn(rs, 1, scale=z) + (0.05 if perturb else 0.0)
        p[f"{prefix}.activate.bias"] = _randn(rs, cout, scale=z)

    def rgb(prefix, cin):
        conv(prefix, cin, 3, 1)
        p[f"{prefix}.bias"] = _randn(rs, 1, 3, 1, 1, scale=z)

    c = channels_for(4, cm)
    styled("conv1", c, c)
    rgb("to_rgb1", c)
    log_size = int(math.log2(size))
    cin = c
    for j, i in enumerate(range(3, log_size + 1)):
        cout = channels_for(2 ** i, cm)
        styled(f"convs.{2 * j}", cin, cout)
        styled(f"convs.{2 * j + 1}", cout, cout)
        rgb(f"to_rgbs.{j}", cout)
        cin = cout
    return p


def make_noise(size: int, seed: int = 1, batch: int = 1) -> List[torch.Tensor]:
    """Noise maps ``[batch,1,2^i,2^i]`` in the generator's order (src/model.py:476-485)."""
    rs = np.random.RandomState(seed)
    maps = [_randn(rs, batch, 1, 4, 4)]
    for i in range(3, int(math.log2(size)) + 1):
        for _ in range(2):
            maps.append(_randn(rs, batch, 1, 2 ** i, 2 ** i))
    return maps


def make_pca_basis(seed: int = 2, dim: int = 512):
    """Orthonormal ``pc [dim,dim]`` (rows = components), decreasing ``sigma [dim,1]`` and a mean
    ``[dim,1]`` - a synthetic stand-in for ``GetPCA.perform_pca`` outputs (src/PCA.py:62-108)."""
    rs = np.random.RandomState(seed)
    q, _ = np.linalg.qr(rs.standard_normal((dim, dim)))
    pc = torch.from_numpy(q.T.astype(np.float32).copy())
    sigma = torch.from_numpy(np.linspace(1.0, 0.05, dim, dtype=np.float32)).reshape(-1, 1)
    mean = _randn(rs, dim, 1, scale=0.1)
    return pc, sigma, mean


def split_basis(pc, sigma, key_len: int = 64, shift: int = 448, fixed_sigma: float = 1.0):
    """``get_uv`` + ``get_alpha_bound`` (src/main.py:23-40) for batch_size 1."""
    dim = pc.shape[0]
    v_cap = pc[shift:shift + key_len].clone()
    u_cap = torch.cat([pc[:shift], pc[shift + key_len:dim]], 0)
    sigma_key = fixed_sigma * torch.ones(key_len, 1)
    sigma_main = torch.cat([sigma[:shift], sigma[shift + key_len:dim]], 0)
    return dict(v_cap=v_cap, u_cap=u_cap, sigma_key=sigma_key, sigma_main=sigma_main,
                max_alpha=3 * sigma_main, min_alpha=-3 * sigma_main)


def seeded(shape, seed, scale=1.0):
    return _randn(np.random.RandomState(seed), *shape, scale=scale)


# torchvision vgg16().features indices of the 13 convolutions by LPIPS slice (src/custom_lpips/pretrained_networks.py:108-117)
VGG_SLICES = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21), (24, 26, 28))
VGG_CHANNELS = (64, 128, 256, 512, 512)


def make_vgg_params(seed: int = 0, lin_weights: Dict[str, torch.Tensor] = None) -> Dict[str, torch.Tensor]:
    """Seeded random ``PNetLin`` state dict (src/custom_lpips/networks_basic.py:27-61): VGG16 conv weights in torchvision's
    initialisation scale (kaiming-normal, fan-out), small seeded biases so the bias path is exercised, and non-negative
    linear heads (the trained heads are non-negative) unless ``lin_weights`` supplies them.  The ImageNet weights cannot be
    downloaded in this environment: every LPIPS number in this repository is on these random weights."""
    rs = np.random.RandomState(seed)
    p: Dict[str, torch.Tensor] = {}
    cin = 3
    for si, convs in enumerate(VGG_SLICES):
        cout = VGG_CHANNELS[si]
        for idx in convs:
            std = (2.0 / (cout * 9)) ** 0.5
            p[f"net.slice{si + 1}.{idx}.weight"] = torch.from_numpy((rs.standard_normal((cout, cin, 3, 3)) * std).astype(np.float32))
            p[f"net.slice{si + 1}.{idx}.bias"] = torch.from_numpy((rs.standard_normal(cout) * 0.05).astype(np.float32))
            cin = cout
    for k in range(5):
        name = f"lin{k}.model.1.weight"
        if lin_weights is not None and name in lin_weights:
            p[name] = lin_weights[name].clone().float()
        else:
            p[name] = torch.from_numpy(np.abs(rs.standard_normal((1, VGG_CHANNELS[k], 1, 1))).astype(np.float32) * 0.1)
    return p
