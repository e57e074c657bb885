"""Robustness attacks applied ONCE to a target image before attribution (src/main.py:124 -> GetGen.augmentation,
src/generator.py:163-168 -> attack_initializer, src/attack_methods/attack_initializer.py:12-35), test-time settings
(``is_train=False``): Gaussian blur 25x25, additive Gaussian noise, JPEG (the reference's DiffJPEG with hard rounding) and
their combination.  Outside the step loop; images are ``[B, 3, H, W]`` CUDA tensors in [-1, 1].

* ``GaussianBlur``: torchvision's ``T.GaussianBlur((25, 25), sigma)`` (src/attack_methods/Gaussian_blur.py:11-33): reflect-pad by
  12, then the normalised separable Gaussian as one FIR pass through the native ``op.upfirdn2d``.
* ``GaussianNoise``: ``image + N(0, std)`` clamped to [-1, 1] (Gaussian_noise.py:10-47).
* ``Jpeg``: image -> [0, 255] YCbCr -> 4:2:0 chroma -> 8x8 DCT -> quantise (tables scaled by the quality factor, hard
  rounding at test time) -> de-quantise -> IDCT -> chroma up-sampling -> RGB -> clamp (Jpeg_compression.py:5-20,
  DiffJPEG_master/DiffJPEG.py, modules/compression.py, modules/decompression.py, utils.py).
* ``Combination``: blur, noise, JPEG in that order, each always applied at test time (Combination.py:4-29).
"""
from __future__ import annotations

import math

import numpy as np
import torch
import torch.nn.functional as F
from torch import nn

from op import upfirdn2d

# standard JPEG luminance / chrominance quantisation tables as the reference holds them (transposed), DiffJPEG_master/utils.py:7-21
_Y_TABLE = np.array([[16, 11, 10, 16, 24, 40, 51, 61], [12, 12, 14, 19, 26, 58, 60, 55], [14, 13, 16, 24, 40, 57, 69, 56],
                     [14, 17, 22, 29, 51, 87, 80, 62], [18, 22, 37, 56, 68, 109, 103, 77], [24, 35, 55, 64, 81, 104, 113, 92],
                     [49, 64, 78, 87, 103, 121, 120, 101], [72, 92, 95, 98, 112, 100, 103, 99]], dtype=np.float32).T
_C_TABLE = np.full((8, 8), 99, dtype=np.float32)
_C_TABLE[:4, :4] = np.array([[17, 18, 24, 47], [18, 21, 26, 66], [24, 26, 56, 99], [47, 66, 99, 99]], dtype=np.float32).T


def gaussian_kernel1d(size: int, sigma: float) -> torch.Tensor:
    """torchvision ``_get_gaussian_kernel1d``: pdf on linspace(-(size-1)/2, (size-1)/2), normalised."""
    lim = (size - 1) * 0.5
    x = torch.linspace(-lim, lim, steps=size)
    pdf = torch.exp(-0.5 * (x / sigma).pow(2))
    return pdf / pdf.sum()


class GaussianBlur(nn.Module):
    def __init__(self, sigma: float, filter_size: int = 25):
        super().__init__()
        self.sigma, self.filter_size = float(sigma), int(filter_size)
        k1 = gaussian_kernel1d(self.filter_size, self.sigma)
        self.register_buffer("kernel", torch.outer(k1, k1))

    def forward(self, image):
        p = self.filter_size // 2
        x = F.pad(image, [p, p, p, p], mode="reflect")
        return upfirdn2d(x, self.kernel.to(image.device, image.dtype), pad=(0, 0))   # symmetric taps: flip is a no-op


class GaussianNoise(nn.Module):
    def __init__(self, std: float, generator: torch.Generator = None):
        super().__init__()
        self.std, self.generator = float(std), generator

    def forward(self, image):
        noise = torch.empty_like(image).normal_(0.0, self.std, generator=self.generator)
        return torch.clamp(image + noise, -1, 1).float()


def quality_to_factor(quality: float) -> float:
    """DiffJPEG_master/utils.py:36-48."""
    q = 5000.0 / quality if quality < 50 else 200.0 - quality * 2
    return q / 100.0


class Jpeg(nn.Module):
    def __init__(self, quality: float = 50, differentiable: bool = False):
        super().__init__()
        self.factor = quality_to_factor(quality)
        self.differentiable = differentiable
        u = np.arange(8)
        # dct[x, u] = cos((2x + 1) u pi / 16); the 2-D transforms are separable products of it
        basis = np.cos((2 * u[:, None] + 1) * u[None, :] * np.pi / 16).astype(np.float32)
        alpha = np.array([1.0 / np.sqrt(2)] + [1.0] * 7, dtype=np.float32)
        self.register_buffer("basis", torch.from_numpy(basis))
        self.register_buffer("alpha2", torch.from_numpy(np.outer(alpha, alpha)))
        self.register_buffer("y_table", torch.from_numpy(_Y_TABLE))
        self.register_buffer("c_table", torch.from_numpy(_C_TABLE))
        self.register_buffer("to_ycc", torch.tensor([[0.299, 0.587, 0.114], [-0.168736, -0.331264, 0.5],
                                                     [0.5, -0.418688, -0.081312]], dtype=torch.float32))
        self.register_buffer("to_rgb", torch.tensor([[1.0, 0.0, 1.402], [1, -0.344136, -0.714136], [1, 1.772, 0]],
                                                    dtype=torch.float32))

    def _round(self, x):
        if self.differentiable:      # DiffJPEG_master/utils.py:24-33
            r = torch.round(x)
            return r + (x - r) ** 3
        return torch.round(x)

    def _code(self, plane, table):
        """[B, H, W] -> 8x8 blocks -> DCT -> quantise -> de-quantise -> IDCT -> [B, H, W]."""
        B, H, W = plane.shape
        blk = plane.reshape(B, H // 8, 8, W // 8, 8).permute(0, 1, 3, 2, 4) - 128.0               # [B, h8, w8, x, y]
        coef = torch.einsum("bhwxy,xu,yv->bhwuv", blk, self.basis, self.basis) * (self.alpha2 * 0.25)
        q = self._round(coef / (table * self.factor)) * (table * self.factor)
        rec = 0.25 * torch.einsum("bhwuv,xu,yv->bhwxy", q * self.alpha2, self.basis, self.basis) + 128.0
        return rec.permute(0, 1, 3, 2, 4).reshape(B, H, W)

    def forward(self, image):
        if self.basis.device != image.device:     # follow the image like the reference's `.to(opt.device)` (attack_initializer.py:22)
            self.to(image.device)
        x = (image + 1.0) / 2.0 * 255.0                                                           # Jpeg_compression.py:16 + compress(image * 255)
        ycc = torch.einsum("bchw,kc->bkhw", x, self.to_ycc)
        y, cb, cr = ycc[:, 0], ycc[:, 1] + 128.0, ycc[:, 2] + 128.0
        cb = F.avg_pool2d(cb[:, None], 2, 2)[:, 0]
        cr = F.avg_pool2d(cr[:, None], 2, 2)[:, 0]
        y, cb, cr = self._code(y, self.y_table), self._code(cb, self.c_table), self._code(cr, self.c_table)
        cb = cb.repeat_interleave(2, 1).repeat_interleave(2, 2) - 128.0
        cr = cr.repeat_interleave(2, 1).repeat_interleave(2, 2) - 128.0
        rgb = torch.einsum("bkhw,ck->bchw", torch.stack([y, cb, cr], 1), self.to_rgb)
        rgb = torch.clamp(rgb, 0, 255) / 255.0
        return rgb * 2.0 - 1.0


class Combination(nn.Module):
    def __init__(self, attacks):
        super().__init__()
        self.attacks = nn.ModuleList(attacks)

    def forward(self, image):
        for a in self.attacks:
            image = a(image)
        return image


def attack_initializer(attack_method: str, noise_sigma: float = 0.1, blur_sigma: float = 0.5, jpeg_quality: float = 50,
                       generator: torch.Generator = None) -> nn.Module:
    """Test-time attack by the reference's flag names (--augmentation {Noise, Blur, Jpeg, Combination}, src/params.py:27-32)."""
    if attack_method == "Noise":
        return GaussianNoise(noise_sigma, generator)
    if attack_method == "Blur":
        return GaussianBlur(blur_sigma)
    if attack_method == "Jpeg":
        return Jpeg(jpeg_quality)
    if attack_method == "Combination":
        return Combination([GaussianBlur(blur_sigma), GaussianNoise(noise_sigma, generator), Jpeg(jpeg_quality)])
    raise ValueError("Not available Attacks")
