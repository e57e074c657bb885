"""Fingerprint embedding on top of the native generator - the ``GetGen`` surface of
src/generator.py (generate_with_alpha :69-107, get_new_latent :148-161, generate_image :170-174)
without the argparse global: settings are constructor arguments.

    wx = w0 + sd * V^T diag(sigma) k,   w0 = U^T alpha + mu,   V = pc[shift:shift+key_len]

PCA (src/PCA.py:62-108) is one-off set-up outside the hot path: the mapping network of the 10 000 samples and the fp64
mean / covariance run on the native kernels (lfp_mapping_forward, lfp_pca_covariance); only the 512 x 512 symmetric
eigen-decomposition is a library call (torch.linalg.eigh).
"""
from __future__ import annotations

import math
from typing import List, Optional, Sequence

import numpy as np
import torch

from model import Generator


def get_noise(img_size: int, device) -> List[torch.Tensor]:
    """Fixed noise maps with the reference's RNG protocol (src/utils.py:128-138): the 4x4 map from
    ``default_rng(2002)``, the rest from the global numpy RNG (seeded 2022 by Generator.__init__,
    src/model.py:404)."""
    rng = np.random.default_rng(seed=2002)
    maps = [torch.tensor(rng.standard_normal((1, 1, 4, 4)), dtype=torch.float32, device=device)]
    for i in range(3, int(math.log2(img_size)) + 1):
        for _ in range(2):
            maps.append(torch.tensor(np.random.standard_normal((1, 1, 2 ** i, 2 ** i)), dtype=torch.float32,
                                     device=device))
    return maps


def perform_pca(g_ema: Generator, n_samples: int = 10000, seed: Optional[int] = None):
    """``GetPCA.perform_pca`` (src/PCA.py:62-108): style vectors of ``n_samples`` random z, principal
    axes sorted by decreasing variance.  Returns ``(pc [512,512], sigma_512 [512,1], mean [512,1])``.
    Component signs are arbitrary (as with sklearn); pass ``pc`` explicitly where parity matters."""
    dev = g_ema.input.input.device
    gen = torch.Generator(device=dev)
    if seed is not None:
        gen.manual_seed(seed)
    from lfp_native.mapping import MappingPlan, covariance
    with torch.no_grad():
        z = torch.randn(n_samples, g_ema.style_dim, device=dev, generator=gen)
        n_mlp = len(g_ema.style) - 1
        plan = MappingPlan(g_ema.style_dim, n_mlp, g_ema.style[1].lr_mul, device=dev)
        plan.load({k: v for k, v in g_ema.state_dict().items() if k.startswith("style.")})
        w = plan.forward(z)                       # src/PCA.py:69 g_ema.style(noise_sample)
        mean, cov = covariance(w)                 # what sklearn's PCA().fit diagonalises (src/PCA.py:72-73)
        evals, evecs = torch.linalg.eigh(cov)
        order = torch.argsort(evals, descending=True)
        pc = evecs[:, order].t().float().contiguous()
        sigma = evals[order].clamp_min(0).sqrt().float().reshape(-1, 1)
    return pc, sigma, mean.float().reshape(-1, 1)


class GetGen:
    def __init__(self, img_size: int = 256, key_len: int = 64, shift: int = 448, sigma: float = 1.0, sd: int = 1,
                 ckpt: Optional[str] = None, device="cuda:0", pca=None, seed: Optional[int] = None,
                 batch_size: int = 1, augmentation: str = "None", noise_sigma: float = 0.1, blur_sigma: float = 0.5,
                 jpeg_quality: float = 50):
        self.device = torch.device(device)
        self.img_size, self.key_len, self.batch_size, self.sd_moved = img_size, key_len, batch_size, sd
        self.style_space_dim, self.mapping_network_layer = 512, 8
        self.num_main_pc = self.style_space_dim - key_len
        g = Generator(img_size, self.style_space_dim, self.mapping_network_layer)
        if ckpt is not None:
            g.load_state_dict(torch.load(ckpt, map_location="cpu")["g_ema"], strict=False)  # src/generator.py:50
        self.g_ema = g.eval().to(self.device)
        for p in self.g_ema.parameters():
            p.requires_grad_(False)
        self.pc, self.sigma_512, self.latent_mean = pca if pca is not None else perform_pca(self.g_ema, seed=seed)
        self.pc, self.sigma_512 = self.pc.to(self.device), self.sigma_512.to(self.device)
        self.latent_mean = self.latent_mean.to(self.device)
        # get_uv (src/main.py:30-40)
        self.v_cap = self.pc[shift:shift + key_len].contiguous()
        self.u_cap = torch.cat([self.pc[:shift], self.pc[shift + key_len:]], 0).contiguous()
        self.sigma_64 = sigma * torch.ones(key_len, 1, device=self.device)
        self.sigma_448 = torch.cat([self.sigma_512[:shift], self.sigma_512[shift + key_len:]], 0)
        self.key = None
        # robustness attack applied once per target image (src/params.py:27-32 flag names and defaults)
        self.augmentation_name = augmentation
        self._attack = None
        if augmentation != "None":
            import attacks
            self._attack = attacks.attack_initializer(augmentation, noise_sigma, blur_sigma, jpeg_quality).to(self.device)

    def get_new_latent(self, v, s, k, w0):
        """``w0 + sd * (V^T diag(s)) k`` (src/generator.py:148-161)."""
        vs = v.t() @ torch.diag(s.reshape(-1))
        return w0 + self.sd_moved * (vs @ k.reshape(-1, 1))

    def generate_image(self, style_vector, noise):
        """src/generator.py:170-174."""
        img, _ = self.g_ema([style_vector.reshape(1, -1)], noise=noise, input_is_latent=True)
        return img

    def generate_with_alpha(self, alpha, u_cap_t, sigma_64, v_cap, noise, key=None):
        """src/generator.py:69-107; ``alpha`` [n_main, B].  Returns (imgs, w0 [B,512], wx [B,512], key)."""
        if key is None:
            key = torch.randint(2, (self.key_len, alpha.shape[1]), device=self.device)
        self.key = key
        w0 = (u_cap_t @ alpha + self.latent_mean).t()
        wx = w0 + self.sd_moved * ((sigma_64 * key).t() @ v_cap)
        with torch.no_grad():
            imgs, _ = self.g_ema([wx], noise=noise, input_is_latent=True)
        return imgs.detach(), w0.detach(), wx.detach(), key

    def augmentation(self, target_img):
        """Image augmentation, default None (src/generator.py:163-168): the test-time attack of ``attacks.py``."""
        if self._attack is None:
            return target_img
        with torch.no_grad():
            return self._attack(target_img)
