"""Diagnostic (GPU box): per-latent-slot gradient error of the fused synthesis vs the CPU oracle."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import fixtures as fx, oracle
from model import Generator
for size, B, seed in ((8, 1, 28), (16, 1, 36), (16, 5, 36), (32, 3, 31), (64, 2, 50)):
    params = fx.make_params(size, seed)
    g = Generator(size, 512, 8); g.load_state_dict(params, strict=False); g = g.eval().cuda()
    noise = fx.make_noise(size, seed + 1)
    lat = fx.seeded((B, oracle.n_latent(size), 512), seed + 2)
    lr = lat.clone().requires_grad_(True)
    ref = oracle.synthesis(params, lr, noise)
    ct = fx.seeded(tuple(ref.shape), seed + 3)
    (gref,) = torch.autograd.grad((ref * ct).sum(), lr)
    lg = lat.cuda().requires_grad_(True)
    img, _ = g([lg], input_is_latent=True, noise=[n.cuda() for n in noise])
    (gl,) = torch.autograd.grad((img * ct.cuda()).sum(), lg)
    gl = gl.cpu()
    print(f"size {size} B {B}: img err {float((img.detach().cpu()-ref.detach()).abs().max()):.2e}")
    for b in range(B):
        print("   b", b, " ".join(f"{float((gl[b,s]-gref[b,s]).norm()/gref[b,s].norm()):.1e}" for s in range(lat.shape[1])))
