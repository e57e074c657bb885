"""Runs the small-plane FIR kernel (blur, blur adjoint, down-2, up-2) and the flat bias-act kernel (forward, backward) once
each on [64, 512, 16, 16]-class tensors, for an ncu capture:
    ncu --set full --clock-control none --import-source on -k regex:"planes_kernel|bias_act_flat" -o out python tools/small_planes_profile.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import torch
from op import fused_leaky_relu, upfirdn2d
dev = "cuda"
k = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = (k[:, None] * k[None, :]) / 64
k4 = k2 * 4
n, c, h = 64, 512, 16
x = torch.randn(n, c, h, h, device=dev)
xo = torch.randn(n, c, h + 1, h + 1, device=dev)
xs = x[:, :, : h // 2, : h // 2].contiguous()
b = torch.randn(c, device=dev)
upfirdn2d(xo, k4, pad=(1, 1))
upfirdn2d(x, k4, pad=(2, 2))
upfirdn2d(x, k2, down=2, pad=(1, 1))
upfirdn2d(xs, k4, up=2, pad=(2, 1))
xg = x.clone().requires_grad_(True)
y = fused_leaky_relu(xg, b)
torch.autograd.grad(y, xg, torch.ones_like(y))
torch.cuda.synchronize()
