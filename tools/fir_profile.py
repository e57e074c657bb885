"""Runs the four hot op.upfirdn2d configurations once each (for an ncu capture of the public op's kernels)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import torch
from op import upfirdn2d
dev = "cuda"
k = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = (k[:, None] * k[None, :]) / 64
n, c, h = 8, 32, 1024
x = torch.randn(n, c, h, h, device=dev)
xo = torch.randn(n, c, h + 1, h + 1, device=dev)
xs = x[:, :, : h // 2, : h // 2].contiguous()
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    upfirdn2d(xo, k2 * 4, pad=(1, 1))
    upfirdn2d(x, k2 * 4, pad=(2, 2))
    upfirdn2d(x, k2, down=2, pad=(1, 1))
    upfirdn2d(xs, k2 * 4, up=2, pad=(2, 1))
torch.cuda.synchronize()
