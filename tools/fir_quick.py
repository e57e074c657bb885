"""Times the blur / blur-bwd / down2 / up2 configurations of op.upfirdn2d at (8,32,1024) (CUDA events, L2 flushed)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import torch
from op import upfirdn2d
dev = "cuda"
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6544.0
k = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = (k[:, None] * k[None, :]) / 64
n, c, h = 8, 32, 1024
x = torch.randn(n, c, h, h, device=dev)
xo = torch.randn(n, c, h + 1, h + 1, device=dev)
xs = x[:, :, : h // 2, : h // 2].contiguous()
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, nbytes, name):
    for _ in range(3): fn()
    ts = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print(f"{name:10s} {ms*1e3:8.1f} us {nbytes/ms/1e6:7.0f} GB/s {nbytes/ms/1e6/PEAK:5.2f}")
N = x.numel()
t(lambda: upfirdn2d(xo, k2 * 4, pad=(1, 1)), 4 * (xo.numel() + N), "blur")
t(lambda: upfirdn2d(x, k2 * 4, pad=(2, 2)), 4 * (xo.numel() + N), "blur_bwd")
t(lambda: upfirdn2d(x, k2, down=2, pad=(1, 1)), 5 * N, "down2")
t(lambda: upfirdn2d(xs, k2 * 4, up=2, pad=(2, 1)), 5 * N, "up2")
