"""Op-level microbenchmark (BASELINE.json configs[4]): op.upfirdn2d / op.fused_leaky_relu forward and backward on the
hot-path shapes, CUDA-event timed, reported as algorithmic GB/s against the measured HBM peak.
    python tools/op_microbench.py [--json out.json]"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
from op import fused_leaky_relu, upfirdn2d

try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0
dev = "cuda"


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()                      # evict L2 between iterations
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


rows = []
k = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = (k[:, None] * k[None, :]) / 64
for (n, c, h) in [(1, 32, 1024), (8, 32, 1024), (8, 64, 512), (16, 512, 64), (64, 512, 16)]:
    x = torch.randn(n, c, h, h, device=dev)
    b = torch.randn(c, device=dev)
    numel = x.numel()
    t = timeit(lambda: fused_leaky_relu(x, b))
    rows.append(("fused_leaky_relu fwd", (n, c, h, h), 8 * numel, t))
    xg = x.clone().requires_grad_(True)
    y = fused_leaky_relu(xg, None)
    gy = torch.randn_like(y)
    t = timeit(lambda: torch.autograd.grad(y, xg, gy, retain_graph=True))
    rows.append(("fused_leaky_relu bwd", (n, c, h, h), 12 * numel, t))
    del xg, y, gy
    # Blur after the up-conv: [N,C,2H+1,2H+1] -> [N,C,2H,2H]  (src/model.py:75-91)
    xo = torch.randn(n, c, h + 1, h + 1, device=dev)
    t = timeit(lambda: upfirdn2d(xo, k2 * 4, pad=(1, 1)))
    rows.append(("upfirdn2d blur pad(1,1)", (n, c, h + 1, h + 1), 4 * (xo.numel() + numel), t))
    t = timeit(lambda: upfirdn2d(x, k2 * 4, pad=(2, 2)))
    rows.append(("upfirdn2d blur-bwd pad(2,2)", (n, c, h, h), 4 * (xo.numel() + numel), t))
    del xo
    if h <= 512 or n == 1:
        t = timeit(lambda: upfirdn2d(x, k2, down=2, pad=(1, 1)))
        rows.append(("upfirdn2d down2 pad(1,1)", (n, c, h, h), 4 * (numel + numel // 4), t))
        xs = x[:, :, : h // 2, : h // 2].contiguous()
        t = timeit(lambda: upfirdn2d(xs, k2 * 4, up=2, pad=(2, 1)))
        rows.append(("upfirdn2d up2 pad(2,1)", (n, c, h // 2, h // 2), 4 * (numel + numel // 4), t))
        del xs
    del x
out = []
print(f"{'op':30s} {'shape':24s} {'ms':>8s} {'GB/s':>8s} {'of HBM':>7s}")
for name, shape, nbytes, t in rows:
    gbs = nbytes / t / 1e9
    print(f"{name:30s} {str(shape):24s} {t*1e3:8.3f} {gbs:8.0f} {gbs/PEAK:7.2f}")
    out.append({"op": name, "shape": list(shape), "ms": t * 1e3, "gbs": gbs, "frac_of_measured_hbm": gbs / PEAK})
if len(sys.argv) > 2 and sys.argv[1] == "--json":
    json.dump({"hbm_peak_gbs": PEAK, "rows": out}, open(sys.argv[2], "w"), indent=1)
