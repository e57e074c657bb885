"""Device time of op.upfirdn2d on many small planes (the shapes the small-plane kernel of csrc/upfirdn2d.cu serves).
    python tools/fir_small_planes.py            # LFP_FIR_NO_PLANES=1: the ring / streaming kernels; LFP_SP_STAGE=n: staged floats per CTA"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import torch
from op import upfirdn2d
dev = "cuda"
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timeit(fn, iters=20, warm=3):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_(); torch.cuda._sleep(1_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


k = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = (k[:, None] * k[None, :]) / 64
k4 = k2 * 4
tag = f"stage={os.environ.get('LFP_SP_STAGE', 'default')} planes={'off' if os.environ.get('LFP_FIR_NO_PLANES') else 'on'}"
for (n, c, h) in [(256, 512, 8), (64, 512, 16), (64, 512, 32), (16, 512, 32)]:
    x = torch.randn(n, c, h, h, device=dev)
    xo = torch.randn(n, c, h + 1, h + 1, device=dev)
    xs = x[:, :, : h // 2, : h // 2].contiguous()
    rows = [("blur", lambda: upfirdn2d(xo, k4, pad=(1, 1)), 4 * (xo.numel() + x.numel())),
            ("blur-bwd", lambda: upfirdn2d(x, k4, pad=(2, 2)), 4 * (xo.numel() + x.numel())),
            ("down2", lambda: upfirdn2d(x, k2, down=2, pad=(1, 1)), 5 * x.numel()),
            ("up2", lambda: upfirdn2d(xs, k4, up=2, pad=(2, 1)), 5 * x.numel())]
    out = []
    for name, fn, nbytes in rows:
        t = timeit(fn)
        out.append(f"{name} {t*1e6:6.1f} us {nbytes/t/1e9/PEAK:4.2f}")
    print(f"{tag} ({n},{c},{h},{h}): " + " | ".join(out), flush=True)
