"""Context for the FIR rooflines: what plain copies with the same footprints reach on this GPU (CUDA events, L2 flushed).
contiguous clone, 2-D cropped copy [.,1025,1025] -> [.,1024,1024] (the blur's footprint), and its reverse (padding)."""
import torch, json, os
dev = "cuda"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6544.0
n, c, h = 8, 32, 1024
x = torch.randn(n, c, h, h, device=dev)
xo = torch.randn(n, c, h + 1, h + 1, device=dev)
y = torch.empty_like(x)
yo = torch.empty_like(xo)
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
def t(fn, nbytes, name):
    for _ in range(3): fn()
    ts = []
    for _ in range(15):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1))
    ts.sort(); ms = ts[len(ts) // 2]
    print(f"{name:34s} {ms*1e3:8.1f} us {nbytes/ms/1e6:7.0f} GB/s {nbytes/ms/1e6/PEAK:5.2f}")
N = x.numel()
t(lambda: y.copy_(x), 8 * N, "contiguous copy 1024^2")
t(lambda: yo.copy_(xo), 8 * xo.numel(), "contiguous copy 1025^2")
t(lambda: y.copy_(xo[:, :, :h, :h]), 8 * N, "crop copy 1025^2 -> 1024^2")
t(lambda: yo[:, :, :h, :h].copy_(x), 8 * N, "pad copy 1024^2 -> 1025^2")
t(lambda: torch.add(x, 1.0, out=y), 8 * N, "add scalar 1024^2")
t(lambda: torch.nn.functional.avg_pool2d(xo, 2, stride=1), 4 * (xo.numel() + N), "avg_pool2d 2x2 s1 1025->1024")
