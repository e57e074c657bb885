"""A/B of the stand-alone layer's layout / reduction kernels: run once as is and once with LFP_MC_SCALAR_LAYOUT=1 (the scalar
32 x 32 shared-memory transposes and the scalar dot kernel of the first version).  Prints fwd and fwd+bwd device times.
    python tools/modconv_layout_ab.py; LFP_MC_SCALAR_LAYOUT=1 python tools/modconv_layout_ab.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG):
    sys.path.insert(0, p)
import torch
from model import ModulatedConv2d
from lfp_native import capi
dev = "cuda"


def timeit(fn, iters=8, warm=2):
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_(); torch.cuda._sleep(2_000_000)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2]


tag = "scalar" if os.environ.get("LFP_MC_SCALAR_LAYOUT") else "vector"
for (cin, cout, res, up, B) in [(32, 32, 1024, False, 16), (64, 32, 1024, True, 16), (64, 64, 512, False, 16), (256, 256, 128, False, 16),
                                (512, 512, 16, False, 16), (512, 3, 64, False, 16)]:
    k = 1 if cout == 3 else 3
    m = ModulatedConv2d(cin, cout, k, 512, upsample=up, demodulate=cout != 3).to(dev)
    m.precision = capi.PREC_TF32
    hin = res // 2 if up else res
    x = torch.randn(B, cin, hin, hin, device=dev, requires_grad=True)
    st = torch.randn(B, 512, device=dev, requires_grad=True)
    gy = torch.randn(B, cout, res, res, device=dev)
    t_f = timeit(lambda: m(x.detach(), st.detach()))

    def fb():
        y = m(x, st)
        torch.autograd.grad(y, [x, st], gy)

    t_fb = timeit(fb)
    print(f"{tag} modconv {cin}->{cout}{' up' if up else ''} k{k} @{res} B={B}: fwd {t_f:.3f} ms  fwd+bwd {t_fb:.3f} ms", flush=True)
    del m, x, st, gy
