"""Op-level microbenchmark (BASELINE.json configs[4]): op.upfirdn2d / op.fused_leaky_relu forward and backward on the
hot-path shapes, reported as algorithmic GB/s against the measured HBM peak, and model.ModulatedConv2d (native stand-alone
layer, lfp_modconv_*) forward and forward+backward over 4-1024 px and batch 1-256, reported as TFLOP/s (2 Cin Cout 9 Hout^2
plain, 2 Cin Cout 9 Hin^2 up; backward = data gradient, the same again) and as a fraction of max(tensor, HBM) bound.
CUDA-event timed with the host running ahead of the device (see timeit), L2 flushed between iterations.
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
    """Median device time of fn().  The GPU is kept busy (L2 flush + a ~1 ms spin kernel) while the host enqueues fn(), so
    the two events bracket the kernels of fn() only; without the run-ahead every row under ~60 us measured the Python /
    ctypes enqueue time instead of the kernel (round-2 tables before this change)."""
    torch.cuda.synchronize()
    for _ in range(warm):
        fn()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(iters):
        flush.zero_()                      # evict L2 between iterations
        torch.cuda._sleep(2_000_000)       # ~1 ms of device spin: the host runs ahead of the GPU
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ts.sort()
    return ts[len(ts) // 2] * 1e-3


def host_us(fn, iters=20):
    """Host time of one call (enqueue only, no synchronisation inside the loop)."""
    import time
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(iters):
        fn()
    t1 = time.perf_counter()
    torch.cuda.synchronize()
    return (t1 - t0) / iters * 1e6


rows = []
k = torch.tensor([1., 3., 3., 1.], device=dev)
k2 = (k[:, None] * k[None, :]) / 64
k2x4 = k2 * 4          # the gain-4 kernel of Blur / Upsample, built once (not a launch inside the timed call)
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
    t = timeit(lambda: upfirdn2d(xo, k2x4, pad=(1, 1)))
    rows.append(("upfirdn2d blur pad(1,1)", (n, c, h + 1, h + 1), 4 * (xo.numel() + numel), t))
    t = timeit(lambda: upfirdn2d(x, k2x4, pad=(2, 2)))
    rows.append(("upfirdn2d blur-bwd pad(2,2)", (n, c, h, h), 4 * (xo.numel() + numel), t))
    del xo
    if h <= 512 or n == 1:
        t = timeit(lambda: upfirdn2d(x, k2, down=2, pad=(1, 1)))
        rows.append(("upfirdn2d down2 pad(1,1)", (n, c, h, h), 4 * (numel + numel // 4), t))
        xs = x[:, :, : h // 2, : h // 2].contiguous()
        t = timeit(lambda: upfirdn2d(xs, k2x4, up=2, pad=(2, 1)))
        rows.append(("upfirdn2d up2 pad(2,1)", (n, c, h // 2, h // 2), 4 * (numel + numel // 4), t))
        del xs
    del x
# ---- ModulatedConv2d sweep: every generator layer shape, batch 1 .. 256 capped by the activation footprint ----
from model import ModulatedConv2d
from lfp_native import capi
try:
    TC_PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["bf16_tflops_sustained"] / 2.0   # kind::tf32 = half the bf16 rate
except Exception:
    TC_PEAK = 700.0
conv_rows = []
LAYERS = [(512, 512, 4, False), (512, 512, 8, True), (512, 512, 16, False), (512, 512, 32, True), (512, 512, 64, False),
          (512, 256, 128, True), (256, 256, 128, False), (256, 128, 256, True), (128, 128, 256, False), (128, 64, 512, True),
          (64, 64, 512, False), (64, 32, 1024, True), (32, 32, 1024, False)]
for (cin, cout, res, up) in ([] if os.environ.get("OPMB_SKIP_MODCONV") else LAYERS):   # OPMB_SKIP_MODCONV=1: FIR / bias-act rows only
    m = ModulatedConv2d(cin, cout, 3, 512, upsample=up).to(dev)
    m.precision = capi.PREC_TF32
    hin = res // 2 if up else res
    for B in (1, 4, 16, 64, 256):
        if B * max(cin * hin * hin, cout * res * res) * 4 > (3 << 30):    # keep one activation under 3 GB
            continue
        x = torch.randn(B, cin, hin, hin, device=dev, requires_grad=True)
        st = torch.randn(B, 512, device=dev, requires_grad=True)
        gy = torch.randn(B, cout, res, res, device=dev)
        flops = 2.0 * B * cin * cout * 9 * (hin * hin if up else res * res)
        nbytes = 4.0 * (B * cin * hin * hin + B * cout * res * res + 9 * cin * cout)
        t_f = timeit(lambda: m(x.detach(), st.detach()), iters=8)

        def fb():
            y = m(x, st)
            torch.autograd.grad(y, [x, st], gy)

        t_fb = timeit(fb, iters=8)
        conv_rows.append((f"modconv {cin}->{cout}{' up' if up else ''} @{res}", B, flops, nbytes, t_f, t_fb))
        del x, st, gy
    del m
print(f"{'ModulatedConv2d (tf32 path)':34s} {'B':>4s} {'fwd ms':>8s} {'TF/s':>7s} {'of bound':>8s} {'fwd+bwd ms':>10s} {'TF/s':>7s}")
conv_out = []
for name, B, flops, nbytes, t_f, t_fb in conv_rows:
    bound = max(flops / (TC_PEAK * 1e12), nbytes / (PEAK * 1e9))
    print(f"{name:34s} {B:4d} {t_f*1e3:8.3f} {flops/t_f/1e12:7.1f} {bound/t_f:8.2f} {t_fb*1e3:10.3f} {2*flops/t_fb/1e12:7.1f}")
    conv_out.append({"op": name, "batch": B, "fwd_ms": t_f * 1e3, "fwd_tflops": flops / t_f / 1e12, "fwd_frac_of_bound": bound / t_f,
                     "fwd_bwd_ms": t_fb * 1e3, "fwd_bwd_tflops": 2 * flops / t_fb / 1e12})
out = []
print("# FIR / bias-act rows: algorithmic bytes / device time against the measured HBM copy rate; tensors under ~60 MB (the 16 px rows)\n"
      "# leave their output in the 126 MB L2, so their fraction measures launch ramp and on-chip work, not DRAM")
print(f"{'op':30s} {'shape':24s} {'ms':>8s} {'GB/s':>8s} {'of HBM':>7s}")
for name, shape, nbytes, t in rows:
    gbs = nbytes / t / 1e9
    print(f"{name:30s} {str(shape):24s} {t*1e3:8.3f} {gbs:8.0f} {gbs/PEAK:7.2f}")
    out.append({"op": name, "shape": list(shape), "ms": t * 1e3, "gbs": gbs, "frac_of_measured_hbm": gbs / PEAK})
# host cost of one call of the public ops (what bounds a Python loop over small tensors)
xs_ = torch.randn(1, 512, 16, 16, device=dev)
bs_ = torch.randn(512, device=dev)
m_ = ModulatedConv2d(512, 512, 3, 512).to(dev)
m_.precision = capi.PREC_TF32
st_ = torch.randn(1, 512, device=dev)
host = {"upfirdn2d": host_us(lambda: upfirdn2d(xs_, k2x4, pad=(2, 2))), "fused_leaky_relu": host_us(lambda: fused_leaky_relu(xs_, bs_)),
        "ModulatedConv2d.forward": host_us(lambda: m_(xs_, st_))}
print("host enqueue time per call (us): " + ", ".join(f"{k} {v:.1f}" for k, v in host.items()))
if len(sys.argv) > 2 and sys.argv[1] == "--json":
    json.dump({"hbm_peak_gbs": PEAK, "tf32_peak_tflops": TC_PEAK, "rows": out, "modconv_rows": conv_out, "host_enqueue_us": host}, open(sys.argv[2], "w"), indent=1)
