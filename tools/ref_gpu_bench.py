#!/usr/bin/env python
"""The reference's OWN CUDA path on this B200: JIT-built upfirdn2d / fused_bias_act extensions
+ cuDNN grouped convs, batch 1, driven through the unmodified ``main.optimization`` (src/main.py:45-89) of the copy
under baseline/_ref/src, MSE loss (LPIPS weights are not available offline).

    python tools/ref_gpu_bench.py [--size 1024] [--steps 30] > profiles/r02_reference_gpu.json

Reports trajectory-steps/s with ``torch.backends.cudnn.allow_tf32`` True (the reference's de-facto default) and False,
plus the synthesis forward and forward+backward-to-latent times.  Not part of bench.py's reference arm (that arm is
the CPU path by the task's definition); this is the honest GPU comparison point for DESIGN.md.
"""
import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--steps", type=int, default=30)
    a = ap.parse_args()
    import torch
    import fixtures as fx
    from oracle import reference_harness as rh
    params = fx.make_params(a.size, seed=1346)
    noise = fx.make_noise(a.size, seed=2002)
    pc, sigma, mean = fx.make_pca_basis(2)
    t0 = time.time()
    loop = rh.ReferenceLoop(a.size, params, noise, pc, sigma, mean, 64, 448, 1.0, device="cuda:0", real_ops=True)
    setup_s = time.time() - t0
    lhs = rh.lhs_sample(1, loop.n_main, 300)
    out = {"gpu": torch.cuda.get_device_name(0), "size": a.size, "batch": 1, "loss": "mse", "setup_seconds": setup_s,
           "torch": torch.__version__, "cudnn": torch.backends.cudnn.version()}
    g = loop.gen.g_ema
    for tf32 in (True, False):
        torch.backends.cudnn.allow_tf32 = tf32
        loop.run(lhs, 5)                       # warm-up (cuDNN autotune off by default; first-call plans)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        loop.run(lhs, a.steps)
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / a.steps
        # synthesis only
        w = fx.seeded((1, 512), 9).cuda().requires_grad_(True)

        def fwd():
            return g([w], input_is_latent=True, noise=loop.noise)[0]

        def fwdbwd():
            img = fwd()
            torch.autograd.grad(img.square().mean(), w)

        res = {}
        for name, fn in (("synthesis_fwd_ms", lambda: fwd().detach()), ("synthesis_fwd_bwd_ms", fwdbwd)):
            for _ in range(3):
                fn()
            torch.cuda.synchronize()
            e0.record()
            for _ in range(10):
                fn()
            e1.record()
            torch.cuda.synchronize()
            res[name] = e0.elapsed_time(e1) / 10
        out["allow_tf32=%s" % tf32] = dict(res, optimization_ms_per_step=ms, trajectory_steps_per_s=1e3 / ms, steps=a.steps)
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
