"""Diagnostic (GPU box): tensor-core (tf32) path against the fp32 CUDA-core path, layer by layer."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import fixtures as fx
from lfp_native import capi
from lfp_native.synthesis import SynthesisPlan

sizes = [int(a) for a in sys.argv[1:]] or [16, 32, 64]
for size in sizes:
    B, seed = 2, 70 + size
    params = fx.make_params(size, seed)
    plan = SynthesisPlan(size, device="cuda"); plan.load(params)
    noise = [n.cuda() for n in fx.make_noise(size, seed + 1)]
    lat = fx.seeded((B, plan.n_latent, 512), seed + 2).cuda()
    ct = fx.seeded((B, 3, size, size), seed + 3).cuda()
    out = {}
    for name, prec in (("fp32", capi.PREC_FP32), ("tf32", capi.PREC_TF32)):
        ws = plan.new_workspace(B)
        img = plan.forward(lat, noise, ws, prec)
        acts = plan.read_activations(B, ws)
        dl = plan.backward(ct, B, ws, prec)
        torch.cuda.synchronize()
        out[name] = (img, acts, dl)
    print(f"== size {size}: image max err {float((out['fp32'][0]-out['tf32'][0]).abs().max()):.3e} (scale {float(out['fp32'][0].abs().max()):.2f})")
    for i, (a, b) in enumerate(zip(out["fp32"][1], out["tf32"][1])):
        print(f"   conv {i} {tuple(a.shape)}: max err {float((a-b).abs().max()):.3e} scale {float(a.abs().max()):.2f} nan {int(torch.isnan(b).sum())}")
    g0, g1 = out["fp32"][2], out["tf32"][2]
    print("   dlatent rel err per slot:", " ".join(f"{float((g0[:,s]-g1[:,s]).norm()/g0[:,s].norm()):.1e}" for s in range(g0.shape[1])))
