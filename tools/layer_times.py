"""Per-launch CUDA-event timing of one synthesis forward+backward (GPU box): python tools/layer_times.py [size] [B] [fp32|tf32]"""
import os, sys, ctypes as C
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import torch
import fixtures as fx
from lfp_native import capi
from lfp_native.synthesis import SynthesisPlan

size = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
B = int(sys.argv[2]) if len(sys.argv) > 2 else 20
prec = capi.PREC_TF32 if (len(sys.argv) > 3 and sys.argv[3] == "tf32") else capi.PREC_FP32
params = fx.make_params(size, 1346)
plan = SynthesisPlan(size, device="cuda"); plan.load(params)
noise = [n.cuda() for n in fx.make_noise(size, 2002)]
lat = fx.seeded((B, plan.n_latent, 512), 3).cuda()
ct = fx.seeded((B, 3, size, size), 4).cuda()
ws = plan.new_workspace(B)
L = capi.lib()
ITERS = int(os.environ.get('LT_ITERS', '3'))
for it in range(ITERS):
    if it == ITERS - 1:
        capi.check(L.lfp_synth_profile_begin(plan._h, 0b11111))
    img = plan.forward(lat, noise, ws, prec)
    dl = plan.backward(ct, B, ws, prec)
torch.cuda.synchronize()
n = len(capi.KINDS)
ms, cnt, fl, by = (C.c_double * n)(), (C.c_int64 * n)(), (C.c_double * n)(), (C.c_double * n)()
capi.check(L.lfp_synth_profile_end(plan._h, ms, cnt, fl, by))
N = 4096
kinds, lms, lfl, lby = (C.c_int * N)(), (C.c_float * N)(), (C.c_double * N)(), (C.c_double * N)()
k = L.lfp_synth_profile_launches(plan._h, N, kinds, lms, lfl, lby)
tot = sum(lms[i] for i in range(k))
print(f"size {size} B {B} {'tf32' if prec else 'fp32'}: {k} profiled launches, {tot:.3f} ms (sum of kernels)")
for i in range(k):
    t = lms[i]
    print(f"{i:3d} {capi.KINDS[kinds[i]]:10s} {t*1e3:9.1f} us  {lfl[i]/1e9:8.2f} GF {lfl[i]/t/1e9 if t>0 else 0:8.1f} TF/s  {lby[i]/1e6:9.1f} MB {lby[i]/t/1e6 if t>0 else 0:8.1f} GB/s")
for j in range(n):
    print(f"{capi.KINDS[j]:10s} {ms[j]:8.3f} ms {cnt[j]:4d} launches  {fl[j]/max(ms[j],1e-9)/1e9:8.1f} TF/s {by[j]/max(ms[j],1e-9)/1e6:8.1f} GB/s")
