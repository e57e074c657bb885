import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")): sys.path.insert(0, p)
import torch, fixtures as fx
from lfp_native import capi
from lfp_native.synthesis import SynthesisPlan
from attribution import AttributionEngine
size=32
params = fx.make_params(size, 11); noise = fx.make_noise(size, 12); pc, s512, mean = fx.make_pca_basis(2)
plan = SynthesisPlan(size, device="cuda"); plan.load(params)
eng = AttributionEngine(plan, noise, pc, s512, mean, precision=capi.PREC_FP32)
sp = fx.split_basis(pc, s512, 64, 448, 1.0)
B=3
target = eng.render(fx.seeded((1, 512), 23).cuda()).clone()
a0 = (sp["sigma_main"].t() * fx.seeded((B, 448), 81)).contiguous()
for nsteps in (1,2,3):
    ref = eng.run(a0, target, steps=nsteps, native=False)
    for graph in (False, True):
        st = eng.init_state(a0)
        stp = eng.native_stepper(st, target, max_steps=16)
        stp.run(nsteps, graph=graph)
        torch.cuda.synchronize()
        print(nsteps, graph, {k: float((st[k]-ref[k]).abs().max()) for k in ("alpha","key","m_a","v_a","loss")})
