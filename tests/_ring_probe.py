"""Helper of test_ring_kernels_gpu.py: one synthesis forward + backward at 512 px (fp32 path) in this process, written to an
.npz.  The ring / non-ring kernel choice is read from the environment once per process (LFP_FIR_RING, LFP_ACTBWD_RING), so
the test runs this script twice."""
import os, sys
HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
for p in (ROOT, os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200"), HERE):
    sys.path.insert(0, p)
import numpy as np
import torch
import fixtures as fx
from lfp_native import capi
from lfp_native.synthesis import SynthesisPlan

size, seed, B = 512, 11, 2
params = fx.make_params(size, seed, 2)
plan = SynthesisPlan(size, channel_multiplier=2, device="cuda")
plan.load(params)
lat = fx.seeded((B, plan.n_latent, 512), seed + 1)
noise = []
for i in range(plan.num_noise):
    res = 4 if i == 0 else 8 << ((i - 1) // 2)
    noise.append(fx.seeded((1, 1, res, res), seed + 10 + i))
ws = plan.new_workspace(B)
img = plan.forward(lat.cuda(), [n.cuda() for n in noise], ws, capi.PREC_FP32)
ct = fx.seeded(tuple(img.shape), seed + 3)
d_lat = plan.backward(ct.cuda(), B, ws, capi.PREC_FP32)
np.savez(sys.argv[1], img=img.cpu().numpy(), d_lat=d_lat.cpu().numpy())
