"""Diagnostic (GPU box): per-step drift of the batched engine against the CPU oracle loop at 32 px,
free-running and teacher-forced, to separate chaotic divergence from arithmetic mismatch."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "attributing-image-generative-models-using-latent-fingerprints-sg2_b200")
for p in (ROOT, PKG, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np, torch
import fixtures as fx, oracle
from lfp_native.synthesis import SynthesisPlan
from attribution import AttributionEngine, get_lr

size, seed = 32, 11
params = fx.make_params(size, seed); noise = fx.make_noise(size, seed + 1)
pc, sigma, mean = fx.make_pca_basis(2); sp = fx.split_basis(pc, sigma, 64, 448, 1.0)
plan = SynthesisPlan(size, device="cuda"); plan.load(params)
eng = AttributionEngine(plan, noise, pc, sigma, mean)
g = np.load(os.path.join(ROOT, "tests/golden/attribution.npz"))
target = torch.from_numpy(g["embed/gwa_img"]); lhs = torch.from_numpy(g["loop/lhs"])
a0 = eng.alpha0_from_lhs(lhs[:1])
st = eng.init_state(a0)
# oracle free-running, recording per-step state
a = a0.cpu().t().clone().requires_grad_(True); k = torch.zeros(64, 1, requires_grad=True)
opt = torch.optim.Adam([a, k], lr=0.2)
for i in range(12):
    # teacher-forced engine step from the oracle's current state
    stf = eng.init_state(a.detach().t().contiguous())
    stf["key"].copy_(k.detach().t()); 
    w0, wx = eng.embed(stf["alpha"], stf["key"])
    mse_e, dwx_e, _ = eng.loss_and_grad(wx, target.cuda())
    opt.zero_grad()
    w0o = oracle.latent_from_alpha(sp["u_cap"], a, mean)
    wxo = oracle.embed_fingerprint(sp["v_cap"], sp["sigma_key"], torch.sigmoid(k), w0o, 1.0).requires_grad_(True) if False else oracle.embed_fingerprint(sp["v_cap"], sp["sigma_key"], torch.sigmoid(k), w0o, 1.0)
    wxo.retain_grad()
    est = oracle.generator_forward(params, [wxo.reshape(1, -1)], size, input_is_latent=True, noise=noise)
    mse_o = oracle.mse_loss(target, est)
    loss = mse_o + 0.1 * oracle.alpha_bound(a, sp["max_alpha"], sp["min_alpha"])
    opt.param_groups[0]["lr"] = get_lr(i)
    loss.backward()
    gw = wxo.grad.reshape(-1)
    rel = float((dwx_e[0].cpu() - gw).norm() / gw.norm())
    eng.step(st, target.cuda())
    print(f"step {i}: forced mse rel {abs(float(mse_e[0]) - float(mse_o)) / float(mse_o):.2e} dwx rel {rel:.2e} | "
          f"free loss rel {abs(float(st['loss'][0]) - float(loss)) / float(loss):.2e} "
          f"min|g_alpha| {float(a.grad.abs().min()):.2e}")
    opt.step()
    print(f"        free alpha maxdiff {float((st['alpha'][0].cpu() - a.detach()[:, 0]).abs().max()):.2e} key maxdiff {float((st['key'][0].cpu() - k.detach()[:, 0]).abs().max()):.2e}")
