"""Pin the CPU oracle against vectors recorded from the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import math
import os
import sys

import numpy as np
import pytest
import torch

import fixtures as fx
import oracle
from golden.make_golden import (GEN_CASES, LRELU_CASES, MODCONV_CASES, UPFIRDN_CASES, kernel_of,
                                modconv_inputs)


@pytest.mark.parametrize("i", range(len(UPFIRDN_CASES)))
def test_upfirdn2d_matches_reference(golden, i):
    name, shape, kspec, up, down, pad = UPFIRDN_CASES[i]
    x = fx.seeded(shape, 100 + i).requires_grad_(True)
    k = kernel_of(kspec)
    y = oracle.upfirdn2d(x, k, up=up, down=down, pad=pad)
    ref = golden[f"upfirdn/{name}/y"]
    assert tuple(y.shape) == ref.shape
    np.testing.assert_allclose(y.detach().numpy(), ref, rtol=1e-5, atol=1e-6)
    ct = fx.seeded(tuple(y.shape), 200 + i)
    (gx,) = torch.autograd.grad((y * ct).sum(), x)
    np.testing.assert_allclose(gx.numpy(), golden[f"upfirdn/{name}/gx"], rtol=1e-5, atol=1e-6)
    # the per-sample loop definition agrees too
    if np.prod(shape) < 3000:
        yl = oracle.upfirdn2d_loops(x.detach().numpy(), k.numpy(), up=up, down=down, pad=pad)
        np.testing.assert_allclose(yl, ref, rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("i", range(len(LRELU_CASES)))
def test_fused_leaky_relu_matches_reference(golden, i):
    name, shape, use_bias = LRELU_CASES[i]
    x = fx.seeded(shape, 300 + i)
    x.view(-1)[::5] = 0.0
    b = fx.seeded((shape[1],), 320 + i) if use_bias else None
    if b is not None:
        b[0] = 0.0
    xg = x.clone().requires_grad_(True)
    bg = b.clone().requires_grad_(True) if b is not None else None
    y = oracle.fused_leaky_relu(xg, bg)
    np.testing.assert_allclose(y.detach().numpy(), golden[f"lrelu/{name}/y"], rtol=1e-6, atol=1e-7)
    ct = fx.seeded(tuple(y.shape), 340 + i)
    grads = torch.autograd.grad((y * ct).sum(), [xg] + ([bg] if bg is not None else []))
    np.testing.assert_allclose(grads[0].numpy(), golden[f"lrelu/{name}/gx"], rtol=1e-6, atol=1e-7)
    if b is not None:
        np.testing.assert_allclose(grads[1].numpy(), golden[f"lrelu/{name}/gb"], rtol=1e-5, atol=1e-6)
    # native-op backward (act=3, grad=1) reproduces autograd's grad_input from the saved output
    gi = oracle.fused_bias_act(ct, None, y.detach(), 3, 1, 0.2, 2 ** 0.5)
    np.testing.assert_allclose(gi.numpy(), golden[f"lrelu/{name}/gx"], rtol=1e-6, atol=1e-7)


@pytest.mark.parametrize("i", range(len(MODCONV_CASES)))
@pytest.mark.parametrize("algebra", ["fused", "unfused"])
def test_modulated_conv_matches_reference(golden, i, algebra):
    name, B, Cin, Cout, H, W, k, demod, up, sd = MODCONV_CASES[i]
    t = modconv_inputs(i, B, Cin, Cout, H, W, k, sd)
    x = t["x"].clone().requires_grad_(True)
    s = t["style"].clone().requires_grad_(True)
    fn = oracle.modulated_conv2d if algebra == "fused" else oracle.modulated_conv2d_unfused
    y = fn(x, s, t["weight"], t["mod_w"], t["mod_b"], demodulate=demod, upsample=up)
    ct = fx.seeded(tuple(y.shape), 500 + i)
    gx, gs = torch.autograd.grad((y * ct).sum(), [x, s])
    for got, key in ((y.detach(), "y"), (gx, "gx"), (gs, "gs")):
        ref = golden[f"modconv/{name}/{algebra}/{key}"]
        scale = np.abs(ref).max()
        np.testing.assert_allclose(got.numpy(), ref, rtol=1e-4, atol=2e-5 * scale)
    # and the two algebras agree with each other (reference fused vs unfused noise floor)
    other = golden[f"modconv/{name}/{'unfused' if algebra == 'fused' else 'fused'}/y"]
    np.testing.assert_allclose(y.detach().numpy(), other, rtol=1e-3, atol=1e-4 * np.abs(other).max())


@pytest.mark.parametrize("case", GEN_CASES, ids=[c[0] for c in GEN_CASES])
def test_generator_matches_reference(golden, case):
    name, size, cm, B, seed = case
    params = fx.make_params(size, seed, cm)
    noise = fx.make_noise(size, seed + 1)
    w = fx.seeded((B, 512), seed + 2).requires_grad_(True)
    img = oracle.generator_forward(params, [w], size, input_is_latent=True, noise=noise)
    ref = golden[f"gen/{name}/img"]
    assert np.abs(img.detach().numpy() - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())
    ct = fx.seeded(tuple(img.shape), seed + 3)
    (gw,) = torch.autograd.grad((img * ct).sum(), w)
    gref = golden[f"gen/{name}/gw"]
    assert np.linalg.norm(gw.numpy() - gref) <= 1e-4 * np.linalg.norm(gref)
    z = fx.seeded((3, 512), seed + 4)
    with torch.no_grad():
        m = oracle.mapping(params, z)
        img_z = oracle.generator_forward(params, [z[:B]], size, noise=noise)
    np.testing.assert_allclose(m.numpy(), golden[f"gen/{name}/mapping"], rtol=1e-4, atol=1e-5)
    zref = golden[f"gen/{name}/img_from_z"]
    assert np.abs(img_z.numpy() - zref).max() <= 1e-4 * max(1.0, np.abs(zref).max())
    assert tuple(golden[f"gen/{name}/latent_shape"]) == (B, oracle.n_latent(size), 512)
    # unfused algebra (the one the CUDA path uses) stays within the reference's own noise floor
    with torch.no_grad():
        img_u = oracle.generator_forward(params, [w.detach()], size, input_is_latent=True, noise=noise,
                                         fused=False)
    assert np.abs(img_u.numpy() - ref).max() <= 1e-3 * max(1.0, np.abs(ref).max())


def test_embed_and_helpers_match_reference(golden):
    pc, sigma, mean = fx.make_pca_basis(2)
    sp = fx.split_basis(pc, sigma, 64, 448, 1.0)
    k = torch.sigmoid(fx.seeded((64, 1), 31))
    w0 = fx.seeded((512, 1), 32)
    wx = oracle.embed_fingerprint(sp["v_cap"], sp["sigma_key"], k, w0, 1)
    np.testing.assert_allclose(wx.numpy(), golden["embed/get_new_latent"], rtol=1e-6, atol=1e-6)
    a = fx.seeded((448, 1), 34, scale=2.0)
    ab = oracle.alpha_bound(a, sp["max_alpha"], sp["min_alpha"])
    np.testing.assert_allclose(ab.numpy(), golden["embed/alpha_bound"], rtol=1e-6)
    np.testing.assert_allclose([oracle.lr_at(i) for i in (0, 1, 99, 1999)], golden["embed/get_lr"],
                               rtol=1e-12)
    np.random.seed(2022)
    nz = oracle.get_noise(32)
    head = np.stack([n.reshape(-1)[:8].numpy() for n in nz])
    np.testing.assert_array_equal(head, golden["embed/get_noise_head"])


def test_generate_with_alpha_matches_reference(golden):
    size, seed = 32, 11
    params = fx.make_params(size, seed)
    noise = fx.make_noise(size, seed + 1)
    pc, sigma, mean = fx.make_pca_basis(2)
    sp = fx.split_basis(pc, sigma, 64, 448, 1.0)
    alpha = sp["sigma_main"] * fx.seeded((448, 1), 33)
    key = torch.from_numpy(golden["embed/gwa_key"])
    with torch.no_grad():
        img, w0, wx = oracle.generate_with_alpha(params, size, alpha, sp["u_cap"], sp["v_cap"],
                                                 sp["sigma_key"], mean, key, noise, sd=1)
    np.testing.assert_allclose(w0.numpy(), golden["embed/gwa_w0"], rtol=1e-5, atol=1e-6)
    np.testing.assert_allclose(wx.numpy(), golden["embed/gwa_wx"], rtol=1e-5, atol=1e-6)
    ref = golden["embed/gwa_img"]
    assert np.abs(img.numpy() - ref).max() <= 1e-4 * max(1.0, np.abs(ref).max())


def test_attribution_loop_matches_reference(golden):
    """The oracle's restated loop against the reference's own ``main.optimization`` (12 steps,
    2 guesses, 32 px, MSE stand-in loss)."""
    size, seed = 32, 11
    params = fx.make_params(size, seed)
    noise = fx.make_noise(size, seed + 1)
    pc, sigma, mean = fx.make_pca_basis(2)
    sp = fx.split_basis(pc, sigma, 64, 448, 1.0)
    target = torch.from_numpy(golden["embed/gwa_img"])
    guesses = torch.from_numpy(golden["loop/lhs"])

    def render(wx):
        return oracle.generator_forward(params, [wx.reshape(1, -1)], size, input_is_latent=True,
                                        noise=noise)

    best, results = oracle.attribute_image(render, target, guesses, sp["u_cap"], sp["v_cap"],
                                           sp["sigma_key"], sp["sigma_main"], mean, sp["max_alpha"],
                                           sp["min_alpha"], steps=12)
    losses = np.array([r[0] for r in results])
    np.testing.assert_allclose(losses, golden["loop/loss"], rtol=2e-4)
    for j, r in enumerate(results):
        np.testing.assert_allclose(r[1].numpy(), golden["loop/alpha"][j], rtol=0, atol=2e-3)
        np.testing.assert_allclose(r[2].numpy(), golden["loop/key"][j], rtol=0, atol=2e-3)
    np.testing.assert_allclose(best[2].numpy(), golden["loop/best_key"], rtol=0, atol=2e-3)
    true_key = torch.from_numpy(golden["embed/gwa_key"]).float()
    acc = (oracle.decode_key(best[2]) == true_key).float().mean().item()
    assert abs(acc - float(golden["loop/acc"])) < 1e-6


def test_lhs_centered_is_latin():
    rng = np.random.default_rng(0)
    s = oracle.latin_hypercube_centered(20, 7, rng)
    assert s.shape == (20, 7)
    for d in range(7):
        assert sorted(np.round(s[:, d] * 20 - 0.5).astype(int).tolist()) == list(range(20))


def test_lpips_oracle_matches_reference_pnetlin():
    """oracle/lpips_oracle.py against the reference's own PNetLin (tests/golden/make_golden_lpips.py: random VGG16 weights,
    shipped linear heads): distance and its gradient w.r.t. the second image."""
    import os
    from oracle import lpips_oracle as lo
    from golden.make_golden_lpips import CASES
    with np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "lpips.npz")) as z:
        g = {k: z[k] for k in z.files}
    heads = {f"lin{k}.model.1.weight": torch.from_numpy(g[f"lpips/head{k}"]) for k in range(5)}
    params = lo.make_vgg_params(seed=5, lin_weights=heads)
    for name, B, size, seed in CASES:
        h, w = (size, size) if isinstance(size, int) else size
        a = fx.seeded((B, 3, h, w), seed, scale=0.5)
        b = fx.seeded((B, 3, h, w), seed + 100, scale=0.5).requires_grad_(True)
        val = lo.lpips(params, a, b)
        (gb,) = torch.autograd.grad(val.sum(), b)
        np.testing.assert_allclose(val.detach().numpy(), g[f"lpips/{name}/val"], rtol=1e-5, atol=1e-8)
        np.testing.assert_allclose(gb.numpy(), g[f"lpips/{name}/grad_in1"], rtol=1e-4, atol=1e-6 * np.abs(g[f"lpips/{name}/grad_in1"]).max())
