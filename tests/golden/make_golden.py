"""Generate tests/golden/*.npz by running the UNMODIFIED reference on CPU.

Run in the build container only (needs /root/reference):

    python tests/golden/make_golden.py

The reference is imported in place from /root/reference/src.  Three things are
stubbed, none of which is on the CPU code path being recorded:
  * ``torch.utils.cpp_extension.load`` (the reference JIT-builds its CUDA ops at
    import, src/op/upfirdn2d.py:11-17, src/op/fused_act.py:11-17; on CPU tensors it
    never calls them),
  * ``custom_lpips`` (needs skimage / pip lpips / downloaded VGG weights, all
    absent offline): replaced by an MSE "perceptual" loss, stated in the fixture,
  * ``scipy`` ``LatinHypercube(centered=True)`` (rejected by scipy 1.18): the LHS
    samples are supplied explicitly.
Inputs are derived from seeds via tests/fixtures.py, so the tests can rebuild
them without the reference.
"""
import os
import sys
import types

import numpy as np
import torch
import torch.nn.functional as F

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import fixtures as fx  # noqa: E402

REF_SRC = "/root/reference/src"


def import_reference():
    import torch.utils.cpp_extension as ce
    ce.load = lambda *a, **k: None
    lp = types.ModuleType("custom_lpips")

    class PerceptualLoss:  # stand-in: MSE (src/utils.py:46-47 is the reference's own MSE option)
        def __init__(self, *a, **k):
            pass

        def __call__(self, a, b):
            return F.mse_loss(a, b)

    lp.PerceptualLoss = PerceptualLoss
    sys.modules["custom_lpips"] = lp
    sys.path.insert(0, REF_SRC)
    sys.argv = ["main.py", "--model", "sg2", "--img_size", "32", "--steps", "12", "--n", "2",
                "--key_len", "64", "--shift", "448", "--sigma", "1.0"]
    import params
    params.opt.device = "cpu"
    import model
    import op
    from op.upfirdn2d import upfirdn2d_native
    import utils as rutils
    import generator as rgen
    import main as rmain
    return dict(model=model, op=op, native=upfirdn2d_native, utils=rutils, gen=rgen, main=rmain,
                opt=params.opt)


UPFIRDN_CASES = [
    # name, shape, kernel spec, up, down, pad
    ("blur_pad11_odd", (2, 3, 9, 9), "fir4x4_g4", 1, 1, (1, 1)),
    ("blur_bwd_pad22", (2, 3, 8, 8), "fir4x4_g4", 1, 1, (2, 2)),
    ("up2_pad21_nonsquare", (1, 3, 5, 7), "fir4x4_g4", 2, 1, (2, 1)),
    ("down2_pad11", (2, 2, 8, 10), "fir4x4", 1, 2, (1, 1)),
    ("up2_asym4x4", (1, 2, 6, 5), "asym4x4", 2, 1, (2, 1)),
    ("down2_asym4x4", (1, 2, 9, 8), "asym4x4", 1, 2, (1, 1)),
    ("blur_asym4x4", (1, 2, 7, 6), "asym4x4", 1, 1, (1, 2)),
    ("k3x3_mode2", (1, 2, 6, 6), "asym3x3", 1, 1, (1, 1)),
    ("k2x2_up2_mode4", (1, 2, 4, 5), "asym2x2", 2, 1, (1, 0)),
    ("k2x2_down2_mode6", (1, 2, 8, 6), "asym2x2", 1, 2, (0, 0)),
    ("neg_pad_crop", (1, 2, 9, 9), "asym4x4", 1, 1, (-1, 2, 1, -1)),
    ("mixed_updown_xy", (1, 2, 6, 7), "asym3x4", (2, 1), (1, 2), (1, 2, 2, 1)),
    ("up3_down2_k5", (1, 1, 7, 6), "asym5x5", 3, 2, (3, 2)),
    ("many_planes", (3, 70, 5, 5), "fir4x4_g4", 1, 1, (1, 1)),
]


def kernel_of(spec):
    if spec == "fir4x4":
        k = np.outer([1, 3, 3, 1], [1, 3, 3, 1]).astype(np.float32)
        return torch.from_numpy(k / k.sum())
    if spec == "fir4x4_g4":
        return kernel_of("fir4x4") * 4
    if spec.startswith("asym"):
        kh, kw = (int(c) for c in spec[4:].split("x"))
        k = (np.arange(kh * kw, dtype=np.float32).reshape(kh, kw) + 1.0)
        k = k * np.linspace(0.5, 1.5, kw, dtype=np.float32)[None, :]
        return torch.from_numpy((k / k.sum()).astype(np.float32))
    raise KeyError(spec)


def gen_upfirdn(ref, out):
    for i, (name, shape, kspec, up, down, pad) in enumerate(UPFIRDN_CASES):
        x = fx.seeded(shape, 100 + i)
        k = kernel_of(kspec)
        up2 = (up, up) if isinstance(up, int) else up
        dn2 = (down, down) if isinstance(down, int) else down
        pad4 = (pad[0], pad[1], pad[0], pad[1]) if len(pad) == 2 else pad
        y = ref["native"](x, k, *up2, *dn2, *pad4)
        # public entry point must agree with the native helper on CPU tensors
        y2 = ref["op"].upfirdn2d(x, k, up=up, down=down, pad=pad)
        assert torch.equal(y, y2)
        out[f"upfirdn/{name}/y"] = y.numpy()
        # gradient wrt input for a seeded cotangent
        xg = x.clone().requires_grad_(True)
        yg = ref["op"].upfirdn2d(xg, k, up=up, down=down, pad=pad)
        ct = fx.seeded(tuple(yg.shape), 200 + i)
        (gx,) = torch.autograd.grad((yg * ct).sum(), xg)
        out[f"upfirdn/{name}/gx"] = gx.numpy()


LRELU_CASES = [
    ("rank4_bias", (2, 5, 4, 4), True),
    ("rank2_bias", (3, 7), True),
    ("rank4_nobias", (2, 3, 2, 2), False),
    ("rank3_bias", (2, 4, 6), True),
]


def gen_lrelu(ref, out):
    for i, (name, shape, use_bias) in enumerate(LRELU_CASES):
        x = fx.seeded(shape, 300 + i)
        x.view(-1)[::5] = 0.0  # exact zeros: x>0 convention
        b = fx.seeded((shape[1],), 320 + i) if use_bias else None
        if b is not None:
            b[0] = 0.0
        xg = x.clone().requires_grad_(True)
        bg = b.clone().requires_grad_(True) if b is not None else None
        y = ref["op"].fused_leaky_relu(xg, bg)
        ct = fx.seeded(tuple(y.shape), 340 + i)
        grads = torch.autograd.grad((y * ct).sum(), [xg] + ([bg] if bg is not None else []))
        out[f"lrelu/{name}/y"] = y.detach().numpy()
        out[f"lrelu/{name}/gx"] = grads[0].numpy()
        if b is not None:
            out[f"lrelu/{name}/gb"] = grads[1].numpy()


MODCONV_CASES = [
    # name, B, Cin, Cout, H, W, k, demod, up, style_dim
    ("plain3x3", 2, 8, 6, 5, 7, 3, True, False, 16),
    ("up3x3", 2, 8, 4, 4, 4, 3, True, True, 16),
    ("rgb1x1", 2, 8, 3, 6, 6, 1, False, False, 16),
    ("plain3x3_b1", 1, 16, 16, 8, 8, 3, True, False, 32),
    ("up3x3_nonsq", 1, 6, 10, 3, 5, 3, True, True, 16),
]


def modconv_inputs(i, B, Cin, Cout, H, W, k, sd):
    return dict(
        x=fx.seeded((B, Cin, H, W), 400 + i),
        style=fx.seeded((B, sd), 420 + i),
        weight=fx.seeded((1, Cout, Cin, k, k), 440 + i),
        mod_w=fx.seeded((Cin, sd), 460 + i),
        mod_b=torch.ones(Cin) + fx.seeded((Cin,), 480 + i, scale=0.1),
    )


def gen_modconv(ref, out):
    M = ref["model"]
    for i, (name, B, Cin, Cout, H, W, k, demod, up, sd) in enumerate(MODCONV_CASES):
        t = modconv_inputs(i, B, Cin, Cout, H, W, k, sd)
        for fused in (True, False):
            m = M.ModulatedConv2d(Cin, Cout, k, sd, demodulate=demod, upsample=up, fused=fused)
            with torch.no_grad():
                m.weight.copy_(t["weight"])
                m.modulation.weight.copy_(t["mod_w"])
                m.modulation.bias.copy_(t["mod_b"])
            x = t["x"].clone().requires_grad_(True)
            s = t["style"].clone().requires_grad_(True)
            y = m(x, s)
            ct = fx.seeded(tuple(y.shape), 500 + i)
            gx, gs = torch.autograd.grad((y * ct).sum(), [x, s])
            tag = "fused" if fused else "unfused"
            out[f"modconv/{name}/{tag}/y"] = y.detach().numpy()
            out[f"modconv/{name}/{tag}/gx"] = gx.numpy()
            out[f"modconv/{name}/{tag}/gs"] = gs.numpy()


def ref_generator(ref, size, seed, cm=2):
    g = ref["model"].Generator(size, 512, 8, channel_multiplier=cm)
    missing = g.load_state_dict(fx.make_params(size, seed, cm), strict=False)
    assert not missing.unexpected_keys, missing
    # only buffers (blur kernels, stored noises) may be missing
    assert all(("kernel" in k or k.startswith("noises.")) for k in missing.missing_keys), missing
    return g.eval()


GEN_CASES = [("g32_b2", 32, 2, 2, 11), ("g64_b1", 64, 2, 1, 12), ("g16_cm1_b3", 16, 1, 3, 13)]


def gen_generator(ref, out):
    for name, size, cm, B, seed in GEN_CASES:
        g = ref_generator(ref, size, seed, cm)
        noise = fx.make_noise(size, seed + 1)
        w = fx.seeded((B, 512), seed + 2).requires_grad_(True)
        img, _ = g([w], input_is_latent=True, noise=noise)
        ct = fx.seeded(tuple(img.shape), seed + 3)
        (gw,) = torch.autograd.grad((img * ct).sum(), w)
        out[f"gen/{name}/img"] = img.detach().numpy()
        out[f"gen/{name}/gw"] = gw.numpy()
        z = fx.seeded((3, 512), seed + 4)
        with torch.no_grad():
            out[f"gen/{name}/mapping"] = g.style(z).numpy()
            img_z, lat = g([z[:B]], noise=noise, return_latents=True)
        out[f"gen/{name}/img_from_z"] = img_z.numpy()
        out[f"gen/{name}/latent_shape"] = np.array(lat.shape)


def gen_state_dict_keys(ref):
    """Names and shapes of Generator.state_dict() for the checkpoint-compatibility test."""
    import json
    table = {}
    for size, cm in ((32, 2), (256, 2), (1024, 2), (64, 1)):
        g = ref["model"].Generator(size, 512, 8, channel_multiplier=cm)
        table[f"{size}_cm{cm}"] = {k: list(v.shape) for k, v in g.state_dict().items()}
    with open(os.path.join(HERE, "state_dict_keys.json"), "w") as f:
        json.dump(table, f, indent=0, sort_keys=True)
    print("state_dict_keys.json", {k: len(v) for k, v in table.items()})


def gen_embed_and_loop(ref, out):
    size, seed = 32, 11
    g = ref_generator(ref, size, seed)
    for p in g.parameters():  # as in the reference: parameters keep requires_grad
        p.requires_grad_(True)
    noise = fx.make_noise(size, seed + 1)
    pc, sigma, mean = fx.make_pca_basis(2)
    sp = fx.split_basis(pc, sigma, 64, 448, 1.0)
    G = ref["gen"].GetGen
    fake = types.SimpleNamespace(sd_moved=1, key_len=64, batch_size=1, device="cpu", model="sg2",
                                 style_mixing=False, g_ema=g, latent_mean=mean, style_space_dim=512,
                                 num_main_pc=448)
    fake.get_new_latent = types.MethodType(G.get_new_latent, fake)
    fake.generate_image = types.MethodType(G.generate_image, fake)
    # get_new_latent (src/generator.py:148-161)
    k = torch.sigmoid(fx.seeded((64, 1), 31))
    w0 = fx.seeded((512, 1), 32)
    out["embed/get_new_latent"] = fake.get_new_latent(sp["v_cap"], sp["sigma_key"], k, w0).numpy()
    # generate_with_alpha (src/generator.py:69-107)
    torch.manual_seed(5)
    alpha = sp["sigma_main"] * fx.seeded((448, 1), 33)
    img, w0_t, wx_t, key = G.generate_with_alpha(fake, alpha, sp["u_cap"].t(), sp["sigma_key"],
                                                 sp["v_cap"], noise)
    out["embed/gwa_img"] = img.numpy()
    out["embed/gwa_w0"] = w0_t.numpy()
    out["embed/gwa_wx"] = wx_t.numpy()
    out["embed/gwa_key"] = key.numpy()
    # alpha_bound / get_lr / get_noise
    a = fx.seeded((448, 1), 34, scale=2.0)
    out["embed/alpha_bound"] = ref["utils"].alpha_bound(a, sp["max_alpha"], sp["min_alpha"]).numpy()
    out["embed/get_lr"] = np.array([ref["main"].get_lr(i) for i in (0, 1, 99, 1999)])
    np.random.seed(2022)
    ref["opt"].img_size = 32
    nz = ref["utils"].get_noise()
    out["embed/get_noise_head"] = np.stack([n.reshape(-1)[:8].numpy() for n in nz])

    # the reference loop itself: main.optimization (src/main.py:45-89), MSE stand-in loss
    rm = ref["main"]
    lhs = np.stack([(np.random.RandomState(40 + j).permutation(2) + 0.5) / 2 for j in range(448)], 1)
    rm.samlping = types.SimpleNamespace(random=lambda n: lhs)
    rm.sigma_448 = sp["sigma_main"]
    rm.generator = fake
    fake.key = key
    rm.u_cap, rm.v_cap, rm.sigma_64 = sp["u_cap"], sp["v_cap"], sp["sigma_key"]
    rm.noise = noise
    rm.sigmoid = torch.nn.Sigmoid()
    rm.max_alpha, rm.min_alpha = sp["max_alpha"], sp["min_alpha"]
    rm.target_w0 = w0_t
    rm.loss, rm.a, rm.k = [], [], []
    alpha_best, key_best, acc = rm.optimization(img)
    out["loop/lhs"] = lhs.astype(np.float32)
    out["loop/loss"] = np.array(rm.loss, dtype=np.float64)
    out["loop/alpha"] = np.stack([t.detach().numpy() for t in rm.a])
    out["loop/key"] = np.stack([t.detach().numpy() for t in rm.k])
    out["loop/best_alpha"] = alpha_best.detach().numpy()
    out["loop/best_key"] = key_best.detach().numpy()
    out["loop/acc"] = np.array(float(acc))


def main():
    torch.set_num_threads(8)
    ref = import_reference()
    gen_state_dict_keys(ref)
    for fname, fn in [("ops.npz", lambda o: (gen_upfirdn(ref, o), gen_lrelu(ref, o))),
                      ("modconv.npz", lambda o: gen_modconv(ref, o)),
                      ("generator.npz", lambda o: gen_generator(ref, o)),
                      ("attribution.npz", lambda o: gen_embed_and_loop(ref, o))]:
        out = {}
        fn(out)
        path = os.path.join(HERE, fname)
        np.savez_compressed(path, **out)
        print(fname, len(out), "arrays", os.path.getsize(path) // 1024, "KiB")


if __name__ == "__main__":
    main()
