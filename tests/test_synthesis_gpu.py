"""GPU parity of the fused synthesis path (Generator.forward -> lfp_synth_forward/backward) against
the reference-generated golden vectors and the CPU oracle.

Tolerances (fp32 path): images are compared on their own scale - random-init images are not in
[-1,1] - so the north-star bar "max-abs <= 1e-3 on [-1,1] pixels" is applied as
max-abs <= 1e-3 * max(1, max|ref|) (asserted at 1e-4; measured ~1e-5).

Latent gradients: leaky-ReLU has a kink at 0.  A unit whose pre-activation is at rounding level
(|v| ~ 1e-6; there are a few per 10^5 units) lands on either side depending on summation order,
which changes its derivative from 1 to 0.2 and the latent gradient by ~1e-3 relative - the
reference's own fused vs unfused algebra differ by exactly that much on the same fixtures
(measured on CPU: 1 flipped unit -> 1.1e-3).  So the gradient bar is split in two:
  * `grad_parity`: with the branch pattern of the CUDA path imposed on the oracle
    (oracle.lrelu_with_sign), gradients agree to 5e-5 relative L2 (pure arithmetic parity);
  * the branch patterns themselves may differ only at units with |activation| < 1e-4.
Free-running comparisons (golden vectors from the reference, Generator-level tests) use 5e-3."""
import numpy as np
import pytest
import torch

import fixtures as fx
import oracle
from golden.make_golden import GEN_CASES

pytestmark = pytest.mark.gpu
DEV = "cuda"


def build_generator(size, seed, cm=2):
    from model import Generator
    g = Generator(size, 512, 8, channel_multiplier=cm)
    missing = g.load_state_dict(fx.make_params(size, seed, cm), strict=False)
    assert not missing.unexpected_keys
    return g.eval().to(DEV)


KINK_TOL = 5e-3   # free-running gradient tolerance (see module docstring)


def grad_parity(size, seed, lat, noise, cm=2, precision=None, img_tol=1e-4, grad_tol=5e-5, flip_band=1e-4, act_tol=None):
    """Plan-level forward/backward against the oracle with the CUDA path's lrelu branches imposed."""
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    precision = capi.PREC_FP32 if precision is None else precision
    act_tol = img_tol if act_tol is None else act_tol
    params = fx.make_params(size, seed, cm)
    plan = SynthesisPlan(size, channel_multiplier=cm, device=DEV)
    plan.load(params)
    B = lat.shape[0]
    ws = plan.new_workspace(B)
    img = plan.forward(lat.to(DEV), [n.to(DEV) for n in noise], ws, precision)
    acts = [a.cpu() for a in plan.read_activations(B, ws)]
    ct = fx.seeded(tuple(img.shape), seed + 3)
    d_lat = plan.backward(ct.to(DEV), B, ws, precision).cpu()
    # oracle, free-running: image and activations
    ref_acts = []
    with torch.no_grad():
        ref = oracle.synthesis(params, lat, noise, activations=ref_acts)
    img_close(img.cpu().numpy(), ref.numpy(), img_tol)
    flips = 0
    for li, (a, r) in enumerate(zip(acts, ref_acts)):
        # per-layer comparison: a failure names the StyledConv (0 = conv1, 1 + i = convs.i) that first went wrong
        assert a.shape == r.shape, f"conv {li}: shape {tuple(a.shape)} vs {tuple(r.shape)}"
        scale = max(1.0, float(r.abs().max()))
        err = float((a - r).abs().max())
        assert err <= act_tol * scale, f"StyledConv {li} ({r.shape[1]} ch @ {r.shape[2]} px): max-abs {err:.3e} vs scale {scale:.3e}"
        diff = (a > 0) != (r > 0)
        flips += int(diff.sum())
        if diff.any():   # branches may differ only where the activation is numerically zero
            assert float(r[diff].abs().max()) < flip_band * scale, f"StyledConv {li}: lrelu branch differs away from 0"
    # oracle with the CUDA path's branch pattern: arithmetic parity of the gradient
    lr = lat.clone().requires_grad_(True)
    refm = oracle.synthesis(params, lr, noise, sign_masks=[a > 0 for a in acts])
    (gref,) = torch.autograd.grad((refm * ct).sum(), lr)
    rel = float((d_lat - gref).norm() / gref.norm())
    assert rel <= grad_tol, (rel, flips)
    per_slot = ((d_lat - gref).flatten(2).norm(dim=2) / gref.flatten(2).norm(dim=2)).max()
    assert float(per_slot) <= 4 * grad_tol, float(per_slot)
    return rel, flips


def img_close(a, ref, tol=1e-3):
    err = np.abs(a - ref).max()
    assert err <= tol * max(1.0, np.abs(ref).max()), f"max-abs {err} vs scale {np.abs(ref).max()}"
    return err


@pytest.mark.parametrize("case", GEN_CASES, ids=[c[0] for c in GEN_CASES])
def test_generator_golden(golden, case):
    name, size, cm, B, seed = case
    g = build_generator(size, seed, cm)
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    w = fx.seeded((B, 512), seed + 2).to(DEV).requires_grad_(True)
    img, none = g([w], input_is_latent=True, noise=noise)
    assert none is None and img.shape == (B, 3, size, size)
    ref = golden[f"gen/{name}/img"]
    img_close(img.detach().cpu().numpy(), ref, 1e-4)
    ct = fx.seeded(tuple(img.shape), seed + 3).to(DEV)
    (gw,) = torch.autograd.grad((img * ct).sum(), w)
    gref = golden[f"gen/{name}/gw"]
    rel = np.linalg.norm(gw.cpu().numpy() - gref) / np.linalg.norm(gref)
    assert rel <= KINK_TOL, rel
    # mapping network + z input
    z = fx.seeded((3, 512), seed + 4).to(DEV)
    with torch.no_grad():
        np.testing.assert_allclose(g.style(z).cpu().numpy(), golden[f"gen/{name}/mapping"], rtol=1e-3, atol=1e-4)
        img_z, lat = g([z[:B]], noise=noise, return_latents=True)
    img_close(img_z.cpu().numpy(), golden[f"gen/{name}/img_from_z"], 1e-3)
    assert tuple(lat.shape) == tuple(golden[f"gen/{name}/latent_shape"])


@pytest.mark.parametrize("size,B", [(8, 1), (16, 5), (128, 2), (256, 1)])
def test_generator_vs_oracle(size, B):
    seed = 20 + size
    params = fx.make_params(size, seed)
    g = build_generator(size, seed)
    noise = fx.make_noise(size, seed + 1)
    w = fx.seeded((B, 512), seed + 2)
    wr = w.clone().requires_grad_(True)
    ref = oracle.generator_forward(params, [wr], size, input_is_latent=True, noise=noise)
    ct = fx.seeded(tuple(ref.shape), seed + 3)
    (gref,) = torch.autograd.grad((ref * ct).sum(), wr)
    wg = w.to(DEV).requires_grad_(True)
    img, _ = g([wg], input_is_latent=True, noise=[n.to(DEV) for n in noise])
    img_close(img.detach().cpu().numpy(), ref.detach().numpy(), 1e-4)
    (gw,) = torch.autograd.grad((img * ct.to(DEV)).sum(), wg)
    rel = np.linalg.norm(gw.cpu().numpy() - gref.numpy()) / np.linalg.norm(gref.numpy())
    assert rel <= KINK_TOL, rel
    # arithmetic parity proper: per-slot latents, CUDA branch pattern imposed on the oracle
    lat = fx.seeded((B, oracle.n_latent(size), 512), seed + 5)
    grad_parity(size, seed, lat, noise)


@pytest.mark.parametrize("size,cm,B", [(64, 1, 2), (32, 2, 4)])
def test_gradient_parity_with_imposed_branches(size, cm, B):
    seed = 90 + size
    noise = fx.make_noise(size, seed + 1, batch=B)
    lat = fx.seeded((B, oracle.n_latent(size), 512), seed + 2)
    grad_parity(size, seed, lat, noise, cm=cm)


@pytest.mark.parametrize("size,cm,B", [(64, 2, 2), (128, 1, 1), (256, 2, 1)])
def test_tf32_tensor_core_path_parity(size, cm, B):
    """tcgen05 kind::tf32 path (operands rounded to 10-bit mantissa, fp32 accumulation in TMEM) - the
    numerics the reference itself runs on GPU (cuDNN TF32 convs).  Stated tolerance:
    image max-abs <= 5e-3 * max(1, max|ref|); latent gradient <= 1e-2 relative L2 with the CUDA path's
    leaky-ReLU branches imposed on the fp32 oracle; branches may differ where |activation| < 2e-2."""
    from lfp_native import capi
    seed = 120 + size
    noise = fx.make_noise(size, seed + 1)
    lat = fx.seeded((B, oracle.n_latent(size), 512), seed + 2)
    rel, flips = grad_parity(size, seed, lat, noise, cm=cm, precision=capi.PREC_TF32, img_tol=5e-3, grad_tol=1e-2,
                             flip_band=2e-2)
    print(f"tf32 size {size}: gradient rel err {rel:.2e}, {flips} branch flips")


def test_per_sample_noise_style_mixing_and_per_slot_gradient():
    size, seed, B = 32, 31, 3
    params = fx.make_params(size, seed)
    g = build_generator(size, seed)
    noise = fx.make_noise(size, seed + 1, batch=B)           # per-sample noise maps [B,1,h,h]
    lat = fx.seeded((B, oracle.n_latent(size), 512), seed + 2)   # a different latent per slot
    lr = lat.clone().requires_grad_(True)
    ref = oracle.synthesis(params, lr, noise)
    ct = fx.seeded(tuple(ref.shape), seed + 3)
    (gref,) = torch.autograd.grad((ref * ct).sum(), lr)
    lg = lat.to(DEV).requires_grad_(True)
    img, _ = g([lg], input_is_latent=True, noise=[n.to(DEV) for n in noise])
    img_close(img.detach().cpu().numpy(), ref.detach().numpy(), 1e-4)
    (gl,) = torch.autograd.grad((img * ct.to(DEV)).sum(), lg)
    assert gl.shape == lat.shape
    rel = np.linalg.norm(gl.cpu().numpy() - gref.numpy()) / np.linalg.norm(gref.numpy())
    assert rel <= KINK_TOL, rel
    grad_parity(size, seed, lat, noise)
    # style mixing goes through the same latent assembly as the reference (src/model.py:536-548)
    w1, w2 = fx.seeded((B, 512), 40), fx.seeded((B, 512), 41)
    mixed = torch.cat([w1[:, None].repeat(1, 4, 1), w2[:, None].repeat(1, oracle.n_latent(size) - 4, 1)], 1)
    with torch.no_grad():
        img_mix, _ = g([w1.to(DEV), w2.to(DEV)], input_is_latent=True, noise=[n.to(DEV) for n in noise], inject_index=4)
        ref_mix = oracle.synthesis(params, mixed, noise)
    img_close(img_mix.cpu().numpy(), ref_mix.numpy(), 1e-4)


def test_batch_independence_is_bitwise():
    """A trajectory's image and gradient do not depend on what shares its batch (what makes the
    sharded attribution run reproduce the unsharded one bit for bit)."""
    size, seed = 64, 50
    g = build_generator(size, seed)
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    w = fx.seeded((4, 512), seed + 2).to(DEV)
    ct = fx.seeded((4, 3, size, size), seed + 3).to(DEV)

    def run(idx):
        wi = w[idx].clone().requires_grad_(True)
        img, _ = g([wi], input_is_latent=True, noise=noise)
        (gw,) = torch.autograd.grad((img * ct[idx]).sum(), wi)
        return img.detach(), gw

    img_all, g_all = run(slice(0, 4))
    for i in range(4):
        img_i, g_i = run(slice(i, i + 1))
        assert torch.equal(img_i[0], img_all[i])
        assert torch.equal(g_i[0], g_all[i])


def test_tf32_batch_independence_is_bitwise_and_default_follows_cudnn_flag():
    """Tensor-core path: tiles are reduced per (sample, tile) in a fixed order, so a trajectory is bit-identical
    whatever shares its batch.  Also: with ``precision=None`` the Generator follows torch.backends.cudnn.allow_tf32,
    the switch that decides the reference's conv arithmetic."""
    from lfp_native import capi
    size, seed = 128, 52
    g = build_generator(size, seed)
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    w = fx.seeded((3, 512), seed + 2).to(DEV)
    ct = fx.seeded((3, 3, size, size), seed + 3).to(DEV)

    def run(idx):
        wi = w[idx].clone().requires_grad_(True)
        img, _ = g([wi], input_is_latent=True, noise=noise)
        (gw,) = torch.autograd.grad((img * ct[idx]).sum(), wi)
        return img.detach(), gw

    g.precision = capi.PREC_TF32
    img_all, g_all = run(slice(0, 3))
    for i in range(3):
        img_i, g_i = run(slice(i, i + 1))
        assert torch.equal(img_i[0], img_all[i])
        assert torch.equal(g_i[0], g_all[i])
    g.precision = None
    old = torch.backends.cudnn.allow_tf32
    try:
        torch.backends.cudnn.allow_tf32 = True
        img_tf, _ = run(slice(0, 1))
        torch.backends.cudnn.allow_tf32 = False
        img_fp, _ = run(slice(0, 1))
    finally:
        torch.backends.cudnn.allow_tf32 = old
    assert torch.equal(img_tf[0], img_all[0])
    assert not torch.equal(img_fp[0], img_all[0])
    img_close(img_tf.cpu().numpy(), img_fp.cpu().numpy(), 5e-3)


@pytest.mark.parametrize("size", [512, 1024])
@pytest.mark.parametrize("prec", ["fp32", "tf32"])
def test_full_size_oracle_parity(size, prec):
    """The benchmarked sizes (BASELINE.json: 1024 px attribution, 512 px config 4) against the CPU oracle on both
    arithmetic paths: image, every saved StyledConv activation (named per layer), leaky-ReLU branch pattern, and
    the per-slot latent gradient with the CUDA path's branches imposed on the oracle.  These are the only
    tests that reach the resident-weight C <= 64 tcgen05 instantiations, the NHWC ring FIR kernels, the ring
    act_bwd kernel and the C = 64 / 32 ToRGB kernels (src/model.py:499-572 at full size).
    Tolerances as stated in DESIGN.md section 2: fp32 1e-4 image / 5e-5 gradient; tf32 5e-3 / 1e-2."""
    from lfp_native import capi
    seed = 300 + size
    noise = fx.make_noise(size, seed + 1)
    lat = fx.seeded((1, oracle.n_latent(size), 512), seed + 2)
    if prec == "fp32":
        rel, flips = grad_parity(size, seed, lat, noise, cm=2, precision=capi.PREC_FP32)
    else:
        rel, flips = grad_parity(size, seed, lat, noise, cm=2, precision=capi.PREC_TF32, img_tol=5e-3, grad_tol=1e-2,
                                 flip_band=2e-2)
    print(f"{prec} size {size}: gradient rel err {rel:.2e}, {flips} branch flips")


def test_full_size_1024_batched_equals_single_bitwise():
    """B = 20 (the bench's batch) picks different slice widths / work distributions than B = 1 at the low
    resolutions; the per-trajectory result must not depend on it (tensor-core path, 1024 px)."""
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    size, seed, B = 1024, 77, 3
    params = fx.make_params(size, seed)
    plan = SynthesisPlan(size, device=DEV)
    plan.load(params)
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    lat = fx.seeded((B, plan.n_latent, 512), seed + 2).to(DEV)
    ct = fx.seeded((B, 3, size, size), seed + 3).to(DEV)
    ws = plan.new_workspace(B)
    img = plan.forward(lat, noise, ws, capi.PREC_TF32).clone()
    dl = plan.backward(ct, B, ws, capi.PREC_TF32).clone()
    del ws
    ws1 = plan.new_workspace(1)
    for b in (0, B - 1):
        img1 = plan.forward(lat[b:b + 1].contiguous(), noise, ws1, capi.PREC_TF32)
        assert torch.equal(img1[0], img[b])
        dl1 = plan.backward(ct[b:b + 1].contiguous(), 1, ws1, capi.PREC_TF32)
        assert torch.equal(dl1[0], dl[b])
    assert torch.isfinite(dl).all()


def test_interleaved_forwards_keep_backward_correct():
    size, seed = 16, 60
    g = build_generator(size, seed)
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    w1 = fx.seeded((2, 512), 61).to(DEV).requires_grad_(True)
    w2 = fx.seeded((2, 512), 62).to(DEV).requires_grad_(True)
    img1, _ = g([w1], input_is_latent=True, noise=noise)
    img2, _ = g([w2], input_is_latent=True, noise=noise)   # second forward before the first backward
    (g1,) = torch.autograd.grad(img1.square().sum(), w1)
    (g2,) = torch.autograd.grad(img2.square().sum(), w2)
    w1b = w1.detach().clone().requires_grad_(True)
    img1b, _ = g([w1b], input_is_latent=True, noise=noise)
    (g1b,) = torch.autograd.grad(img1b.square().sum(), w1b)
    assert torch.equal(g1, g1b)
    assert not torch.equal(g1, g2)


def test_host_buffer_entry_point_matches():
    """lfp_synth_forward_backward_host: HOST pointers in and out (INTEGRATION.md binding)."""
    import ctypes as C
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    size, seed, B = 16, 70, 2
    params = fx.make_params(size, seed)
    plan = SynthesisPlan(size, device=DEV)
    plan.load(params)
    noise = fx.make_noise(size, seed + 1)
    lat = fx.seeded((B, plan.n_latent, 512), seed + 2).contiguous()
    ct = fx.seeded((B, 3, size, size), seed + 3).contiguous()
    img = torch.empty(B, 3, size, size)
    dlat = torch.empty_like(lat)
    nptr = (C.c_void_p * plan.num_noise)(*[n.data_ptr() for n in noise])
    nb = (C.c_int * plan.num_noise)(*[1] * plan.num_noise)
    capi.check(capi.lib().lfp_synth_forward_backward_host(plan._h, B, lat.data_ptr(), nptr, nb, img.data_ptr(),
                                                          ct.data_ptr(), dlat.data_ptr(), capi.PREC_FP32))
    lr = lat.clone().requires_grad_(True)
    ref = oracle.synthesis(params, lr, noise)
    (gref,) = torch.autograd.grad((ref * ct).sum(), lr)
    img_close(img.numpy(), ref.detach().numpy(), 1e-4)
    assert np.linalg.norm(dlat.numpy() - gref.numpy()) <= KINK_TOL * np.linalg.norm(gref.numpy())


def test_forward_only_workspace_matches_and_refuses_backward():
    """lfp_synth_generate (ping-pong activations, nothing kept) gives the same bits as lfp_synth_forward, its workspace is
    a fraction of the full one, and a backward on it is refused (LFP_ESTATE) instead of reading stale memory."""
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    size, seed, B = 128, 81, 3
    plan = SynthesisPlan(size, device=DEV)
    plan.load(fx.make_params(size, seed))
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    lat = fx.seeded((B, plan.n_latent, 512), seed + 2).to(DEV)
    for prec in (capi.PREC_FP32, capi.PREC_TF32):
        ws = plan.new_workspace(B)
        full = plan.forward(lat, noise, ws, prec).clone()
        wsg = plan.new_workspace(B, forward_only=True)
        gen = plan.forward(lat, noise, wsg, prec, forward_only=True)
        assert torch.equal(full, gen)
        assert plan.workspace_bytes(B, forward_only=True) < plan.workspace_bytes(B)
        with pytest.raises(capi.LfpError, match="no matching forward|workspace too small"):
            plan.backward(torch.zeros_like(gen), B, wsg, prec)
        plan.forward(lat, noise, ws, prec, forward_only=True)   # a forward-only pass on a FULL-size workspace
        with pytest.raises(capi.LfpError, match="no matching forward"):
            plan.backward(torch.zeros_like(gen), B, ws, prec)


def test_two_workspaces_interleaved_through_the_c_abi():
    """Forward on workspace A, forward on workspace B (different noise), then both backwards: each backward uses the
    noise and batch recorded for ITS workspace (the plan kept one global record)."""
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    size, seed = 32, 83
    plan = SynthesisPlan(size, device=DEV)
    plan.load(fx.make_params(size, seed))
    nA = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    nB = [n.to(DEV) for n in fx.make_noise(size, seed + 7)]
    lat = fx.seeded((2, plan.n_latent, 512), seed + 2).to(DEV)
    ct = fx.seeded((2, 3, size, size), seed + 3).to(DEV)

    def alone(noise):
        ws = plan.new_workspace(2)
        plan.forward(lat, noise, ws, capi.PREC_FP32)
        return plan.backward(ct, 2, ws, capi.PREC_FP32).clone()

    gA, gB = alone(nA), alone(nB)
    assert not torch.equal(gA, gB)
    wsA, wsB = plan.new_workspace(2), plan.new_workspace(2)
    plan.forward(lat, nA, wsA, capi.PREC_FP32)
    plan.forward(lat, nB, wsB, capi.PREC_FP32)
    assert torch.equal(plan.backward(ct, 2, wsA, capi.PREC_FP32), gA)
    assert torch.equal(plan.backward(ct, 2, wsB, capi.PREC_FP32), gB)
    # a backward with another batch size or arithmetic than the forward on that workspace is refused
    with pytest.raises(capi.LfpError, match="no matching forward"):
        plan.backward(ct, 2, wsA, capi.PREC_TF32)


def test_generation_batch64_at_1024px_forward_only():
    """BASELINE.json configs[1]: fingerprinted generation, batch 64 at 1024 px, one call (a [64,32,1024,1024] activation is
    exactly 2^31 elements: the reference's int-indexed ops cannot run it un-chunked).  Samples of the
    batched call equal single-sample calls bit for bit, and those are oracle-checked by test_full_size_oracle_parity."""
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    size, seed, B = 1024, 1324, 64   # the seed of test_full_size_oracle_parity[*-1024]
    params = fx.make_params(size, seed)
    plan = SynthesisPlan(size, device=DEV)
    plan.load(params)
    noise = [n.to(DEV) for n in fx.make_noise(size, seed + 1)]
    lat1 = fx.seeded((1, plan.n_latent, 512), seed + 2)
    lat = torch.cat([lat1, fx.seeded((B - 1, plan.n_latent, 512), seed + 9)]).to(DEV)
    img = plan.generate(lat, noise, capi.PREC_TF32).clone()
    assert torch.isfinite(img).all()
    for b in (0, 31, 63):
        one = plan.generate(lat[b:b + 1].contiguous(), noise, capi.PREC_TF32)
        assert torch.equal(one[0], img[b]), b
    # sample 0 is the latent of the oracle-checked full-size test: compare with the oracle directly as well
    with torch.no_grad():
        ref = oracle.synthesis(params, lat1, fx.make_noise(size, seed + 1))
    img_close(img[:1].cpu().numpy(), ref.numpy(), 5e-3)
    print(f"generation B=64 @1024: forward-only workspace {plan.workspace_bytes(B, forward_only=True) / 1e9:.1f} GB "
          f"(full: {plan.workspace_bytes(B) / 1e9:.1f} GB)")


def test_host_entry_point_in_a_loop_does_not_allocate():
    """lfp_synth_forward_backward_host keeps its staging buffers: device memory in use is flat across repeated calls."""
    import ctypes as C
    from lfp_native import capi
    from lfp_native.synthesis import SynthesisPlan
    size, seed, B = 32, 70, 2
    plan = SynthesisPlan(size, device=DEV)
    plan.load(fx.make_params(size, seed))
    noise = fx.make_noise(size, seed + 1)
    lat = fx.seeded((B, plan.n_latent, 512), seed + 2).contiguous()
    ct = fx.seeded((B, 3, size, size), seed + 3).contiguous()
    img, dlat = torch.empty(B, 3, size, size), torch.empty_like(lat)
    nptr = (C.c_void_p * plan.num_noise)(*[n.data_ptr() for n in noise])
    nb = (C.c_int * plan.num_noise)(*[1] * plan.num_noise)

    def call():
        capi.check(capi.lib().lfp_synth_forward_backward_host(plan._h, B, lat.data_ptr(), nptr, nb, img.data_ptr(),
                                                              ct.data_ptr(), dlat.data_ptr(), capi.PREC_FP32))
        return torch.cuda.mem_get_info()[0]

    call()
    first = dlat.clone()
    free0 = call()
    for _ in range(5):
        free = call()
    assert free == free0
    assert torch.equal(first, dlat)
