"""GPU parity of the stand-alone native modulated convolution (lfp_modconv_*, include/lfp_sg2.h group 5, reached through
model.ModulatedConv2d / StyledConv / ToRGB) against the reference-generated goldens (tests/golden/modconv.npz:
ModulatedConv2d fused and unfused, plain / up / 1x1-no-demod, out, dX, dstyle) and against the CPU oracle at every
(Cin, Cout, H') of the generator on both arithmetic paths.

Tolerances: fp32 path rtol 1e-4 (+ 2e-5 of the tensor's scale, as the oracle-vs-golden test); tf32 path 5e-3 of the
tensor's scale for the output, 1e-2 relative L2 for the gradients (the stated tf32 bar of DESIGN.md section 2)."""
import numpy as np
import pytest
import torch

import fixtures as fx
import oracle
from golden.make_golden import MODCONV_CASES, modconv_inputs

pytestmark = pytest.mark.gpu
DEV = "cuda"


def make_module(Cin, Cout, k, sd, demod, up, t, precision):
    from model import ModulatedConv2d
    m = ModulatedConv2d(Cin, Cout, k, sd, demodulate=demod, upsample=up)
    with torch.no_grad():
        m.weight.copy_(t["weight"])
        m.modulation.weight.copy_(t["mod_w"])
        m.modulation.bias.copy_(t["mod_b"])
    m = m.to(DEV)
    m.precision = precision
    return m


@pytest.mark.parametrize("i", range(len(MODCONV_CASES)), ids=[c[0] for c in MODCONV_CASES])
def test_modconv_matches_reference_golden(golden, i):
    from lfp_native import capi
    name, B, Cin, Cout, H, W, k, demod, up, sd = MODCONV_CASES[i]
    t = modconv_inputs(i, B, Cin, Cout, H, W, k, sd)
    m = make_module(Cin, Cout, k, sd, demod, up, t, capi.PREC_FP32)
    before = capi.launch_count()
    x = t["x"].to(DEV).requires_grad_(True)
    s = t["style"].to(DEV).requires_grad_(True)
    y = m(x, s)
    ct = fx.seeded(tuple(y.shape), 500 + i).to(DEV)
    gx, gs = torch.autograd.grad((y * ct).sum(), [x, s])
    assert capi.launch_count() > before          # the native path ran (no cuDNN composite)
    for got, key in ((y.detach(), "y"), (gx, "gx"), (gs, "gs")):
        for algebra in ("fused", "unfused"):
            ref = golden[f"modconv/{name}/{algebra}/{key}"]
            scale = np.abs(ref).max()
            np.testing.assert_allclose(got.cpu().numpy(), ref, rtol=1e-3 if algebra == "fused" else 1e-4,
                                       atol=(1e-4 if algebra == "fused" else 2e-5) * scale, err_msg=f"{name}/{algebra}/{key}")


# every (Cin, Cout, H_out, up) of Generator(1024, 512, 8): src/model.py:418-473, SURVEY.md 9a
LAYER_SHAPES = ([(512, 512, 4, False)] + [(512, 512, r, u) for r in (8, 16, 32, 64) for u in (True, False)] +
                [(512, 256, 128, True), (256, 256, 128, False), (256, 128, 256, True), (128, 128, 256, False),
                 (128, 64, 512, True), (64, 64, 512, False), (64, 32, 1024, True), (32, 32, 1024, False)])


def layer_check(Cin, Cout, res, up, B, k=3, demod=True, seed=0, precs=("fp32", "tf32")):
    """One layer against the oracle (computed once) on both arithmetic paths."""
    from lfp_native import capi
    H = res // 2 if up else res
    sd = 512
    rs = np.random.RandomState(1000 + seed + Cin + 7 * Cout + res)
    t = dict(weight=torch.from_numpy(rs.standard_normal((1, Cout, Cin, k, k)).astype(np.float32)),
             mod_w=torch.from_numpy(rs.standard_normal((Cin, sd)).astype(np.float32)),
             mod_b=torch.ones(Cin) + 0.1 * torch.from_numpy(rs.standard_normal(Cin).astype(np.float32)))
    x = torch.from_numpy(rs.standard_normal((B, Cin, H, H)).astype(np.float32))
    style = torch.from_numpy(rs.standard_normal((B, sd)).astype(np.float32))
    xr, sr = x.clone().requires_grad_(True), style.clone().requires_grad_(True)
    yr = oracle.modulated_conv2d(xr, sr, t["weight"], t["mod_w"], t["mod_b"], demodulate=demod, upsample=up)
    ct = torch.from_numpy(rs.standard_normal(tuple(yr.shape)).astype(np.float32))
    gxr, gsr = torch.autograd.grad((yr * ct).sum(), [xr, sr])
    out = {}
    for prec_name in precs:
        m = make_module(Cin, Cout, k, sd, demod, up, t, capi.PREC_FP32 if prec_name == "fp32" else capi.PREC_TF32)
        xg, sg = x.to(DEV).requires_grad_(True), style.to(DEV).requires_grad_(True)
        y = m(xg, sg)
        gx, gs = torch.autograd.grad((y * ct.to(DEV)).sum(), [xg, sg])
        tol_y, tol_g = (1e-4, 1e-4) if prec_name == "fp32" else (5e-3, 1e-2)
        err = float((y.detach().cpu() - yr.detach()).abs().max()) / max(1.0, float(yr.detach().abs().max()))
        rel_x = float((gx.cpu() - gxr).norm() / gxr.norm())
        rel_s = float((gs.cpu() - gsr).norm() / gsr.norm())
        assert err <= tol_y and rel_x <= tol_g and rel_s <= tol_g, (Cin, Cout, res, up, B, prec_name, err, rel_x, rel_s)
        out[prec_name] = (err, rel_x, rel_s)
    return out


@pytest.mark.parametrize("shape", LAYER_SHAPES, ids=[f"{a}to{b}at{r}{'up' if u else ''}" for a, b, r, u in LAYER_SHAPES])
def test_modconv_every_generator_shape_against_oracle(shape):
    """out, dX and dstyle at every (Cin, Cout, H') of Generator(1024), fp32 CUDA-core and tcgen05 tf32
    paths, B in {1, 2} everywhere and B = 20 (the attribution batch) on the low-resolution layers, where the tensor-core
    kernel picks narrower channel slices for small batches."""
    Cin, Cout, res, up = shape
    for B in ((1, 2, 20) if res <= 16 else ((1, 2) if res <= 256 else (1,))):
        layer_check(Cin, Cout, res, up, B)


def test_torgb_1x1_no_demod_and_odd_channel_counts():
    """ToRGB's conv (1x1, Cout = 3, no demodulation: src/model.py:376) at the generator's channel counts, and a layer
    whose channel counts are not multiples of anything (padded inside the native op)."""
    for Cin, res in ((512, 4), (256, 128), (32, 256)):
        layer_check(Cin, 3, res, False, 2, k=1, demod=False)
    layer_check(24, 10, 16, False, 3)
    layer_check(10, 24, 16, True, 2)


def test_styledconv_and_torgb_modules_run_native():
    """StyledConv (conv + noise + fused bias-act) and ToRGB (1x1 conv + bias + upsampled skip) used stand-alone,
    src/model.py:332-388, against the oracle's styled_conv / to_rgb."""
    from model import StyledConv, ToRGB
    from lfp_native import capi
    size, seed = 16, 5
    params = fx.make_params(size, seed)
    sc = StyledConv(512, 512, 3, 512, upsample=True)
    rgb = ToRGB(512, 512)
    sc.load_state_dict({k[len("convs.0."):]: v for k, v in params.items() if k.startswith("convs.0.")}, strict=False)
    rgb.load_state_dict({k[len("to_rgbs.0."):]: v for k, v in params.items() if k.startswith("to_rgbs.0.")}, strict=False)
    sc, rgb = sc.to(DEV), rgb.to(DEV)
    sc.conv.precision = rgb.conv.precision = capi.PREC_FP32
    x = fx.seeded((2, 512, 4, 4), 6)
    st = fx.seeded((2, 512), 7)
    nz = fx.seeded((1, 1, 8, 8), 8)
    skip = fx.seeded((2, 3, 4, 4), 9)
    before = capi.launch_count()
    y = sc(x.to(DEV), st.to(DEV), noise=nz.to(DEV))
    img = rgb(y, st.to(DEV), skip.to(DEV))
    assert capi.launch_count() - before >= 8
    yr = oracle.styled_conv(params, "convs.0", x, st, nz, upsample=True)
    imgr = oracle.to_rgb(params, "to_rgbs.0", yr, st, skip)
    np.testing.assert_allclose(y.detach().cpu().numpy(), yr.numpy(), rtol=1e-4, atol=1e-4 * float(yr.abs().max()))
    np.testing.assert_allclose(img.detach().cpu().numpy(), imgr.numpy(), rtol=1e-4, atol=1e-4 * float(imgr.abs().max()))


def test_tall_work_items_match_oracle_and_single_sample_bits():
    """128 -> 128 at 256 px with B = 4 runs on tall work items (two 16-row halves per item share the streamed weight slices,
    conv_tc.cu Args::nhalf; chosen from 3 samples up at this shape), B = 1 does not: the oracle check covers the tall path on
    both forward and data-gradient kernels, and each sample of the batch has the bits of its single-sample call."""
    from lfp_native import capi
    Cin = Cout = 128
    res, B, sd, k = 256, 4, 512, 3
    layer_check(Cin, Cout, res, False, B, seed=3, precs=("tf32",))
    rs = np.random.RandomState(77)
    t = dict(weight=torch.from_numpy(rs.standard_normal((1, Cout, Cin, k, k)).astype(np.float32)),
             mod_w=torch.from_numpy(rs.standard_normal((Cin, sd)).astype(np.float32)),
             mod_b=torch.ones(Cin))
    x = torch.from_numpy(rs.standard_normal((B, Cin, res, res)).astype(np.float32)).to(DEV)
    style = torch.from_numpy(rs.standard_normal((B, sd)).astype(np.float32)).to(DEV)
    ct = torch.from_numpy(rs.standard_normal((B, Cout, res, res)).astype(np.float32)).to(DEV)
    m = make_module(Cin, Cout, k, sd, True, False, t, capi.PREC_TF32)

    def run(sl):
        xg, sg = x[sl].clone().requires_grad_(True), style[sl].clone().requires_grad_(True)
        y = m(xg, sg)
        gx, gs = torch.autograd.grad((y * ct[sl]).sum(), [xg, sg])
        return y.detach(), gx, gs

    yb, gxb, gsb = run(slice(0, B))
    for b in (0, B - 1):
        y1, gx1, gs1 = run(slice(b, b + 1))
        assert torch.equal(y1[0], yb[b]) and torch.equal(gx1[0], gxb[b]) and torch.equal(gs1[0], gsb[b]), b
