"""GPU parity of the two legacy ops (through the drop-in ``op`` package -> C ABI) against the
reference-generated golden vectors and the CPU oracle."""
import numpy as np
import pytest
import torch

import fixtures as fx
import oracle
from golden.make_golden import LRELU_CASES, UPFIRDN_CASES, kernel_of

pytestmark = pytest.mark.gpu
DEV = "cuda"


@pytest.mark.parametrize("i", range(len(UPFIRDN_CASES)))
def test_upfirdn2d_golden(golden, i):
    import op
    name, shape, kspec, up, down, pad = UPFIRDN_CASES[i]
    x = fx.seeded(shape, 100 + i).to(DEV).requires_grad_(True)
    k = kernel_of(kspec).to(DEV)
    y = op.upfirdn2d(x, k, up=up, down=down, pad=pad)
    ref = golden[f"upfirdn/{name}/y"]
    assert tuple(y.shape) == ref.shape
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref, rtol=1e-5, atol=1e-6)
    ct = fx.seeded(tuple(y.shape), 200 + i).to(DEV)
    (gx,) = torch.autograd.grad((y * ct).sum(), x)
    np.testing.assert_allclose(gx.cpu().numpy(), golden[f"upfirdn/{name}/gx"], rtol=1e-5, atol=1e-6)


HOT = [  # the configurations the synthesis path hits, at sizes that use the tiled kernel
    ("blur", (2, 5, 129, 129), 1, 1, (1, 1)),
    ("blur_bwd", (2, 5, 128, 128), 1, 1, (2, 2)),
    ("up2", (2, 3, 70, 66), 2, 1, (2, 1)),
    ("down2", (2, 3, 140, 132), 1, 2, (1, 1)),
    ("blur_wide", (1, 2, 40, 300), 1, 1, (1, 1)),
    ("tiny4", (3, 4, 4, 4), 2, 1, (2, 1)),
    # the bulk-copy row-ring kernel (wide 1:1 maps whose element count is a multiple of 4): several column segments and row
    # bands, more row groups than ring slots, every row skew (in_w % 4), odd plane counts, crop pads
    ("ring_odd_planes", (3, 1, 70, 300), 1, 1, (1, 1)),
    ("ring_w517", (2, 2, 130, 517), 1, 1, (2, 2)),
    ("ring_w258", (1, 2, 66, 258), 1, 1, (1, 1)),
    ("ring_w259", (2, 2, 65, 259), 1, 1, (2, 1)),
    ("ring_one_plane", (1, 1, 64, 1024), 1, 1, (1, 1)),
    ("ring_crop", (1, 4, 96, 200), 1, 1, (-1, 4)),
    ("ring_narrow", (4, 8, 17, 17), 1, 1, (1, 1)),     # 2-warp segments
    ("ring_mid", (2, 4, 65, 100), 1, 1, (2, 2)),       # 4-warp segments, one short band
    # the small-plane kernel (>= 32 planes of at most 34 x 34 samples, pads 0..3; runs of planes staged zero-haloed in shared
    # memory): every up / down mode, odd plane sizes, every pad, a last run shorter than the others; the crops and the larger
    # planes of this block fall through to the ring / streaming kernels
    ("planes_blur17", (8, 8, 17, 17), 1, 1, (1, 1)),
    ("planes_bwd16", (8, 8, 16, 16), 1, 1, (2, 2)),
    ("planes_up8", (4, 16, 8, 8), 2, 1, (2, 1)),
    ("planes_down16", (4, 16, 16, 16), 1, 2, (1, 1)),
    ("planes_blur33", (2, 16, 33, 33), 1, 1, (1, 1)),
    ("planes_down32", (33, 1, 32, 32), 1, 2, (1, 1)),
    ("planes_up16", (33, 1, 16, 16), 2, 1, (2, 1)),
    ("planes_up_odd", (3, 11, 7, 5), 2, 1, (2, 1)),
    ("planes_down_odd", (3, 11, 11, 9), 1, 2, (1, 1)),
    ("planes_up_pad3", (33, 1, 7, 9), 2, 1, (3, 3)),
    ("planes_down_pad0", (33, 1, 7, 9), 1, 2, (0, 0)),
    ("planes_pad3", (33, 1, 7, 9), 1, 1, (3, 3)),
    ("planes_up_pad4", (33, 1, 5, 6), 2, 1, (1, 2, 3, 0)),
    ("planes_many", (700, 1, 5, 3), 2, 1, (0, 3, 1, 2)),
    ("planes_one_px", (33, 1, 1, 1), 2, 1, (2, 1)),
    ("fallback_blur65", (2, 20, 65, 65), 1, 1, (1, 1)),
    ("fallback_up33", (2, 16, 33, 33), 2, 1, (2, 1)),
    ("fallback_crop", (5, 9, 9, 13), 1, 1, (-1, 4)),
    ("fallback_pad4", (37, 1, 10, 12), 1, 2, (0, 2, 3, -1)),
]


@pytest.mark.parametrize("case", HOT, ids=[c[0] for c in HOT])
@pytest.mark.parametrize("kspec", ["fir4x4_g4", "asym4x4", "asym3x3", "asym2x2"])
def test_upfirdn2d_tiled_vs_oracle(case, kspec):
    import op
    name, shape, up, down, pad = case
    x = fx.seeded(shape, 7)
    k = kernel_of(kspec)
    ref = oracle.upfirdn2d(x, k, up=up, down=down, pad=pad)
    xg = x.to(DEV).requires_grad_(True)
    y = op.upfirdn2d(xg, k.to(DEV), up=up, down=down, pad=pad)
    assert y.shape == ref.shape
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref.numpy(), rtol=1e-5, atol=2e-6)
    ct = fx.seeded(tuple(ref.shape), 8)
    xr = x.clone().requires_grad_(True)
    (gref,) = torch.autograd.grad((oracle.upfirdn2d(xr, k, up=up, down=down, pad=pad) * ct).sum(), xr)
    (g,) = torch.autograd.grad((y * ct.to(DEV)).sum(), xg)
    np.testing.assert_allclose(g.cpu().numpy(), gref.numpy(), rtol=1e-5, atol=2e-6)


def test_upfirdn2d_many_planes_and_dtypes():
    import op
    k = kernel_of("fir4x4_g4")
    x = fx.seeded((70000, 1, 6, 6), 9)  # N*C > 65535: beyond one grid dimension
    y = op.upfirdn2d(x.to(DEV), k.to(DEV), pad=(1, 1))
    np.testing.assert_allclose(y.cpu().numpy(), oracle.upfirdn2d(x, k, pad=(1, 1)).numpy(), rtol=1e-5, atol=1e-6)
    # plane counts beyond one grid dimension in the specialised kernels: the ring kernel launches in chunks of 65534 plane
    # pairs, the streaming kernels loop over planes; checked on the planes around the boundaries
    big = torch.randn(140000, 1, 17, 17, generator=torch.Generator().manual_seed(5))
    yb = op.upfirdn2d(big.to(DEV), k.to(DEV), pad=(1, 1)).cpu()
    for lo in (0, 65530, 131060, 139990):
        ref = oracle.upfirdn2d(big[lo:lo + 10], k, pad=(1, 1))
        np.testing.assert_allclose(yb[lo:lo + 10].numpy(), ref.numpy(), rtol=1e-5, atol=1e-6)
    small = torch.randn(70000, 1, 8, 8, generator=torch.Generator().manual_seed(6))
    yu = op.upfirdn2d(small.to(DEV), k.to(DEV), up=2, pad=(2, 1)).cpu()
    yd = op.upfirdn2d(small.to(DEV), k.to(DEV), down=2, pad=(1, 1)).cpu()
    for lo in (0, 65530, 69990):
        np.testing.assert_allclose(yu[lo:lo + 10].numpy(), oracle.upfirdn2d(small[lo:lo + 10], k, up=2, pad=(2, 1)).numpy(), rtol=1e-5, atol=1e-6)
        np.testing.assert_allclose(yd[lo:lo + 10].numpy(), oracle.upfirdn2d(small[lo:lo + 10], k, down=2, pad=(1, 1)).numpy(), rtol=1e-5, atol=1e-6)
    x = fx.seeded((2, 3, 33, 31), 10)
    for dt, tol in ((torch.float64, 1e-12), (torch.float16, 2e-3)):
        ref = oracle.upfirdn2d(x.double(), k.double(), up=2, pad=(2, 1))
        y = op.upfirdn2d(x.to(DEV, dt), k.to(DEV, dt), up=2, pad=(2, 1))
        assert y.dtype == dt
        np.testing.assert_allclose(y.double().cpu().numpy(), ref.numpy(), rtol=tol, atol=tol)


def test_upfirdn2d_double_backward():
    import op
    k = kernel_of("asym4x4").to(DEV, torch.float64)
    x = fx.seeded((1, 2, 6, 5), 11).to(DEV, torch.float64).requires_grad_(True)
    assert torch.autograd.gradcheck(lambda t: op.upfirdn2d(t, k, up=2, pad=(2, 1)), (x,), atol=1e-6)
    assert torch.autograd.gradgradcheck(lambda t: op.upfirdn2d(t, k, down=2, pad=(1, 1)), (x,), atol=1e-6)


@pytest.mark.parametrize("i", range(len(LRELU_CASES)))
def test_fused_leaky_relu_golden(golden, i):
    import op
    name, shape, use_bias = LRELU_CASES[i]
    x = fx.seeded(shape, 300 + i)
    x.view(-1)[::5] = 0.0
    b = fx.seeded((shape[1],), 320 + i) if use_bias else None
    if b is not None:
        b[0] = 0.0
    xg = x.to(DEV).requires_grad_(True)
    bg = b.to(DEV).requires_grad_(True) if b is not None else None
    y = op.fused_leaky_relu(xg, bg)
    np.testing.assert_array_equal(y.detach().cpu().numpy(), golden[f"lrelu/{name}/y"])
    ct = fx.seeded(tuple(y.shape), 340 + i).to(DEV)
    grads = torch.autograd.grad((y * ct).sum(), [xg] + ([bg] if bg is not None else []))
    # the CPU golden multiplies (g*scale)*slope, the native op (g*slope)*scale as the reference's CUDA kernel
    # does (src/op/fused_bias_act_kernel.cu:55-62): 1 ulp apart
    np.testing.assert_allclose(grads[0].cpu().numpy(), golden[f"lrelu/{name}/gx"], rtol=3e-7, atol=0)
    if b is not None:
        np.testing.assert_allclose(grads[1].cpu().numpy(), golden[f"lrelu/{name}/gb"], rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("shape", [(2, 32, 64, 64), (3, 512), (2, 6, 7, 9), (1, 16, 130, 2)])
@pytest.mark.parametrize("slope,scale", [(0.2, 2 ** 0.5), (0.1, 1.5)])
def test_fused_leaky_relu_vs_oracle(shape, slope, scale):
    """Vector and scalar kernels, non-default slope/scale (honoured on GPU)."""
    import op
    x = fx.seeded(shape, 12)
    b = fx.seeded((shape[1],), 13)
    xr, br = x.clone().requires_grad_(True), b.clone().requires_grad_(True)
    ref = oracle.fused_leaky_relu(xr, br, slope, scale)
    ct = fx.seeded(shape, 14)
    gxr, gbr = torch.autograd.grad((ref * ct).sum(), [xr, br])
    xg, bg = x.to(DEV).requires_grad_(True), b.to(DEV).requires_grad_(True)
    y = op.fused_leaky_relu(xg, bg, slope, scale)
    np.testing.assert_allclose(y.detach().cpu().numpy(), ref.detach().numpy(), rtol=1e-6, atol=1e-7)
    gx, gb = torch.autograd.grad((y * ct.to(DEV)).sum(), [xg, bg])
    np.testing.assert_allclose(gx.cpu().numpy(), gxr.numpy(), rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(gb.cpu().numpy(), gbr.numpy(), rtol=1e-4, atol=1e-4)
    m = op.FusedLeakyReLU(shape[1]).to(DEV)
    assert list(dict(m.named_parameters())) == ["bias"]
    y0 = m(x.to(DEV))
    np.testing.assert_allclose(y0.detach().cpu().numpy(), oracle.fused_leaky_relu(x, torch.zeros(shape[1])).numpy(),
                               rtol=1e-6, atol=1e-7)


def test_fused_leaky_relu_other_dtypes_and_empty():
    import op
    x = fx.seeded((2, 8, 16, 16), 15)
    b = fx.seeded((8,), 16)
    ref = oracle.fused_leaky_relu(x.double(), b.double())
    y = op.fused_leaky_relu(x.to(DEV, torch.float64), b.to(DEV, torch.float64))
    np.testing.assert_allclose(y.cpu().numpy(), ref.numpy(), rtol=1e-7, atol=1e-7)  # alpha/scale pass through float
    yh = op.fused_leaky_relu(x.to(DEV, torch.float16), b.to(DEV, torch.float16))
    np.testing.assert_allclose(yh.float().cpu().numpy(), ref.numpy(), rtol=5e-3, atol=5e-3)
    e = op.fused_leaky_relu(torch.empty(0, 4, device=DEV), torch.zeros(4, device=DEV))
    assert e.shape == (0, 4)


def test_host_buffer_entry_points():
    """The *_host variants take HOST pointers (what a non-torch caller binds)."""
    import ctypes as C
    from lfp_native import capi
    L = capi.lib()
    x = fx.seeded((2, 3, 17, 17), 17).contiguous()
    k = kernel_of("fir4x4_g4").contiguous()
    out = torch.empty(2, 3, 16, 16)
    capi.check(L.lfp_upfirdn2d_host(x.data_ptr(), k.data_ptr(), out.data_ptr(), capi.F32, 6, 17, 17, 1, 4, 4,
                                    1, 1, 1, 1, 1, 1, 1, 1))
    np.testing.assert_allclose(out.numpy(), oracle.upfirdn2d(x, k, pad=(1, 1)).numpy(), rtol=1e-5, atol=1e-6)
    b = fx.seeded((3,), 18)
    o2 = torch.empty_like(x)
    capi.check(L.lfp_fused_bias_act_host(x.data_ptr(), b.data_ptr(), None, o2.data_ptr(), capi.F32, x.numel(),
                                         17 * 17, 3, 3, 0, 0.2, 2 ** 0.5))
    np.testing.assert_allclose(o2.numpy(), oracle.fused_leaky_relu(x, b).numpy(), rtol=1e-6, atol=1e-7)


def test_more_than_2_pow_31_elements():
    """64-bit indexing (the reference's kernels overflow `int` here): 2^31 + 2^20 fp16 elements through
    fused_leaky_relu, and an upfirdn2d whose input + output exceed 2^31 elements; checked slice-wise against torch."""
    from op import fused_leaky_relu, upfirdn2d
    n_c, h, w = 4, 16384, 32784   # 4 * 16384 * 32784 = 2_148_532_224 > 2^31
    x = torch.empty((1, n_c, h, w), dtype=torch.float16, device=DEV)
    x.view(-1)[: 1 << 20].normal_()
    x.view(-1)[1 << 20:] = x.view(-1)[: 1 << 20].repeat((x.numel() >> 20) + 1)[: x.numel() - (1 << 20)]
    b = torch.tensor([0.5, -0.25, 0.125, 1.0], dtype=torch.float16, device=DEV)
    y = fused_leaky_relu(x, b)
    for c in (0, 3):
        for rows in (slice(0, 64), slice(h - 64, h)):
            ref = torch.nn.functional.leaky_relu(x[0, c, rows].float() + b[c].float(), 0.2) * 2 ** 0.5
            np.testing.assert_allclose(y[0, c, rows].float().cpu().numpy(), ref.cpu().numpy(), rtol=2e-3, atol=2e-3)
    del y
    k = torch.tensor([[1., 3., 3., 1.]], device=DEV)
    k = (k.t() @ k) / 64 * 4
    xf = x[:, :2].float()     # 2 * 16384 * 32784 fp32 = 1.07e9 in, same out
    del x
    yf = upfirdn2d(xf, k, pad=(2, 1))
    assert yf.shape == xf.shape
    ref = oracle.upfirdn2d(xf[:, 1:2, h - 40:, w - 72:].cpu(), k.cpu(), pad=(2, 1))
    np.testing.assert_allclose(yf[0, 1, h - 30:, w - 60:].cpu().numpy(), ref[0, 0, -30:, -60:].numpy(), rtol=1e-5, atol=1e-5)
