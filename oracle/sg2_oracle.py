"""Functional CPU restatement of the reference hot path (see oracle/__init__.py).

Everything here is a pure function of explicit tensors: there are no modules, no
globals and no argparse.  ``params`` is a dict keyed by the reference
``Generator.state_dict()`` names, so the same dict can be
loaded into the reference ``model.Generator`` with ``load_state_dict`` when the
golden vectors are produced.

Each function cites the reference lines it restates (paths relative to
/root/reference/).  Arithmetic is done in the dtype of the inputs (fp32 for the
parity tests, fp64 for identity checks).
"""
from __future__ import annotations

import math
from typing import Callable, Dict, List, Optional, Sequence, Tuple

import numpy as np
import torch
import torch.nn.functional as F

__all__ = [
    "CHANNEL_TABLE", "channels_for", "num_layers", "n_latent", "make_kernel",
    "upfirdn2d", "upfirdn2d_loops", "upfirdn2d_out_size", "fused_bias_act",
    "fused_leaky_relu", "equal_linear", "pixel_norm", "mapping",
    "modulated_conv2d", "modulated_conv2d_unfused", "styled_conv", "to_rgb",
    "synthesis", "generator_forward", "latent_from_alpha", "embed_fingerprint",
    "generate_with_alpha", "alpha_bound", "lr_at", "get_noise",
    "latin_hypercube_centered", "adam_step", "attribute_one_guess",
    "attribute_image", "decode_key", "mse_loss",
]

# src/model.py:418-428
CHANNEL_TABLE = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256, 128: 128, 256: 64,
                 512: 32, 1024: 16}


def channels_for(res: int, channel_multiplier: int = 2) -> int:
    """Channel count at a resolution (src/model.py:418-428)."""
    base = CHANNEL_TABLE[res]
    return base if res <= 32 else base * channel_multiplier


def num_layers(size: int) -> int:
    """Number of noise inputs, (log2(size)-2)*2+1 (src/model.py:437-438)."""
    return (int(math.log2(size)) - 2) * 2 + 1


def n_latent(size: int) -> int:
    """Number of per-layer latent slots, 2*log2(size)-2 (src/model.py:474)."""
    return int(math.log2(size)) * 2 - 2


# --------------------------------------------------------------------------
# FIR resampling  (src/op/upfirdn2d.py, src/op/upfirdn2d_kernel.cu)
# --------------------------------------------------------------------------

def make_kernel(taps: Sequence[float], dtype=torch.float32) -> torch.Tensor:
    """Separable FIR taps -> normalised 2-D kernel (src/model.py:22-30)."""
    k = torch.as_tensor(list(taps), dtype=dtype)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


def _pair(v):
    if isinstance(v, (tuple, list)):
        return int(v[0]), int(v[1])
    return int(v), int(v)


def _pad4(pad):
    pad = tuple(int(p) for p in pad)
    if len(pad) == 2:  # (p0, p1) applies to both axes (src/op/upfirdn2d.py:156-157)
        return pad[0], pad[1], pad[0], pad[1]
    return pad


def upfirdn2d_out_size(in_h, in_w, kh, kw, up, down, pad):
    """Output extent (src/op/upfirdn2d.py:104-105, upfirdn2d_kernel.cu:236-239)."""
    up_x, up_y = _pair(up)
    dn_x, dn_y = _pair(down)
    px0, px1, py0, py1 = _pad4(pad)
    out_h = (in_h * up_y + py0 + py1 - kh + dn_y) // dn_y
    out_w = (in_w * up_x + px0 + px1 - kw + dn_x) // dn_x
    return out_h, out_w


def upfirdn2d(x: torch.Tensor, kernel: torch.Tensor, up=1, down=1, pad=(0, 0)) -> torch.Tensor:
    """Zero-insert upsample, pad/crop, true convolution with ``kernel``, decimate.

    Restates ``upfirdn2d`` / ``upfirdn2d_native`` (src/op/upfirdn2d.py:149-209)
    and therefore the CUDA kernels it stands in for
    (src/op/upfirdn2d_kernel.cu:107-207): the FIR is applied *flipped* (:137),
    negative pads crop, the result keeps rows/cols ``0, down, 2*down, ...``.
    ``x`` is ``[N, C, H, W]``; ``up``/``down`` are ints or ``(x, y)`` pairs;
    ``pad`` is ``(p0, p1)`` or ``(x0, x1, y0, y1)``.
    """
    up_x, up_y = _pair(up)
    dn_x, dn_y = _pair(down)
    px0, px1, py0, py1 = _pad4(pad)
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    planes = x.reshape(n * c, 1, h, w)
    # zero-insertion: sample (y, x) lands on (y*up_y, x*up_x); trailing zeros are kept
    stuffed = planes.new_zeros(n * c, 1, h * up_y, w * up_x)
    stuffed[:, :, ::up_y, ::up_x] = planes
    # F.pad crops for negative amounts, which is what the reference does in two steps
    stuffed = F.pad(stuffed, [px0, px1, py0, py1])
    taps = torch.flip(kernel.to(x.dtype), [0, 1]).reshape(1, 1, kh, kw)
    full = F.conv2d(stuffed, taps)
    out = full[:, :, ::dn_y, ::dn_x]
    out_h, out_w = upfirdn2d_out_size(h, w, kh, kw, (up_x, up_y), (dn_x, dn_y), (px0, px1, py0, py1))
    return out.reshape(n, c, out_h, out_w)


def upfirdn2d_loops(x: np.ndarray, kernel: np.ndarray, up=1, down=1, pad=(0, 0)) -> np.ndarray:
    """Direct per-output-sample definition in numpy (small inputs only).

    Follows the index arithmetic of the generic CUDA kernel
    (src/op/upfirdn2d_kernel.cu:49-105): for output ``(oy, ox)`` the position on
    the zero-stuffed, padded grid is ``mid = o*down - pad0``; tap ``t`` of the
    un-flipped kernel meets stuffed sample ``mid + (k-1-t)``, which is a real
    input sample only when divisible by ``up``.  Accumulates in float64.
    """
    up_x, up_y = _pair(up)
    dn_x, dn_y = _pair(down)
    px0, px1, py0, py1 = _pad4(pad)
    n, c, h, w = x.shape
    kh, kw = kernel.shape
    out_h, out_w = upfirdn2d_out_size(h, w, kh, kw, (up_x, up_y), (dn_x, dn_y), (px0, px1, py0, py1))
    out = np.zeros((n, c, out_h, out_w), dtype=np.float64)
    xd = x.astype(np.float64)
    kd = kernel.astype(np.float64)
    for oy in range(out_h):
        for ox in range(out_w):
            acc = np.zeros((n, c), dtype=np.float64)
            for ty in range(kh):
                sy = oy * dn_y - py0 + (kh - 1 - ty)
                if sy < 0 or sy % up_y or sy // up_y >= h:
                    continue
                for tx in range(kw):
                    sx = ox * dn_x - px0 + (kw - 1 - tx)
                    if sx < 0 or sx % up_x or sx // up_x >= w:
                        continue
                    acc += xd[:, :, sy // up_y, sx // up_x] * kd[ty, tx]
            out[:, :, oy, ox] = acc
    return out.astype(x.dtype)


# --------------------------------------------------------------------------
# bias + leaky-ReLU + gain  (src/op/fused_act.py, fused_bias_act_kernel.cu)
# --------------------------------------------------------------------------

def fused_bias_act(x: torch.Tensor, bias: Optional[torch.Tensor], ref: Optional[torch.Tensor],
                   act: int, grad: int, alpha: float, scale: float) -> torch.Tensor:
    """The native op's arithmetic (src/op/fused_bias_act_kernel.cu:19-65).

    ``x += b[(i / step_b) % size_b]`` (:37-39) with ``step_b`` the product of
    dims >= 2 (:84-88); ``act*10+grad``: 10/11 identity, 30 ``x>0 ? x : x*alpha``,
    31 ``ref>0 ? x : x*alpha``, 12/32 zero (:40-60); result times ``scale`` (:62).
    ``None`` (or an empty tensor) means "absent", as in the reference.
    """
    if bias is not None and bias.numel():
        shape = [1, -1] + [1] * (x.ndim - 2)
        x = x + bias.reshape(shape).to(x.dtype)
    code = act * 10 + grad
    if code in (12, 32):
        y = torch.zeros_like(x)
    elif code == 30:
        y = torch.where(x > 0, x, x * alpha)
    elif code == 31:
        r = ref if (ref is not None and ref.numel()) else torch.zeros_like(x)
        y = torch.where(r > 0, x, x * alpha)
    else:
        y = x
    return y * scale


def fused_leaky_relu(x: torch.Tensor, bias: Optional[torch.Tensor] = None,
                     negative_slope: float = 0.2, scale: float = 2 ** 0.5) -> torch.Tensor:
    """``lrelu(x + b[c]) * scale`` (src/op/fused_act.py:110-127).

    Follows the GPU branch (``fused_bias_act`` with act=3, grad=0, :72), which
    honours ``negative_slope``; the reference's CPU branch hard-codes 0.2
    (:116,122) - identical at the default slope.
    """
    return fused_bias_act(x, bias, None, 3, 0, negative_slope, scale)


# --------------------------------------------------------------------------
# Mapping network and linear layers  (src/model.py:14-19, 132-161, 407-416)
# --------------------------------------------------------------------------

def equal_linear(x, weight, bias, lr_mul: float = 1.0, activation: bool = False):
    """Equalised-LR linear (src/model.py:151-161)."""
    scale = (1.0 / math.sqrt(weight.shape[1])) * lr_mul
    if activation:
        return fused_leaky_relu(F.linear(x, weight * scale), bias * lr_mul)
    return F.linear(x, weight * scale, bias * lr_mul)


def pixel_norm(x):
    """src/model.py:18-19."""
    return x * torch.rsqrt(torch.mean(x * x, dim=1, keepdim=True) + 1e-8)


def mapping(params: Dict[str, torch.Tensor], z: torch.Tensor, n_mlp: int = 8, lr_mlp: float = 0.01):
    """z -> w: PixelNorm then ``n_mlp`` fused-lrelu linears (src/model.py:407-416)."""
    h = pixel_norm(z)
    for i in range(1, n_mlp + 1):
        h = equal_linear(h, params[f"style.{i}.weight"], params[f"style.{i}.bias"],
                         lr_mul=lr_mlp, activation=True)
    return h


# --------------------------------------------------------------------------
# Modulated convolution  (src/model.py:169-302)
# --------------------------------------------------------------------------

def modulated_conv2d(x, style, weight, mod_weight, mod_bias, *, demodulate=True,
                     upsample=False, blur_kernel=(1, 3, 3, 1)):
    """Per-sample modulated (de-modulated) conv, the reference's fused algebra.

    src/model.py:258-302: ``s = Linear(style)`` (:258, bias_init 1, lr_mul 1);
    ``w = scale * W * s`` (:259); ``demod = rsqrt(sum w^2 + 1e-8)`` (:261-263);
    plain: grouped ``conv2d(pad=k//2)`` (:292-300); upsample: grouped
    ``conv_transpose2d(stride 2, pad 0)`` then ``Blur(kernel*4, pad=(1,1))``
    (:269-282, :186-197).  ``weight`` is ``[1, Cout, Cin, k, k]``.
    """
    b, cin, h, w_ = x.shape
    _, cout, _, k, _ = weight.shape
    s = equal_linear(style, mod_weight, mod_bias)                 # [B, Cin]
    wscale = 1.0 / math.sqrt(cin * k * k)
    wmod = (wscale * weight) * s.reshape(b, 1, cin, 1, 1)         # [B, Cout, Cin, k, k]
    if demodulate:
        d = torch.rsqrt(wmod.pow(2).sum(dim=(2, 3, 4)) + 1e-8)    # [B, Cout]
        wmod = wmod * d.reshape(b, cout, 1, 1, 1)
    xg = x.reshape(1, b * cin, h, w_)
    if upsample:
        wt = wmod.permute(0, 2, 1, 3, 4).reshape(b * cin, cout, k, k)
        y = F.conv_transpose2d(xg, wt, stride=2, padding=0, groups=b)
        y = y.reshape(b, cout, y.shape[2], y.shape[3])
        factor = 2
        p = (len(blur_kernel) - factor) - (k - 1)
        pad = ((p + 1) // 2 + factor - 1, p // 2 + 1)
        fir = make_kernel(blur_kernel, dtype=x.dtype) * (factor ** 2)
        return upfirdn2d(y, fir, pad=pad)
    y = F.conv2d(xg, wmod.reshape(b * cout, cin, k, k), padding=k // 2, groups=b)
    return y.reshape(b, cout, y.shape[2], y.shape[3])


def modulated_conv2d_unfused(x, style, weight, mod_weight, mod_bias, *, demodulate=True,
                             upsample=False, blur_kernel=(1, 3, 3, 1)):
    """Same result via activation modulation + output demodulation.

    The reference's ``fused=False`` branch (src/model.py:229-256): scale the
    input channels by ``s``, convolve with the *shared* weight, scale output
    channels by ``demod``.  This is the algebra the CUDA path uses.
    """
    b, cin, h, w_ = x.shape
    _, cout, _, k, _ = weight.shape
    ws = weight[0] * (1.0 / math.sqrt(cin * k * k))               # [Cout, Cin, k, k]
    s = equal_linear(style, mod_weight, mod_bias)
    xm = x * s.reshape(b, cin, 1, 1)
    if upsample:
        y = F.conv_transpose2d(xm, ws.transpose(0, 1), stride=2, padding=0)
        fir = make_kernel(blur_kernel, dtype=x.dtype) * 4
        y = upfirdn2d(y, fir, pad=(1, 1))
    else:
        y = F.conv2d(xm, ws, padding=k // 2)
    if demodulate:
        wsq = ws.pow(2).sum(dim=(2, 3))                           # [Cout, Cin]
        d = torch.rsqrt((s * s) @ wsq.t() + 1e-8)                 # [B, Cout]
        y = y * d.reshape(b, cout, 1, 1)
    return y


def lrelu_with_sign(x, bias, sign_mask, negative_slope: float = 0.2, scale: float = 2 ** 0.5):
    """``fused_leaky_relu`` with the branch of every unit taken from ``sign_mask`` (bool, True =
    positive side) instead of from ``x + b > 0``.

    Checker-only helper: leaky-ReLU is continuous, so for a unit whose pre-activation is
    numerically zero the value is the same on either branch but the derivative is not
    (1 vs ``negative_slope``).  Two correct fp32 implementations pick different branches
    for such units (measured: 1-2 units per 10^5 between the reference's own fused and
    unfused algebra), which moves the latent gradient by ~1e-3.  Gradient parity is
    therefore checked with the branch pattern of the implementation under test imposed
    here, and the pattern itself is checked separately (it may differ only where
    ``|x + b|`` is at rounding level).
    """
    v = x + bias.reshape([1, -1] + [1] * (x.ndim - 2)).to(x.dtype)
    return torch.where(sign_mask, v, v * negative_slope) * scale


def styled_conv(params, prefix, x, style, noise, *, upsample=False, fused=True, sign_mask=None):
    """ModulatedConv2d -> noise injection -> bias + lrelu*sqrt2 (src/model.py:360-366)."""
    conv = modulated_conv2d if fused else modulated_conv2d_unfused
    y = conv(x, style, params[f"{prefix}.conv.weight"], params[f"{prefix}.conv.modulation.weight"],
             params[f"{prefix}.conv.modulation.bias"], demodulate=True, upsample=upsample)
    if noise is None:  # src/model.py:312-314
        noise = torch.randn(y.shape[0], 1, y.shape[2], y.shape[3], dtype=y.dtype)
    y = y + params[f"{prefix}.noise.weight"] * noise              # src/model.py:316
    if sign_mask is not None:
        return lrelu_with_sign(y, params[f"{prefix}.activate.bias"], sign_mask)
    return fused_leaky_relu(y, params[f"{prefix}.activate.bias"])


def to_rgb(params, prefix, x, style, skip=None, *, fused=True):
    """1x1 modulated conv (no demod) + bias + upsampled skip (src/model.py:379-388)."""
    conv = modulated_conv2d if fused else modulated_conv2d_unfused
    y = conv(x, style, params[f"{prefix}.conv.weight"], params[f"{prefix}.conv.modulation.weight"],
             params[f"{prefix}.conv.modulation.bias"], demodulate=False)
    y = y + params[f"{prefix}.bias"]
    if skip is not None:
        fir = make_kernel((1, 3, 3, 1), dtype=x.dtype) * 4        # src/model.py:37-46
        y = y + upfirdn2d(skip, fir, up=2, down=1, pad=(2, 1))
    return y


def synthesis(params, latent, noise: Sequence[Optional[torch.Tensor]], *, fused=True,
              sign_masks=None, activations=None):
    """Latent ``[B, n_latent, 512]`` -> image ``[B, 3, S, S]`` (src/model.py:551-566).

    ``sign_masks`` (checker-only, see ``lrelu_with_sign``): one bool tensor per StyledConv in
    execution order.  ``activations``: optional list that receives every StyledConv output.
    """
    b = latent.shape[0]
    masks = list(sign_masks) if sign_masks is not None else [None] * (latent.shape[1] - 1)
    keep = activations if activations is not None else []
    x = params["input.input"].repeat(b, 1, 1, 1)                  # src/model.py:325-329
    x = styled_conv(params, "conv1", x, latent[:, 0], noise[0], fused=fused, sign_mask=masks[0])
    keep.append(x)
    skip = to_rgb(params, "to_rgb1", x, latent[:, 1], fused=fused)
    n_blocks = (latent.shape[1] - 2) // 2
    slot = 1
    for j in range(n_blocks):
        x = styled_conv(params, f"convs.{2 * j}", x, latent[:, slot], noise[1 + 2 * j],
                        upsample=True, fused=fused, sign_mask=masks[1 + 2 * j])
        keep.append(x)
        x = styled_conv(params, f"convs.{2 * j + 1}", x, latent[:, slot + 1], noise[2 + 2 * j],
                        fused=fused, sign_mask=masks[2 + 2 * j])
        keep.append(x)
        skip = to_rgb(params, f"to_rgbs.{j}", x, latent[:, slot + 2], skip, fused=fused)
        slot += 2
    return skip


def generator_forward(params, styles: List[torch.Tensor], size: int, *, input_is_latent=False,
                      noise=None, fused=True):
    """``Generator.forward`` for a single style (src/model.py:499-572)."""
    w = styles[0] if input_is_latent else mapping(params, styles[0])
    nl = n_latent(size)
    latent = w.unsqueeze(1).repeat(1, nl, 1) if w.ndim < 3 else w  # src/model.py:531-535
    if noise is None:
        noise = [None] * num_layers(size)
    return synthesis(params, latent, noise, fused=fused)


# --------------------------------------------------------------------------
# Fingerprint embed and the attribution loop  (src/generator.py, src/main.py)
# --------------------------------------------------------------------------

def latent_from_alpha(u_cap, alpha, latent_mean):
    """``w0 = U^T alpha + mu`` with column vectors (src/main.py:60)."""
    return u_cap.t() @ alpha + latent_mean


def embed_fingerprint(v_cap, sigma_key, k, w0, sd: float = 1.0):
    """``wx = w0 + sd * V^T diag(sigma) k`` (src/generator.py:148-161).

    ``v_cap`` is ``[key_len, 512]``, ``sigma_key`` ``[key_len, 1]``, ``k``
    ``[key_len, 1]`` (already through the sigmoid in the loop, src/main.py:61),
    ``w0`` ``[512, 1]``.  Same association order as the reference:
    ``(V^T diag(sigma)) k``.
    """
    vs = v_cap.t() @ torch.diag(sigma_key.reshape(-1))
    return w0 + sd * (vs @ k.reshape(-1, 1))


def generate_with_alpha(params, size, alpha, u_cap, v_cap, sigma_key, latent_mean, key, noise,
                        sd: float = 1.0):
    """Target image for a given key (src/generator.py:69-107); ``alpha`` is ``[n_main, B]``,
    ``key`` integer ``[key_len, B]``.  Returns ``(image, w0 [B,512], wx [B,512])``."""
    w0 = (u_cap.t() @ alpha + latent_mean).t()                    # :83
    sk = sigma_key * key.to(sigma_key.dtype)                      # :85
    wx = w0 + sd * (sk.t() @ v_cap)                               # :89
    img = generator_forward(params, [wx], size, input_is_latent=True, noise=noise)
    return img.detach(), w0.detach(), wx.detach()


def alpha_bound(alpha, upper, lower):
    """``sum relu(alpha-upper) + sum relu(lower-alpha)`` (src/utils.py:53-58)."""
    return torch.relu(alpha - upper).sum() + torch.relu(lower - alpha).sum()


def lr_at(step: int, lr0: float = 0.2) -> float:
    """``lr0 * exp(-0.001 (step+1))`` (src/main.py:42-43)."""
    return lr0 * math.exp(-0.001 * (step + 1))


def get_noise(size: int, dtype=torch.float32) -> List[torch.Tensor]:
    """Fixed noise maps (src/utils.py:128-138): the 4x4 map from ``default_rng(2002)``,
    the others from the *global* numpy RNG, which ``Generator.__init__`` seeds with 2022
    (src/model.py:404); callers that need the reference's exact maps must seed likewise."""
    rng = np.random.default_rng(seed=2002)
    maps = [torch.tensor(rng.standard_normal((1, 1, 4, 4)), dtype=dtype)]
    for i in range(3, int(math.log2(size)) + 1):
        for _ in range(2):
            maps.append(torch.tensor(np.random.standard_normal((1, 1, 2 ** i, 2 ** i)), dtype=dtype))
    return maps


def latin_hypercube_centered(n: int, d: int, rng: np.random.Generator) -> np.ndarray:
    """Centred Latin hypercube (src/main.py:103, ``LatinHypercube(centered=True)`` of
    scipy 1.7): per dimension a random permutation of the cell centres ``(i+0.5)/n``."""
    cols = [(rng.permutation(n) + 0.5) / n for _ in range(d)]
    return np.stack(cols, axis=1)


def mse_loss(a, b):
    """``F.mse_loss`` (src/utils.py:46-47)."""
    return torch.mean((a - b) ** 2)


def adam_step(p, g, m, v, t: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """``torch.optim.Adam`` defaults, single tensor, step ``t`` (1-based), in place
    (the optimiser at src/main.py:56,70)."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** t
    bc2 = 1 - beta2 ** t
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-(lr / bc1))


def decode_key(key_logits):
    """``round(sigmoid(key))`` (src/main.py:72,88)."""
    return torch.round(torch.sigmoid(key_logits))


def attribute_one_guess(render: Callable[[torch.Tensor], torch.Tensor], target_img, alpha0, u_cap,
                        v_cap, sigma_key, latent_mean, max_alpha, min_alpha, *, steps: int,
                        lr0: float = 0.2, sd: float = 1.0, key_len: int = 64,
                        loss_fn: Callable = mse_loss, optimise_alpha: bool = True,
                        trace: Optional[list] = None):
    """One trajectory of the hot loop (src/main.py:51-81).

    ``render(wx[512,1]) -> image [1,3,S,S]`` is the frozen generator
    (src/generator.py:170-174).  Order per step: forward, loss
    ``= loss_fn + 0.1*alpha_bound`` (:65), set lr (:67), backward (:69), Adam (:70).
    ``optimise_alpha=False`` gives the well-posed key-only fixture.
    Returns ``(final loss as computed in the last step, alpha, key_logits)``.
    """
    alpha = alpha0.clone().reshape(-1, 1).requires_grad_(optimise_alpha)
    key = torch.zeros(key_len, 1, dtype=alpha.dtype, requires_grad=True)   # src/utils.py:19-21
    tensors = [alpha, key] if optimise_alpha else [key]
    moments = [(torch.zeros_like(t), torch.zeros_like(t)) for t in tensors]
    loss_val = float("nan")
    for i in range(steps):
        w0 = latent_from_alpha(u_cap, alpha, latent_mean)
        wx = embed_fingerprint(v_cap, sigma_key, torch.sigmoid(key), w0, sd)
        est = render(wx)
        loss = loss_fn(target_img, est) + 0.1 * alpha_bound(alpha, max_alpha, min_alpha)
        grads = torch.autograd.grad(loss, tensors)
        lr = lr_at(i, lr0)
        with torch.no_grad():
            for t, g, (m, v) in zip(tensors, grads, moments):
                adam_step(t, g, m, v, i + 1, lr)
        loss_val = float(loss.detach())
        if trace is not None:
            trace.append((loss_val, alpha.detach().clone(), key.detach().clone()))
    return loss_val, alpha.detach(), key.detach()


def attribute_image(render, target_img, guesses, u_cap, v_cap, sigma_key, sigma_main, latent_mean,
                    max_alpha, min_alpha, **kw):
    """All LHS guesses of one image and the arg-min pick (src/main.py:45-89).

    ``guesses`` is ``[n, n_main]`` in (0,1); each becomes
    ``alpha0 = 2*u*sigma_main - sigma_main`` (:52)."""
    results = []
    for u in guesses:
        a0 = 2 * u.reshape(-1, 1) * sigma_main - sigma_main
        results.append(attribute_one_guess(render, target_img, a0, u_cap, v_cap, sigma_key,
                                           latent_mean, max_alpha, min_alpha, **kw))
    best = min(range(len(results)), key=lambda i: results[i][0])
    return results[best], results
