"""StyleGAN2 generator behind the reference's ``model`` API (src/model.py), B200-native.

Constructor signatures, parameter/buffer names and shapes (hence ``state_dict`` keys) match the
reference, so its checkpoints load unchanged.  What differs is execution:

* ``Generator.forward`` runs the whole synthesis network (ConstantInput -> StyledConv/ToRGB
  chain, src/model.py:551-566) as ONE native call into liblfp_sg2 (``lfp_native.synthesis``),
  differentiable w.r.t. the latent.  Generator parameters are frozen constants on this path -
  the only consumers of this generator (src/generator.py, src/main.py) never use their
  gradients.
* The building blocks (``Blur``, ``Upsample``, ``ModulatedConv2d``, ``StyledConv``, ``ToRGB`` ...)
  remain usable on their own; their FIR / bias-act work goes through the drop-in ``op`` package.

Only CUDA tensors are accepted: there is no CPU fallback.
"""
import math

import numpy as np
import torch
from torch import nn
from torch.nn import functional as F

from op import FusedLeakyReLU, fused_leaky_relu, upfirdn2d, conv2d_gradfix
from lfp_native import capi as _capi
from lfp_native.synthesis import SynthesisPlan, synthesize


def make_kernel(k):
    """1-D taps -> normalised separable 2-D FIR (src/model.py:22-30)."""
    k = torch.tensor(k, dtype=torch.float32)
    if k.ndim == 1:
        k = torch.outer(k, k)
    return k / k.sum()


class PixelNorm(nn.Module):
    def forward(self, input):
        return input * torch.rsqrt(input.pow(2).mean(dim=1, keepdim=True) + 1e-8)


class _Resample(nn.Module):
    """Shared body of Upsample / Downsample / Blur: one ``upfirdn2d`` with a registered FIR."""

    def __init__(self, kernel, up, down, pad):
        super().__init__()
        self.register_buffer("kernel", kernel)
        self.up, self.down, self.pad = up, down, pad

    def forward(self, input):
        return upfirdn2d(input, self.kernel, up=self.up, down=self.down, pad=self.pad)


class Upsample(_Resample):
    """x2 FIR upsampling of the RGB skip (src/model.py:33-51)."""

    def __init__(self, kernel, factor=2):
        fir = make_kernel(kernel) * (factor ** 2)
        p = fir.shape[0] - factor
        super().__init__(fir, factor, 1, ((p + 1) // 2 + factor - 1, p // 2))
        self.factor = factor


class Downsample(_Resample):
    """x2 FIR decimation (src/model.py:54-72)."""

    def __init__(self, kernel, factor=2):
        fir = make_kernel(kernel)
        p = fir.shape[0] - factor
        super().__init__(fir, 1, factor, ((p + 1) // 2, p // 2))
        self.factor = factor


class Blur(_Resample):
    """FIR blur with explicit padding (src/model.py:75-91)."""

    def __init__(self, kernel, pad, upsample_factor=1):
        fir = make_kernel(kernel)
        if upsample_factor > 1:
            fir = fir * (upsample_factor ** 2)
        super().__init__(fir, 1, 1, pad)


class EqualConv2d(nn.Module):
    """Equalised-LR conv (src/model.py:94-129); discriminator-side, kept for API parity."""

    def __init__(self, in_channel, out_channel, kernel_size, stride=1, padding=0, bias=True):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_channel, in_channel, kernel_size, kernel_size))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.stride, self.padding = stride, padding
        self.bias = nn.Parameter(torch.zeros(out_channel)) if bias else None

    def forward(self, input):
        return conv2d_gradfix.conv2d(input, self.weight * self.scale, bias=self.bias, stride=self.stride,
                                     padding=self.padding)


class EqualLinear(nn.Module):
    """Equalised-LR linear, optionally followed by the fused bias+lrelu (src/model.py:132-166)."""

    def __init__(self, in_dim, out_dim, bias=True, bias_init=0, lr_mul=1, activation=None):
        super().__init__()
        self.weight = nn.Parameter(torch.randn(out_dim, in_dim).div_(lr_mul))
        self.bias = nn.Parameter(torch.full((out_dim,), float(bias_init))) if bias else None
        self.activation = activation
        self.scale = (1 / math.sqrt(in_dim)) * lr_mul
        self.lr_mul = lr_mul

    def forward(self, input):
        if self.activation:
            return fused_leaky_relu(F.linear(input, self.weight * self.scale), self.bias * self.lr_mul)
        return F.linear(input, self.weight * self.scale, bias=self.bias * self.lr_mul)


class ModulatedConv2d(nn.Module):
    """Style-modulated, demodulated conv (src/model.py:169-302).

    Stand-alone use runs the native layer (``lfp_modconv_forward / backward``, include/lfp_sg2.h group 5): the
    activation-modulated form (the reference's ``fused=False`` branch, :229-256) on the same gather-convolution kernels
    as the whole-synthesis plan - tcgen05 when ``torch.backends.cudnn.allow_tf32`` (what decides the reference's conv
    arithmetic) and the shape tiles, CUDA-core fp32 otherwise - differentiable w.r.t. the input and the style; the
    layer's parameters are frozen constants there, exactly as on the ``Generator.forward`` path.  ``native = False`` (or
    a downsampling layer, which the generator never builds) evaluates the same algebra with torch ops through
    ``conv2d_gradfix``, which also yields parameter gradients.  Inside ``Generator`` the parameters are consumed by the
    synthesis plan instead and this ``forward`` is not called.
    """

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, demodulate=True, upsample=False,
                 downsample=False, blur_kernel=[1, 3, 3, 1], fused=True):
        super().__init__()
        self.eps = 1e-8
        self.kernel_size, self.in_channel, self.out_channel = kernel_size, in_channel, out_channel
        self.upsample, self.downsample = upsample, downsample
        if upsample:
            p = (len(blur_kernel) - 2) - (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2 + 1, p // 2 + 1), upsample_factor=2)
        if downsample:
            p = (len(blur_kernel) - 2) + (kernel_size - 1)
            self.blur = Blur(blur_kernel, pad=((p + 1) // 2, p // 2))
        self.scale = 1 / math.sqrt(in_channel * kernel_size ** 2)
        self.padding = kernel_size // 2
        self.weight = nn.Parameter(torch.randn(1, out_channel, in_channel, kernel_size, kernel_size))
        self.modulation = EqualLinear(style_dim, in_channel, bias_init=1)
        self.demodulate = demodulate
        self.fused = fused
        self.blur_kernel = list(blur_kernel)
        self.native = True        # False: torch-op composite (parameter gradients)
        self.precision = None     # None: follow torch.backends.cudnn.allow_tf32
        self._plans = {}

    def _plan(self, device):
        from lfp_native.modconv import ModConvPlan
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        plan = self._plans.get(key)
        if plan is None:
            plan = ModConvPlan(self.in_channel, self.out_channel, self.kernel_size, self.modulation.weight.shape[1],
                               self.demodulate, self.upsample, self.blur_kernel, device=torch.device("cuda", key[1]))
            self._plans[key] = plan
        plan.sync(self.weight, self.modulation.weight, self.modulation.bias)
        return plan

    def forward(self, input, style):
        if self.native and not self.downsample and self.kernel_size in (1, 3) and len(self.blur_kernel) == 4:
            if not input.is_cuda:
                raise RuntimeError("input must be a CUDA tensor")
            from lfp_native.modconv import modulated_conv2d
            prec = self.precision
            if prec is None:
                prec = _capi.PREC_TF32 if torch.backends.cudnn.allow_tf32 else _capi.PREC_FP32
            return modulated_conv2d(self._plan(input.device), input, style, prec)
        batch = input.shape[0]
        w = self.scale * self.weight[0]
        s = self.modulation(style)
        x = input * s.view(batch, self.in_channel, 1, 1)
        if self.upsample:
            out = self.blur(conv2d_gradfix.conv_transpose2d(x, w.transpose(0, 1), padding=0, stride=2))
        elif self.downsample:
            out = conv2d_gradfix.conv2d(self.blur(x), w, padding=0, stride=2)
        else:
            out = conv2d_gradfix.conv2d(x, w, padding=self.padding)
        if self.demodulate:
            d = torch.rsqrt((s * s) @ w.pow(2).sum(dim=(2, 3)).t() + self.eps)
            out = out * d.view(batch, self.out_channel, 1, 1)
        return out


class NoiseInjection(nn.Module):
    """``image + weight * noise`` (src/model.py:305-316)."""

    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))

    def forward(self, image, noise=None):
        if noise is None:
            b, _, h, w = image.shape
            noise = image.new_empty(b, 1, h, w).normal_()
        return image + self.weight * noise


class ConstantInput(nn.Module):
    """Learned 4x4 constant (src/model.py:319-329)."""

    def __init__(self, channel, size=4):
        super().__init__()
        self.input = nn.Parameter(torch.randn(1, channel, size, size))

    def forward(self, input):
        return self.input.repeat(input.shape[0], 1, 1, 1)


class StyledConv(nn.Module):
    """ModulatedConv2d -> NoiseInjection -> FusedLeakyReLU (src/model.py:332-366)."""

    def __init__(self, in_channel, out_channel, kernel_size, style_dim, upsample=False, blur_kernel=[1, 3, 3, 1],
                 demodulate=True):
        super().__init__()
        self.conv = ModulatedConv2d(in_channel, out_channel, kernel_size, style_dim, upsample=upsample,
                                    blur_kernel=blur_kernel, demodulate=demodulate)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(out_channel)

    def forward(self, input, style, noise=None):
        return self.activate(self.noise(self.conv(input, style), noise=noise))


class ToRGB(nn.Module):
    """1x1 modulated conv (no demod) + bias + upsampled skip (src/model.py:369-388)."""

    def __init__(self, in_channel, style_dim, upsample=True, blur_kernel=[1, 3, 3, 1]):
        super().__init__()
        if upsample:
            self.upsample = Upsample(blur_kernel)
        self.conv = ModulatedConv2d(in_channel, 3, 1, style_dim, demodulate=False)
        self.bias = nn.Parameter(torch.zeros(1, 3, 1, 1))

    def forward(self, input, style, skip=None):
        out = self.conv(input, style) + self.bias
        if skip is not None:
            out = out + self.upsample(skip)
        return out


class Generator(nn.Module):
    """``Generator(size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1,3,3,1], lr_mlp=0.01)``
    (src/model.py:391-572)."""

    def __init__(self, size, style_dim, n_mlp, channel_multiplier=2, blur_kernel=[1, 3, 3, 1], lr_mlp=0.01):
        super().__init__()
        # the reference reseeds the GLOBAL numpy RNG here (src/model.py:404); utils.get_noise() (src/utils.py:128-138)
        # draws every noise map after the 4x4 from that global stream, so callers that build their noise the reference
        # way get the reference's maps only if the constructor does the same
        np.random.seed(2022)
        self.size, self.style_dim = size, style_dim
        self.channel_multiplier, self.blur_kernel = channel_multiplier, list(blur_kernel)
        self.style = nn.Sequential(PixelNorm(), *[
            EqualLinear(style_dim, style_dim, lr_mul=lr_mlp, activation="fused_lrelu") for _ in range(n_mlp)])
        cm = channel_multiplier
        self.channels = {4: 512, 8: 512, 16: 512, 32: 512, 64: 256 * cm, 128: 128 * cm, 256: 64 * cm,
                         512: 32 * cm, 1024: 16 * cm}
        self.input = ConstantInput(self.channels[4])
        self.conv1 = StyledConv(self.channels[4], self.channels[4], 3, style_dim, blur_kernel=blur_kernel)
        self.to_rgb1 = ToRGB(self.channels[4], style_dim, upsample=False)
        self.log_size = int(math.log(size, 2))
        self.num_layers = (self.log_size - 2) * 2 + 1
        self.convs, self.upsamples, self.to_rgbs = nn.ModuleList(), nn.ModuleList(), nn.ModuleList()
        self.noises = nn.Module()
        for layer_idx in range(self.num_layers):
            res = (layer_idx + 5) // 2
            self.noises.register_buffer(f"noise_{layer_idx}", torch.randn(1, 1, 2 ** res, 2 ** res))
        in_channel = self.channels[4]
        for i in range(3, self.log_size + 1):
            out_channel = self.channels[2 ** i]
            self.convs.append(StyledConv(in_channel, out_channel, 3, style_dim, upsample=True,
                                         blur_kernel=blur_kernel))
            self.convs.append(StyledConv(out_channel, out_channel, 3, style_dim, blur_kernel=blur_kernel))
            self.to_rgbs.append(ToRGB(out_channel, style_dim))
            in_channel = out_channel
        self.n_latent = self.log_size * 2 - 2
        # convolution arithmetic of the native path: None = follow torch.backends.cudnn.allow_tf32, which is what
        # decides it for the reference's F.conv2d / F.conv_transpose2d calls (True by default);
        # or pin it with lfp_native.capi.PREC_FP32 / PREC_TF32
        self.precision = None
        self._plans = {}

    # ---- helpers kept from the reference API (src/model.py:476-497) ---------------------------
    def make_noise(self):
        device = self.input.input.device
        noises = [torch.randn(1, 1, 4, 4, device=device)]
        for i in range(3, self.log_size + 1):
            noises += [torch.randn(1, 1, 2 ** i, 2 ** i, device=device) for _ in range(2)]
        return noises

    def mean_latent(self, n_latent):
        z = torch.randn(n_latent, self.style_dim, device=self.input.input.device)
        return self.style(z).mean(0, keepdim=True)

    def get_latent(self, input):
        return self.style(input)

    # ---- native synthesis -----------------------------------------------------------------------
    def _plan(self):
        device = self.input.input.device
        if device.type != "cuda":
            raise RuntimeError("Generator parameters must be CUDA tensors: this implementation has no CPU path")
        key = (device.type, device.index if device.index is not None else torch.cuda.current_device())
        plan = self._plans.get(key)
        if plan is None:
            plan = SynthesisPlan(self.size, self.style_dim, self.channel_multiplier, self.blur_kernel,
                                 device=torch.device("cuda", key[1]))
            self._plans[key] = plan
        plan.sync_from_module(self)
        return plan

    def _noise_list(self, noise, fixed_noise, batch, device, dtype):
        if noise is None:
            noise = ([getattr(self.noises, f"noise_{i}") for i in range(self.num_layers)] if fixed_noise
                     else [None] * self.num_layers)
        out = []
        for i, n in enumerate(noise):
            if n is None:  # per-sample Gaussian noise, drawn in layer order like NoiseInjection does
                res = 4 if i == 0 else 8 << ((i - 1) // 2)
                n = torch.empty(batch, 1, res, res, device=device, dtype=dtype).normal_()
            out.append(n)
        return out

    def forward(self, styles, return_latents=False, get_latent_only=False, inject_index=None, truncation=1,
                truncation_latent=None, input_is_latent=False, noise=None, fixed_noise=False):
        if not input_is_latent:
            styles = [self.style(s) for s in styles]
            if truncation < 1:
                styles = [truncation_latent + truncation * (s - truncation_latent) for s in styles]
        if len(styles) < 2:
            inject_index = self.n_latent
            latent = styles[0].unsqueeze(1).repeat(1, inject_index, 1) if styles[0].ndim < 3 else styles[0]
        else:  # style mixing: first `inject_index` slots from styles[0], the rest from styles[1]
            if inject_index is None:
                inject_index = self.n_latent - 2
            latent = torch.cat([styles[0].unsqueeze(1).repeat(1, inject_index, 1),
                                styles[1].unsqueeze(1).repeat(1, self.n_latent - inject_index, 1)], 1)
        if get_latent_only:
            return latent
        plan = self._plan()
        nz = self._noise_list(noise, fixed_noise, latent.shape[0], latent.device, torch.float32)
        prec = self.precision
        if prec is None:
            prec = _capi.PREC_TF32 if torch.backends.cudnn.allow_tf32 else _capi.PREC_FP32
        image = synthesize(plan, latent, nz, prec)
        return (image, latent) if return_latents else (image, None)
