"""Oracle layers: plain-torch restatement of the CompressAI building blocks the hot path runs.

Test infrastructure only (see ``oracle/__init__.py``).  ``compressai`` is an un-vendored,
unpinned dependency of the reference (imported at ``/root/reference/anchors/model.py:3-5``);
the formulas below restate its published 1.1.x-1.2.x algorithm and are
anchored on the reference's own in-repo fragments cited per function.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


# --------------------------------------------------------------------------------------
# clamp-with-gradient (reference: utils/ops.py:28-56; compressai.ops.LowerBound, same rule)
# --------------------------------------------------------------------------------------
class _LowBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x)
        ctx.bound = float(bound)
        return torch.clamp(x, min=float(bound))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        keep = (x >= ctx.bound) | (g < 0.0)  # utils/ops.py:40
        return g * keep.to(g.dtype), None


class _UpBound(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, bound):
        ctx.save_for_backward(x)
        ctx.bound = float(bound)
        return torch.clamp(x, max=float(bound))

    @staticmethod
    def backward(ctx, g):
        (x,) = ctx.saved_tensors
        keep = (x <= ctx.bound) | (g > 0.0)  # utils/ops.py:55
        return g * keep.to(g.dtype), None


def low_bound(x, bound):
    """``ops.Low_bound.apply(x, b)`` (utils/ops.py:28-41)."""
    return _LowBound.apply(x, bound)


def up_bound(x, bound):
    """``ops.Up_bound.apply(x, b)`` (utils/ops.py:43-56)."""
    return _UpBound.apply(x, bound)


class LowerBound(nn.Module):
    """compressai.ops.LowerBound: max(x, b) forward, utils/ops.py:36-41 rule backward."""

    def __init__(self, bound):
        super().__init__()
        self.bound = float(bound)

    def forward(self, x):
        return low_bound(x, self.bound)


class NonNegativeParametrizer(nn.Module):
    """compressai.ops.NonNegativeParametrizer (A.1); same math as utils/ops.py:64-90."""

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.minimum = float(minimum)
        self.pedestal = float(reparam_offset) ** 2
        self.bound = (self.minimum + self.pedestal) ** 0.5
        self.lower_bound = LowerBound(self.bound)

    def init(self, x):
        return torch.sqrt(torch.clamp(x + self.pedestal, min=self.pedestal))

    def forward(self, x):
        return self.lower_bound(x) ** 2 - self.pedestal


class GDN(nn.Module):
    """Generalised divisive normalisation (A.1; reference pin utils/ops.py:58-97,
    attack_rd.py:294-298 uses ``beta_reparam``/``gamma_reparam``)."""

    def __init__(self, C, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        self.gamma_reparam = NonNegativeParametrizer()
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(C)))
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(C)))

    def effective(self):
        return self.beta_reparam(self.beta), self.gamma_reparam(self.gamma)

    def forward(self, x):
        C = x.shape[1]
        beta, gamma = self.effective()
        norm = F.conv2d(x * x, gamma.reshape(C, C, 1, 1), beta)
        norm = torch.sqrt(norm) if self.inverse else torch.rsqrt(norm)
        return x * norm


# --------------------------------------------------------------------------------------
# conv helpers (reference: anchors/utils.py:112-130, verbatim hyper-parameters)
# --------------------------------------------------------------------------------------
def conv(cin, cout, kernel_size=5, stride=2):
    return nn.Conv2d(cin, cout, kernel_size, stride=stride, padding=kernel_size // 2)


def deconv(cin, cout, kernel_size=5, stride=2):
    return nn.ConvTranspose2d(cin, cout, kernel_size, stride=stride,
                              output_padding=stride - 1, padding=kernel_size // 2)


def conv3x3(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 3, stride=stride, padding=1)


def conv1x1(cin, cout, stride=1):
    return nn.Conv2d(cin, cout, 1, stride=stride)


def subpel_conv3x3(cin, cout, r=1):
    return nn.Sequential(nn.Conv2d(cin, cout * r * r, 3, padding=1), nn.PixelShuffle(r))


class MaskedConv2d(nn.Conv2d):
    """Type-A/B masked conv (A.2; used by mbt2018 ``context_prediction``,
    reference call site anchors/model.py:103)."""

    def __init__(self, *args, mask_type="A", **kwargs):
        super().__init__(*args, **kwargs)
        self.register_buffer("mask", torch.ones_like(self.weight.data))
        _, _, h, w = self.mask.shape
        self.mask[:, :, h // 2, w // 2 + (mask_type == "B"):] = 0
        self.mask[:, :, h // 2 + 1:] = 0

    def forward(self, x):
        self.weight.data *= self.mask
        return super().forward(x)


class ResidualBlockWithStride(nn.Module):
    def __init__(self, cin, cout, stride=2):
        super().__init__()
        self.conv1 = conv3x3(cin, cout, stride)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(cout, cout)
        self.gdn = GDN(cout)
        self.skip = conv1x1(cin, cout, stride) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        idt = x if self.skip is None else self.skip(x)
        return out + idt


class ResidualBlockUpsample(nn.Module):
    def __init__(self, cin, cout, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(cin, cout, upsample)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv = conv3x3(cout, cout)
        self.igdn = GDN(cout, inverse=True)
        self.upsample = subpel_conv3x3(cin, cout, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return out + self.upsample(x)


class ResidualBlock(nn.Module):
    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = conv3x3(cin, cout)
        self.leaky_relu = nn.LeakyReLU(inplace=True)
        self.conv2 = conv3x3(cout, cout)
        self.skip = conv1x1(cin, cout) if cin != cout else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        idt = x if self.skip is None else self.skip(x)
        return out + idt


class AttentionBlock(nn.Module):
    """compressai.layers.AttentionBlock (cheng2020_attn): the simplified attention of
    Cheng 2020 -- ``conv_a(x) * sigmoid(conv_b(x)) + x``, three residual units (1x1 N->N/2, ReLU, 3x3, ReLU, 1x1
    N/2->N, + x, ReLU) per branch, a closing 1x1 on the gate branch.  20.5 N^2 + 13 N parameters."""

    class ResidualUnit(nn.Module):
        def __init__(self, N):
            super().__init__()
            self.conv = nn.Sequential(conv1x1(N, N // 2), nn.ReLU(inplace=True), conv3x3(N // 2, N // 2),
                                      nn.ReLU(inplace=True), conv1x1(N // 2, N))

        def forward(self, x):
            return F.relu(self.conv(x) + x)

    def __init__(self, N):
        super().__init__()
        U = AttentionBlock.ResidualUnit
        self.conv_a = nn.Sequential(U(N), U(N), U(N))
        self.conv_b = nn.Sequential(U(N), U(N), U(N), conv1x1(N, N))

    def forward(self, x):
        return self.conv_a(x) * torch.sigmoid(self.conv_b(x)) + x


# --------------------------------------------------------------------------------------
# entropy models (A.3 / A.4)
# --------------------------------------------------------------------------------------
def quantize(x, mode, means=None, noise=None):
    """compressai ``EntropyModel.quantize`` (call site anchors/model.py:102).

    ``noise`` lets a test inject the U(-.5,.5) sample so two implementations can share it.
    """
    if mode == "noise":
        if noise is None:
            noise = torch.empty_like(x).uniform_(-0.5, 0.5)
        return x + noise
    out = x.clone()
    if means is not None:
        out = out - means
    out = torch.round(out)
    if mode == "dequantize":
        if means is not None:
            out = out + means
        return out
    assert mode == "symbols", mode
    return out.int()


class EntropyBottleneck(nn.Module):
    """Factorised density model of Balle 2018 (A.3)."""

    def __init__(self, channels, tail_mass=1e-9, init_scale=10.0, filters=(3, 3, 3, 3),
                 likelihood_bound=1e-9):
        super().__init__()
        self.channels = int(channels)
        self.filters = tuple(int(f) for f in filters)
        self.init_scale = float(init_scale)
        self.tail_mass = float(tail_mass)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        C = self.channels
        for i in range(len(self.filters) + 1):
            init = np.log(np.expm1(1 / scale / f[i + 1]))
            m = torch.empty(C, f[i + 1], f[i]).fill_(float(init))
            self.register_parameter(f"_matrix{i}", nn.Parameter(m))
            b = torch.empty(C, f[i + 1], 1).uniform_(-0.5, 0.5)
            self.register_parameter(f"_bias{i}", nn.Parameter(b))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i}", nn.Parameter(torch.zeros(C, f[i + 1], 1)))
        self.quantiles = nn.Parameter(
            torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(C, 1, 1))
        t = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.tensor([-t, 0.0, t]))
        self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.noise_override = None  # test hook: shared U(-.5,.5) sample, shape of the input

    def _medians(self):
        return self.quantiles[:, :, 1:2]

    def _logits_cumulative(self, v, stop_gradient=False):
        for i in range(len(self.filters) + 1):
            m = getattr(self, f"_matrix{i}")
            b = getattr(self, f"_bias{i}")
            if stop_gradient:
                m, b = m.detach(), b.detach()
            v = torch.matmul(F.softplus(m), v) + b
            if i < len(self.filters):
                fa = getattr(self, f"_factor{i}")
                if stop_gradient:
                    fa = fa.detach()
                v = v + torch.tanh(fa) * torch.tanh(v)
        return v

    def _likelihood(self, v):
        lower = self._logits_cumulative(v - 0.5)
        upper = self._logits_cumulative(v + 0.5)
        sign = -torch.sign(lower + upper).detach()
        return torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))

    def loss(self):
        logits = self._logits_cumulative(self.quantiles, stop_gradient=True)
        return torch.abs(logits - self.target).sum()

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        perm = list(range(x.dim()))
        perm[0], perm[1] = 1, 0
        xp = x.permute(*perm).contiguous()
        shape = xp.shape
        v = xp.reshape(shape[0], 1, -1)
        if training:
            nz = None
            if self.noise_override is not None:
                nz = self.noise_override.permute(*perm).reshape(shape[0], 1, -1)
            out = quantize(v, "noise", noise=nz)
        else:
            out = quantize(v, "dequantize", self._medians())
        lik = self.likelihood_lower_bound(self._likelihood(out))
        out = out.reshape(shape).permute(*perm).contiguous()
        lik = lik.reshape(shape).permute(*perm).contiguous()
        return out, lik


class GaussianConditional(nn.Module):
    """Discretised Gaussian likelihood (A.4; reference pin visual_distribution.py:85-100)."""

    def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9, likelihood_bound=1e-9):
        super().__init__()
        self.lower_bound_scale = LowerBound(scale_bound)
        self.likelihood_lower_bound = LowerBound(likelihood_bound)
        self.noise_override = None

    @staticmethod
    def quantize(x, mode, means=None, noise=None):
        return quantize(x, mode, means, noise)

    @staticmethod
    def _standardized_cumulative(t):
        return 0.5 * torch.erfc(-(2 ** -0.5) * t)

    def _likelihood(self, v, scales, means=None):
        if means is not None:
            v = v - means
        scales = self.lower_bound_scale(scales)
        v = torch.abs(v)
        upper = self._standardized_cumulative((0.5 - v) / scales)
        lower = self._standardized_cumulative((-0.5 - v) / scales)
        return upper - lower

    def forward(self, x, scales, means=None, training=None):
        if training is None:
            training = self.training
        if training:
            out = quantize(x, "noise", means, self.noise_override)
        else:
            out = quantize(x, "dequantize", means)
        lik = self.likelihood_lower_bound(self._likelihood(out, scales, means))
        return out, lik
