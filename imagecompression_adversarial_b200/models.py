"""The reference's operator surface on the sm_100a kernels.

``init_model(MODEL, quality, metric, pretrained)`` (anchors/model.py:60-78) returns an ``nn.Module`` with
``forward(x) -> {"x_hat", "likelihoods": {"y"[, "z"]}}`` and the attributes the reference's host code
touches: ``g_a``, ``g_s`` (iterable, Sequential-like: anchors/utils.py:136,141), ``h_a``, ``h_s``,
``entropy_bottleneck``, ``gaussian_conditional`` (+ ``.quantize``), ``context_prediction``,
``entropy_parameters``, ``aux_loss()``; parameter names follow CompressAI so ``state_dict`` round-trips
(train.py:244-247) and only ``*.quantiles`` are auxiliary parameters (coder.py:57-67).

All tensors are fp32 CUDA; activations are kept ``channels_last``.  No module here has a CPU path.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import entropy_coding as ec
from . import functional as Fn
from . import ops
from . import precision
from .program import StackProgram, parse_stack

# compressai.zoo.image.cfgs
ZOO = {
    "factorized": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "hyper": {q: (128, 192) if q <= 5 else (192, 320) for q in range(1, 9)},
    "context": {q: (192, 192) if q <= 4 else (192, 320) for q in range(1, 9)},
    "cheng2020": {q: (128,) if q <= 3 else (192,) for q in range(1, 7)},
    "cheng2020_attn": {q: (128,) if q <= 3 else (192,) for q in range(1, 7)},
    "debug": {q: (3, 192) for q in range(1, 9)},      # anchors/model.py:61-68: ae_onelayer(N=3, M=192) at every quality
}


# Parameter gradients are produced only inside a full train-mode ``net(x)`` forward (the codec update, train.py:353-359)
# or when a module's ``param_grads`` is set by hand.  The attack calls ``net.g_a`` / ``net.g_s`` directly and only steps
# the perturbation (attack_rd.py:546-548): there the reference computes weight gradients and throws them away
# (train.py:357-358 zeroes them), so this path does not compute them at all.
_PARAM_GRADS = [False]


class _param_grads_on:
    def __init__(self, on):
        self.on = on

    def __enter__(self):
        self.prev = _PARAM_GRADS[0]
        _PARAM_GRADS[0] = self.prev or self.on

    def __exit__(self, *exc):
        _PARAM_GRADS[0] = self.prev


class _ContractionModule(nn.Module):
    transposed = False

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=2):
        super().__init__()
        self.in_channels, self.out_channels = in_channels, out_channels
        self.ksize, self.stride = kernel_size, stride
        shape = ((in_channels, out_channels) if self.transposed else (out_channels, in_channels)) + (kernel_size,) * 2
        w = torch.empty(*shape)
        nn.init.kaiming_normal_(w)  # CompressAI 1.1.x CompressionModel._initialize_weights
        self.weight = nn.Parameter(w)
        self.bias = nn.Parameter(torch.zeros(out_channels))
        self.param_grads = True  # the attack turns weight gradients off (they are never read there)

    def forward(self, x):
        return Fn.Contraction.apply(x, self.weight, self.bias, self.ksize, self.stride, self.transposed, L.ACT_NONE,
                                    self.param_grads)   # weight gradients only if the weight requires grad (autograd)

    def extra_repr(self):
        return f"{self.in_channels}, {self.out_channels}, kernel_size={self.ksize}, stride={self.stride}"


class Conv2d(_ContractionModule):
    """nn.Conv2d(i, o, k, stride=s, padding=k//2)  (anchors/utils.py:112-119)."""
    transposed = False


class ConvTranspose2d(_ContractionModule):
    """nn.ConvTranspose2d(i, o, k, stride=s, output_padding=s-1, padding=k//2)  (anchors/utils.py:122-130)."""
    transposed = True


def conv(cin, cout, kernel_size=5, stride=2):
    return Conv2d(cin, cout, kernel_size, stride)


def conv3x3(cin, cout, stride=1):
    return Conv2d(cin, cout, 3, stride)


def conv1x1(cin, cout, stride=1):
    return Conv2d(cin, cout, 1, stride)


class PixelShuffle(nn.Module):
    def __init__(self, r):
        super().__init__()
        self.r = int(r)

    def forward(self, x):
        return Fn.PixelShuffleFn.apply(x, self.r)


def subpel_conv3x3(cin, cout, r=1):
    """compressai.layers.subpel_conv3x3: conv3x3 to cout*r*r channels + PixelShuffle(r)."""
    return nn.Sequential(Conv2d(cin, cout * r * r, 3, 1), PixelShuffle(r))


class MaskedConv2d(Conv2d):
    """compressai.layers.MaskedConv2d, type-A mask (mbt2018 ``context_prediction``, call site anchors/model.py:103):
    the weight is masked in place at every forward, as CompressAI does."""

    def __init__(self, in_channels, out_channels, kernel_size=5, stride=1, mask_type="A"):
        super().__init__(in_channels, out_channels, kernel_size, stride)
        mask = torch.ones_like(self.weight.data)
        k = kernel_size
        mask[:, :, k // 2, k // 2 + (mask_type == "B"):] = 0
        mask[:, :, k // 2 + 1:] = 0
        self.register_buffer("mask", mask)

    def forward(self, x):
        with torch.no_grad():
            self.weight.mul_(self.mask)
        return super().forward(x)


def deconv(cin, cout, kernel_size=5, stride=2):
    return ConvTranspose2d(cin, cout, kernel_size, stride)


class _ConstStateKeys:
    """CompressAI keeps a few constants as persistent 1-element buffers, so its checkpoints (and checkpoints the
    reference's train.py wrote, train.py:443-454) carry keys such as ``g_a.1.beta_reparam.pedestal``,
    ``g_a.1.beta_reparam.lower_bound.bound``, ``entropy_bottleneck.likelihood_lower_bound.bound``,
    ``gaussian_conditional.lower_bound_scale.bound``.  Here those constants are plain floats (kernel arguments); this
    mixin EMITS them in ``state_dict()`` under CompressAI's names and, on load, accepts them (the checkpoint's value is
    adopted), so the strict ``net.load_state_dict(checkpoint["state_dict"])`` of coder.py:107 works on real checkpoints
    and a saved one loads back into CompressAI.  They are optional on load (an oracle / plain state dict has none)."""

    _const_keys = {}   # state-dict key (relative) -> attribute name holding the float

    def _save_to_state_dict(self, destination, prefix, keep_vars):
        super()._save_to_state_dict(destination, prefix, keep_vars)
        for key, attr in self._const_keys.items():
            destination[prefix + key] = torch.tensor([float(getattr(self, attr))], dtype=torch.float32)

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for key, attr in self._const_keys.items():
            v = state_dict.pop(prefix + key, None)
            if v is not None:
                setattr(self, attr, float(v.reshape(-1)[0]))
        super()._load_from_state_dict(state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs)


class NonNegativeParametrizer(_ConstStateKeys, nn.Module):
    """compressai.ops.NonNegativeParametrizer; exposes ``bound``/``pedestal`` (attack_rd.py:294-295 calls it)."""
    _const_keys = {"pedestal": "pedestal", "lower_bound.bound": "bound"}

    def __init__(self, minimum=0.0, reparam_offset=2 ** -18):
        super().__init__()
        self.pedestal = float(reparam_offset) ** 2
        self.bound = (float(minimum) + self.pedestal) ** 0.5

    def init(self, x):
        return torch.sqrt(torch.clamp(x + self.pedestal, min=self.pedestal))

    def forward(self, x):
        return ops.gdn_reparam(x, self.bound, self.pedestal)


class GDN(nn.Module):
    """compressai.layers.GDN (utils/ops.py:58-97): y = x * (beta + gamma x^2)^(-+1/2)."""

    def __init__(self, C, inverse=False, beta_min=1e-6, gamma_init=0.1):
        super().__init__()
        self.inverse = bool(inverse)
        self.beta_reparam = NonNegativeParametrizer(minimum=beta_min)
        self.gamma_reparam = NonNegativeParametrizer()
        self.beta = nn.Parameter(self.beta_reparam.init(torch.ones(C)))
        self.gamma = nn.Parameter(self.gamma_reparam.init(gamma_init * torch.eye(C)))

    def effective_parameters(self, round_tf32=False):
        """(beta_eff [C], gamma_eff [C,C], gamma_eff^T) after the non-negative reparametrisation; the gamma
        matrices optionally rounded to TF32 (they are operands of the tensor-path normalisation GEMM)."""
        be = ops.gdn_reparam(self.beta, self.beta_reparam.bound, self.beta_reparam.pedestal)
        ga = ops.gdn_reparam(self.gamma, self.gamma_reparam.bound, self.gamma_reparam.pedestal,
                             round_tf32=round_tf32)
        gaT = ops.gdn_reparam(self.gamma, self.gamma_reparam.bound, self.gamma_reparam.pedestal, transpose=True,
                              round_tf32=round_tf32)
        return be, ga, gaT

    def forward(self, x):
        if torch.is_grad_enabled() and (self.beta.requires_grad or self.gamma.requires_grad) and _PARAM_GRADS[0]:
            return Fn.GdnTrainFn.apply(x, self.beta, self.gamma, self.inverse, self.beta_reparam.bound,
                                       self.beta_reparam.pedestal, self.gamma_reparam.bound, self.gamma_reparam.pedestal)
        be, ga, _ = self.effective_parameters(round_tf32=not precision.split())
        if Fn._REC is not None:
            Fn._REC.current_gdn = self          # a trace records which module owns this normalisation (tape.py)
        return Fn.GdnFn.apply(x, be, ga, self.inverse)


class Activation(nn.Module):
    def __init__(self, act, inplace=True):
        super().__init__()
        self.act = act

    def forward(self, x):
        return Fn.ActFn.apply(x, self.act)


def ReLU(inplace=True):
    return Activation(L.ACT_RELU)


def LeakyReLU(inplace=True):
    return Activation(L.ACT_LEAKY)


class ResidualBlockWithStride(nn.Module):
    """compressai.layers.ResidualBlockWithStride (cheng2020_anchor g_a)."""

    def __init__(self, cin, cout, stride=2):
        super().__init__()
        self.conv1 = conv3x3(cin, cout, stride)
        self.leaky_relu = LeakyReLU()
        self.conv2 = conv3x3(cout, cout)
        self.gdn = GDN(cout)
        self.skip = conv1x1(cin, cout, stride) if (stride != 1 or cin != cout) else None

    def forward(self, x):
        out = self.gdn(self.conv2(self.leaky_relu(self.conv1(x))))
        return Fn.AddFn.apply(out, x if self.skip is None else self.skip(x))


class ResidualBlockUpsample(nn.Module):
    """compressai.layers.ResidualBlockUpsample (cheng2020_anchor g_s)."""

    def __init__(self, cin, cout, upsample=2):
        super().__init__()
        self.subpel_conv = subpel_conv3x3(cin, cout, upsample)
        self.leaky_relu = LeakyReLU()
        self.conv = conv3x3(cout, cout)
        self.igdn = GDN(cout, inverse=True)
        self.upsample = subpel_conv3x3(cin, cout, upsample)

    def forward(self, x):
        out = self.igdn(self.conv(self.leaky_relu(self.subpel_conv(x))))
        return Fn.AddFn.apply(out, self.upsample(x))


class ResidualBlock(nn.Module):
    """compressai.layers.ResidualBlock."""

    def __init__(self, cin, cout):
        super().__init__()
        self.conv1 = conv3x3(cin, cout)
        self.leaky_relu = LeakyReLU()
        self.conv2 = conv3x3(cout, cout)
        self.skip = conv1x1(cin, cout) if cin != cout else None

    def forward(self, x):
        out = self.leaky_relu(self.conv2(self.leaky_relu(self.conv1(x))))
        return Fn.AddFn.apply(out, x if self.skip is None else self.skip(x))


_INFER_PROGRAMS = {}


def _inference_program(stack, n, h, w, device):
    """Cached forward-only StackProgram of ``stack`` for this input shape (at most a few live entries per stack)."""
    import weakref
    key = (id(stack), n, h, w, str(device), precision.get())
    hit = _INFER_PROGRAMS.get(key)
    if hit is not None and hit[0]() is stack:
        return hit[1]
    for k in [k for k, v in _INFER_PROGRAMS.items() if v[0]() is None or (k[0] == id(stack) and k != key)]:
        del _INFER_PROGRAMS[k]          # dead stacks, and other shapes of this stack: keep one program per stack
    prog = StackProgram(parse_stack(stack), n, h, w, device, need_grad=False)
    _INFER_PROGRAMS[key] = (weakref.ref(stack), prog)
    return prog


_GRAD_PROGRAMS = {}


def _grad_program(stack, n, h, w, device):
    """Cached forward + input-gradient StackProgram of ``stack`` for this input shape: the unmodified reference loop calls
    ``net.g_a(im_in)`` / ``net.g_s(y)`` once per iteration (attack_rd.py:344,349), and building a program (buffers,
    weight packing, TMA plans) per call costs more than running it.  One program per stack is kept."""
    import weakref
    key = (id(stack), n, h, w, str(device), precision.get())
    hit = _GRAD_PROGRAMS.get(key)
    if hit is not None and hit[0]() is stack:
        return hit[1]
    for k in [k for k, v in _GRAD_PROGRAMS.items() if v[0]() is None or (k[0] == id(stack) and k != key)]:
        del _GRAD_PROGRAMS[k]
    prog = StackProgram(parse_stack(stack), n, h, w, device, need_grad=True)
    prog.version = 0
    _GRAD_PROGRAMS[key] = (weakref.ref(stack), prog)
    return prog


class _StackFn(torch.autograd.Function):
    """Whole g_a / g_s stack as ONE autograd node running a fused StackProgram."""

    @staticmethod
    def forward(ctx, x, stack):
        xn = Fn.to_nhwc(x).contiguous()
        n, h, w, _ = xn.shape
        need_grad = x.requires_grad
        if not need_grad:
            # inference (clean pass, final eval: attack_rd.py:406, self_ensemble.py:190): one cached launch program per
            # (stack, shape) -- plans, tensor maps and scratch buffers are reused across calls; the result is copied out
            prog = _inference_program(stack, n, h, w, x.device)
            prog.x_in.copy_(xn)
            prog.refresh_parameters()
            ctx.prog = None
            return Fn.to_nchw(prog.forward().clone())
        # grad-enabled call: the cached program of this (stack, shape); its buffers hold the saved activations until the
        # matching backward.  The input is kept so that a backward that arrives after the stack has been run again (two
        # forwards, then two backwards) can replay its forward first.
        prog = _grad_program(stack, n, h, w, x.device)
        prog.x_in.copy_(xn)
        prog.refresh_parameters()
        prog.version += 1
        ctx.prog, ctx.version = prog, prog.version
        ctx.save_for_backward(xn)
        return Fn.to_nchw(prog.forward().clone())

    @staticmethod
    def backward(ctx, g):
        prog = ctx.prog
        if prog.version != ctx.version:        # the stack ran forward again since: restore this call's activations
            (xn,) = ctx.saved_tensors
            prog.x_in.copy_(xn)
            prog.forward()
            prog.version += 1
            ctx.version = prog.version
        prog.g_out.copy_(Fn.to_nhwc(g))
        return Fn.to_nchw(prog.backward().clone()), None


class CodecStack(nn.Sequential):
    """``nn.Sequential`` of Conv2d/ConvTranspose2d/GDN whose forward runs the fused launch program.

    Iterating ``_modules`` still yields the individual layers (anchors/utils.py:136,141).  Parameter
    gradients are produced only when ``param_grads`` is set (the attack never reads them:
    attack_rd.py:546-548 steps only the perturbation; train.py:357-358 zeroes them before use).
    """
    param_grads = False

    def forward(self, x):
        if (self.param_grads or _PARAM_GRADS[0]) and torch.is_grad_enabled() and \
                any(p.requires_grad for p in self.parameters()):
            with _param_grads_on(True):
                for m in self:   # layer-by-layer autograd path, with weight and GDN-parameter gradients
                    x = m(x)
            return x
        return _StackFn.apply(x, self)


# ---------------------------------------------------------------------------------- entropy models
class LowerBound(nn.Module):
    def __init__(self, bound):
        super().__init__()
        self.bound = float(bound)

    def forward(self, x):
        return Fn.Low_bound.apply(x, self.bound)


_CDF_BUFFERS = ("_offset", "_quantized_cdf", "_cdf_length")


class _ZooStateDict(_ConstStateKeys):
    """Checkpoint compatibility of the entropy models (reference: coder.py:104-116,
    anchors/balle.py:57-72, anchors/utils.py:46-109; key names SURVEY.md A.7).

    * The entropy-coder tables a CompressAI checkpoint carries (``_quantized_cdf``, ``_offset``, ``_cdf_length``, and
      ``scale_table`` for GaussianConditional) are registered EMPTY like CompressAI does, so the reference's
      ``update_registered_buffers`` can resize them and ``state_dict()`` round-trips them.  On load they take the
      checkpoint's sizes ("resize" policy) and are optional (a plain / oracle state dict has none).  The likelihood
      kernels never read them: bpp is estimated from the likelihoods (attack_rd.py:419); entropy coding itself is out
      of scope (section 8f rank 4).
    * CompressAI >= 1.2 stores the EntropyBottleneck chain as ParameterLists (``matrices.0`` ...); the reference's era
      and this package use ``_matrix0`` ...: both spellings load."""

    _zoo_buffers = _CDF_BUFFERS

    def _register_zoo_buffers(self):
        for name in _CDF_BUFFERS:
            self.register_buffer(name, torch.zeros(0, dtype=torch.int32))

    def _load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys, error_msgs):
        for new, old in (("matrices.", "_matrix"), ("biases.", "_bias"), ("factors.", "_factor")):
            for key in [k for k in state_dict if k.startswith(prefix + new)]:
                state_dict[prefix + old + key[len(prefix + new):]] = state_dict.pop(key)
        for key, attr in self._const_keys.items():      # CompressAI's constant buffers (see _ConstStateKeys)
            v = state_dict.pop(prefix + key, None)
            if v is not None:
                setattr(self, attr, float(v.reshape(-1)[0]))
        for name in self._zoo_buffers:
            key, buf = prefix + name, getattr(self, name)
            if key in state_dict:
                if tuple(buf.shape) != tuple(state_dict[key].shape):
                    setattr(self, name, torch.zeros(state_dict[key].shape, dtype=buf.dtype, device=buf.device))
            else:
                state_dict[key] = buf
        nn.Module._load_from_state_dict(self, state_dict, prefix, local_metadata, strict, missing_keys, unexpected_keys,
                                        error_msgs)


class EntropyBottleneck(_ZooStateDict, nn.Module):
    """compressai EntropyBottleneck, forward on the fused likelihood kernel."""
    _const_keys = {"likelihood_lower_bound.bound": "likelihood_bound"}

    def __init__(self, channels, tail_mass=1e-9, init_scale=10.0, filters=(3, 3, 3, 3)):
        super().__init__()
        assert tuple(filters) == (3, 3, 3, 3), "kernel is specialised to the 1-3-3-3-3-1 chain"
        self.channels, self.filters = int(channels), tuple(filters)
        self.init_scale, self.tail_mass = float(init_scale), float(tail_mass)
        f = (1,) + self.filters + (1,)
        scale = self.init_scale ** (1 / (len(self.filters) + 1))
        C = self.channels
        for i in range(len(self.filters) + 1):
            init = math.log(math.expm1(1 / scale / f[i + 1]))
            self.register_parameter(f"_matrix{i}", nn.Parameter(torch.full((C, f[i + 1], f[i]), init)))
            self.register_parameter(f"_bias{i}", nn.Parameter(torch.empty(C, f[i + 1], 1).uniform_(-0.5, 0.5)))
            if i < len(self.filters):
                self.register_parameter(f"_factor{i}", nn.Parameter(torch.zeros(C, f[i + 1], 1)))
        self.quantiles = nn.Parameter(torch.tensor([-self.init_scale, 0.0, self.init_scale]).repeat(C, 1, 1))
        t = math.log(2 / self.tail_mass - 1)
        self.register_buffer("target", torch.tensor([-t, 0.0, t]))
        self._register_zoo_buffers()
        self.likelihood_bound = 1e-9
        self.noise_override = None  # test hook: U(-.5,.5) sample shared with the oracle (NCHW)
        self.last_bits = None

    def _table(self):
        ms = [getattr(self, f"_matrix{i}") for i in range(5)]
        bs = [getattr(self, f"_bias{i}") for i in range(5)]
        fs = [getattr(self, f"_factor{i}") for i in range(4)]
        return ops.eb_prepare(ms, bs, fs, self.channels)

    def loss(self):
        """Auxiliary loss on the quantiles (parameter-only math, 3*C elements)."""
        v = self.quantiles
        for i in range(5):
            m = getattr(self, f"_matrix{i}").detach()
            b = getattr(self, f"_bias{i}").detach()
            v = torch.matmul(torch.nn.functional.softplus(m), v) + b
            if i < 4:
                fa = getattr(self, f"_factor{i}").detach()
                v = v + torch.tanh(fa) * torch.tanh(v)
        return torch.abs(v - self.target).sum()

    # ---- entropy coding (compressai EntropyBottleneck.update / compress / decompress)
    lanes = ec.DEFAULT_LANES          # independent rANS streams per image (1 = compressai's single stream)

    def _logits_cumulative_host(self, v):
        for i in range(5):
            m = getattr(self, f"_matrix{i}").detach().float().cpu()
            b = getattr(self, f"_bias{i}").detach().float().cpu()
            v = torch.matmul(torch.nn.functional.softplus(m), v) + b
            if i < 4:
                fa = getattr(self, f"_factor{i}").detach().float().cpu()
                v = v + torch.tanh(fa) * torch.tanh(v)
        return v

    def update(self, force=False):
        """Builds ``_quantized_cdf`` / ``_cdf_length`` / ``_offset`` from the learned density (host, parameter-sized)."""
        if self._offset.numel() > 0 and not force:
            return False
        dev = self.quantiles.device
        q = self.quantiles.detach().float().cpu()
        medians = q[:, 0, 1]
        minima = torch.ceil(medians - q[:, 0, 0]).int().clamp(min=0)
        maxima = torch.ceil(q[:, 0, 2] - medians).int().clamp(min=0)
        pmf_start = medians - minima
        pmf_length = maxima + minima + 1
        max_length = int(pmf_length.max())
        samples = torch.arange(max_length)[None, :] + pmf_start[:, None, None]
        lower = self._logits_cumulative_host(samples - 0.5)
        upper = self._logits_cumulative_host(samples + 0.5)
        sign = -torch.sign(lower + upper)
        pmf = torch.abs(torch.sigmoid(sign * upper) - torch.sigmoid(sign * lower))[:, 0, :]
        tail_mass = torch.sigmoid(lower[:, 0, :1]) + torch.sigmoid(-upper[:, 0, -1:])
        self._quantized_cdf = ec.pmf_to_cdf(pmf, tail_mass, pmf_length, max_length).to(dev)
        self._cdf_length = (pmf_length + 2).int().to(dev)
        self._offset = (-minima).int().to(dev)
        return True

    def _coder_tables(self):
        if self._offset.numel() == 0:
            raise L.IcadvError("Uninitialized CDFs. Run update() first")
        return ec.CoderTables(self._quantized_cdf, self._cdf_length, self._offset, self.quantiles.device)

    def _channel_indexes(self, n, h, w):
        c = self.channels
        return torch.arange(c, device=self.quantiles.device, dtype=torch.int32).expand(n, h, w, c).contiguous()

    def compress(self, x):
        """-> one byte string per image (symbol order (c, h, w) as compressai's)."""
        xn = Fn.to_nhwc(x.detach()).contiguous()
        med = self.quantiles.detach()[:, 0, 1]
        sym = ops.unary(xn - med, 3).int()
        return ec.rans_encode(sym, self._channel_indexes(*xn.shape[:3]), self._coder_tables(), mode=0, lanes=self.lanes)

    def decompress(self, strings, size):
        n, (h, w) = len(strings), size
        idx = self._channel_indexes(n, h, w)
        med = self.quantiles.detach()[:, 0, 1].expand(n, h, w, self.channels).contiguous()
        out = ec.rans_decode(strings, idx, self._coder_tables(), means=med, mode=0, lanes=self.lanes)
        return Fn.to_nchw(out)

    def forward(self, x, training=None):
        if training is None:
            training = self.training
        med = self.quantiles.detach()[:, 0, 1].contiguous()
        if training and torch.is_grad_enabled() and (x.requires_grad or self._matrix0.requires_grad):
            noise = Fn.to_nhwc(self.noise_override).contiguous() if self.noise_override is not None else None
            params = [getattr(self, f"_matrix{i}") for i in range(5)] + [getattr(self, f"_bias{i}") for i in range(5)] + \
                     [getattr(self, f"_factor{i}") for i in range(4)]
            return Fn.EbTrainFn.apply(x, noise, med, self.likelihood_bound, *params)
        xn = Fn.to_nhwc(x).contiguous()
        noise = None
        if training:
            noise = (Fn.to_nhwc(self.noise_override).contiguous() if self.noise_override is not None
                     else ops.uniform_noise_like(xn))
        x_hat, lik, bits = ops.eb_forward(xn, self._table(), med, training=training, noise=noise,
                                          lik_bound=self.likelihood_bound)
        self.last_bits = bits
        return Fn.to_nchw(x_hat), Fn.to_nchw(lik)


class GaussianConditional(_ZooStateDict, nn.Module):
    """compressai GaussianConditional (A.4) on the fused erfc likelihood kernel."""
    _zoo_buffers = _CDF_BUFFERS + ("scale_table",)
    _const_keys = {"likelihood_lower_bound.bound": "likelihood_bound", "lower_bound_scale.bound": "scale_bound"}

    def __init__(self, scale_table=None, scale_bound=0.11, tail_mass=1e-9):
        super().__init__()
        self.scale_bound, self.likelihood_bound = float(scale_bound), 1e-9
        self._register_zoo_buffers()
        self.register_buffer("scale_table", torch.tensor(sorted(float(v) for v in scale_table)) if scale_table is not None
                             else torch.zeros(0))
        self.noise_override = None
        self.last_bits = None

    # ---- entropy coding (compressai GaussianConditional.update_scale_table / update / build_indexes / compress /
    #      decompress)
    lanes = ec.DEFAULT_LANES
    tail_mass = 1e-9

    def update_scale_table(self, scale_table, force=False):
        if self._offset.numel() > 0 and not force:
            return False
        dev = self._offset.device
        self.scale_table = torch.tensor(sorted(float(v) for v in scale_table), device=dev)
        self.update()
        return True

    def update(self):
        from scipy.stats import norm
        dev = self._offset.device
        st = self.scale_table.detach().float().cpu()
        multiplier = -norm.ppf(self.tail_mass / 2)
        pmf_center = torch.ceil(st * multiplier).int()
        pmf_length = 2 * pmf_center + 1
        max_length = int(pmf_length.max())
        samples = torch.abs(torch.arange(max_length).int() - pmf_center[:, None]).float()
        sc = st[:, None]
        cum = lambda t: 0.5 * torch.erfc(-(2 ** -0.5) * t)
        upper, lower = cum((0.5 - samples) / sc), cum((-0.5 - samples) / sc)
        self._quantized_cdf = ec.pmf_to_cdf(upper - lower, 2 * lower[:, :1], pmf_length, max_length).to(dev)
        self._cdf_length = (pmf_length + 2).int().to(dev)
        self._offset = (-pmf_center).int().to(dev)

    def _coder_tables(self):
        if self._offset.numel() == 0:
            raise L.IcadvError("Uninitialized CDFs. Run update() first")
        return ec.CoderTables(self._quantized_cdf, self._cdf_length, self._offset, self._offset.device)

    def build_indexes(self, scales):
        return ec.build_indexes(scales.detach(), self.scale_table, self.scale_bound)

    def compress(self, inputs, indexes, means=None):
        sym = Fn.to_nhwc(self.quantize(inputs.detach(), "symbols", means)).contiguous()
        return ec.rans_encode(sym, Fn.to_nhwc(indexes).contiguous(), self._coder_tables(), mode=0, lanes=self.lanes)

    def decompress(self, strings, indexes, dtype=torch.float, means=None):
        m = Fn.to_nhwc(means.detach()).contiguous() if means is not None else None
        out = ec.rans_decode(strings, Fn.to_nhwc(indexes).contiguous(), self._coder_tables(), means=m, mode=0,
                             lanes=self.lanes)
        return Fn.to_nchw(out)

    def quantize(self, inputs, mode, means=None):
        """compressai ``quantize`` (call site anchors/model.py:102)."""
        x = inputs.contiguous(memory_format=torch.channels_last)
        if mode == "noise":
            nz = (self.noise_override.contiguous(memory_format=torch.channels_last)
                  if self.noise_override is not None else ops.uniform_noise_like(x))
            if torch.is_grad_enabled() and inputs.requires_grad:
                # compressai: ``inputs + noise`` is differentiable (identity to the inputs, anchors/model.py:102), so the
                # distortion gradient through g_s(y_hat) and context_prediction(y_hat) reaches g_a in training
                return Fn.AddFn.apply(x, nz)
            return ops.unary(x, 4, nz)
        if means is not None:
            m = means.contiguous(memory_format=torch.channels_last)
            r = ops.unary(x - m, 3)                      # round(x - mean): the integer symbols
            if mode == "symbols":
                return r.int()
            assert mode == "dequantize", mode
            return ops.unary(r, 4, m)
        out = ops.unary(x, 3)
        if mode == "dequantize":
            return out
        assert mode == "symbols", mode
        return out.int()

    def forward(self, inputs, scales, means=None, training=None):
        if training is None:
            training = self.training
        if training and torch.is_grad_enabled() and (inputs.requires_grad or scales.requires_grad):
            noise = Fn.to_nhwc(self.noise_override).contiguous() if self.noise_override is not None else None
            return Fn.GcTrainFn.apply(inputs, scales, means, noise, self.scale_bound, self.likelihood_bound)
        y = Fn.to_nhwc(inputs).contiguous()
        s = Fn.to_nhwc(scales).contiguous()
        m = Fn.to_nhwc(means).contiguous() if means is not None else None
        noise = None
        if training:
            noise = (Fn.to_nhwc(self.noise_override).contiguous() if self.noise_override is not None
                     else ops.uniform_noise_like(y))
        y_hat, lik, bits = ops.gc_forward(y, s, m, training=training, noise=noise, scale_bound=self.scale_bound,
                                          lik_bound=self.likelihood_bound)
        self.last_bits = bits
        return Fn.to_nchw(y_hat), Fn.to_nchw(lik)


# ---------------------------------------------------------------------------------- codecs
class CompressionModel(nn.Module):
    def __init__(self, entropy_bottleneck_channels):
        super().__init__()
        self.entropy_bottleneck = EntropyBottleneck(entropy_bottleneck_channels)

    def aux_loss(self):
        return sum(m.loss() for m in self.modules() if isinstance(m, EntropyBottleneck))

    def to(self, *args, **kwargs):
        return super().to(*args, **kwargs)

    def update(self, scale_table=None, force=False):
        """compressai ``CompressionModel.update`` (+ the scale table of the conditional model where there is one):
        builds the entropy coder's CDF tables; call before ``compress`` / ``decompress`` (InvCompress/train.py:452)."""
        updated = False
        gc = getattr(self, "gaussian_conditional", None)
        if gc is not None:
            updated |= gc.update_scale_table(ec.get_scale_table() if scale_table is None else scale_table, force=force)
        for m in self.modules():
            if isinstance(m, EntropyBottleneck):
                updated |= m.update(force=force)
        return updated


def _g_a(N, M):
    return CodecStack(conv(3, N), GDN(N), conv(N, N), GDN(N), conv(N, N), GDN(N), conv(N, M))


def _g_s(N, M):
    return CodecStack(deconv(M, N), GDN(N, inverse=True), deconv(N, N), GDN(N, inverse=True), deconv(N, N),
                      GDN(N, inverse=True), deconv(N, 3))


class FactorizedPrior(CompressionModel):
    def __init__(self, N, M):
        super().__init__(M)
        self.g_a, self.g_s = _g_a(N, M), _g_s(N, M)
        self.N, self.M = N, M

    def forward(self, x):
        with _param_grads_on(self.training):
            y = self.g_a(x)
            y_hat, y_lik = self.entropy_bottleneck(y)   # anchors/model.py:87-89
            return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik}}


    @torch.no_grad()
    def compress(self, x):
        """compressai FactorizedPrior.compress."""
        y = self.g_a(x)
        return {"strings": [self.entropy_bottleneck.compress(y)], "shape": y.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 1
        y_hat = self.entropy_bottleneck.decompress(strings[0], shape)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}


class ScaleHyperprior(CompressionModel):
    def __init__(self, N, M):
        super().__init__(N)
        self.g_a, self.g_s = _g_a(N, M), _g_s(N, M)
        self.h_a = nn.Sequential(conv(M, N, 3, 1), ReLU(), conv(N, N), ReLU(), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, N), ReLU(), deconv(N, N), ReLU(), conv(N, M, 3, 1), ReLU())
        self.gaussian_conditional = GaussianConditional()
        self.N, self.M = N, M

    def forward(self, x):
        with _param_grads_on(self.training):
            y = self.g_a(x)
            z = self.h_a(Fn.ActFn.apply(y, L.ACT_ABS))  # anchors/model.py:92, anchors/balle.py:38
            z_hat, z_lik = self.entropy_bottleneck(z)
            scales_hat = self.h_s(z_hat)
            y_hat, y_lik = self.gaussian_conditional(y, scales_hat)
            return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


    @torch.no_grad()
    def compress(self, x):
        """compressai ScaleHyperprior.compress: z string, then y coded with the scale indexes h_s(z_hat) gives."""
        y = self.g_a(x)
        z = self.h_a(Fn.ActFn.apply(y, L.ACT_ABS))
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        indexes = self.gaussian_conditional.build_indexes(self.h_s(z_hat))
        y_strings = self.gaussian_conditional.compress(y, indexes)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        indexes = self.gaussian_conditional.build_indexes(self.h_s(z_hat))
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, z_hat.dtype)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}


class MeanScaleHyperprior(ScaleHyperprior):
    """compressai MeanScaleHyperprior (mbt2018_mean): the base class of the reference's ``debug`` model
    (anchors/model.py:9-35), whose own forward (:22-35) is what runs there; this forward is CompressAI's."""

    def __init__(self, N, M, **kwargs):
        super().__init__(N, M)
        self.h_a = nn.Sequential(conv(M, N, 3, 1), LeakyReLU(), conv(N, N), LeakyReLU(), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, M), LeakyReLU(), deconv(M, M * 3 // 2), LeakyReLU(),
                                 conv(M * 3 // 2, M * 2, 3, 1))

    def forward(self, x):
        with _param_grads_on(self.training):
            y = self.g_a(x)
            z = self.h_a(y)
            z_hat, z_lik = self.entropy_bottleneck(z)
            gp = self.h_s(z_hat)
            half = gp.shape[1] // 2
            scales_hat, means_hat = Fn.NarrowFn.apply(gp, 0, half), Fn.NarrowFn.apply(gp, half, half)
            y_hat, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
            return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


    @torch.no_grad()
    def compress(self, x):
        y = self.g_a(x)
        z = self.h_a(y)
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        gp = self.h_s(z_hat)
        half = gp.shape[1] // 2
        scales_hat, means_hat = gp[:, :half], gp[:, half:]
        indexes = self.gaussian_conditional.build_indexes(scales_hat.contiguous(memory_format=torch.channels_last))
        y_strings = self.gaussian_conditional.compress(y, indexes, means=means_hat)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape):
        assert isinstance(strings, list) and len(strings) == 2
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        gp = self.h_s(z_hat)
        half = gp.shape[1] // 2
        scales_hat, means_hat = gp[:, :half], gp[:, half:]
        indexes = self.gaussian_conditional.build_indexes(scales_hat.contiguous(memory_format=torch.channels_last))
        y_hat = self.gaussian_conditional.decompress(strings[0], indexes, means=means_hat)
        return {"x_hat": self.g_s(y_hat).clamp_(0, 1)}


class JointAutoregressiveHierarchicalPriors(CompressionModel):
    """compressai mbt2018 (widths pinned by InvCompress/ours.py:22-32); forward of anchors/model.py:96-108."""

    def __init__(self, N, M):
        super().__init__(N)
        self.g_a, self.g_s = _g_a(N, M), _g_s(N, M)
        self.h_a = nn.Sequential(conv(M, N, 3, 1), LeakyReLU(), conv(N, N), LeakyReLU(), conv(N, N))
        self.h_s = nn.Sequential(deconv(N, M), LeakyReLU(), deconv(M, M * 3 // 2), LeakyReLU(),
                                 conv(M * 3 // 2, M * 2, 3, 1))
        self.entropy_parameters = nn.Sequential(
            Conv2d(M * 12 // 3, M * 10 // 3, 1, 1), LeakyReLU(),
            Conv2d(M * 10 // 3, M * 8 // 3, 1, 1), LeakyReLU(),
            Conv2d(M * 8 // 3, M * 6 // 3, 1, 1))
        self.context_prediction = MaskedConv2d(M, 2 * M, kernel_size=5, stride=1)
        self.gaussian_conditional = GaussianConditional()
        self.N, self.M = N, M

    def forward(self, x):
        with _param_grads_on(self.training):
            return self._forward(x)

    def _forward(self, x):
        y = self.g_a(x)
        z = self.h_a(y)                                          # anchors/model.py:98 (no abs)
        z_hat, z_lik = self.entropy_bottleneck(z)
        params = self.h_s(z_hat)
        y_hat = self.gaussian_conditional.quantize(y, "noise" if self.training else "dequantize")   # :102
        ctx = self.context_prediction(y_hat)
        gp = self.entropy_parameters(Fn.CatFn.apply(params, ctx))                                  # :104
        half = gp.shape[1] // 2
        scales_hat, means_hat = Fn.NarrowFn.apply(gp, 0, half), Fn.NarrowFn.apply(gp, half, half)  # :105
        _, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
        return {"x_hat": self.g_s(y_hat), "likelihoods": {"y": y_lik, "z": z_lik}}


    # ---- entropy coding (compressai JointAutoregressiveHierarchicalPriors.compress / decompress with _compress_ar /
    #      _decompress_ar; model-level shape InvCompress/ours.py:100-175).  Every position's (scale, mean) depends on the
    #      latents already coded through the masked 5x5 context, in the encoder as in the decoder.  compressai walks
    #      the positions in raster order on the CPU (H*W sequential steps per image); here the positions that are
    #      mutually independent under the type-A mask form a wavefront (j + 3 i = const), so W + 3(H - 1) steps code
    #      a whole batch, each step = one gathered 1x1 contraction for the context model, the entropy-parameter
    #      network, a quantiser / index kernel and (decoder) one rANS step kernel.  ``wavefront=False`` keeps raster
    #      order; with ``gaussian_conditional.lanes = 1`` that string is the one compressai's coder emits.
    def _ar_pass(self, y, params, decoder, wavefront):
        """y: channels-last latent [N,H,W,M] (encoder) or None (decoder); params: channels-last h_s output.
        Returns (y_hat [N,H,W,M], symbols, indexes, order)."""
        gc = self.gaussian_conditional
        n, h, w, cp = params.shape
        m = self.context_prediction.in_channels
        dev = params.device
        hp, wp = h + 4, w + 4
        y_pad = torch.zeros(n, hp * wp, m, device=dev)
        with torch.no_grad():
            self.context_prediction.weight.mul_(self.context_prediction.mask)
        wt = self.context_prediction.weight.detach()
        live = [(dy, dx) for dy in range(5) for dx in range(5) if dy < 2 or (dy == 2 and dx < 2)]
        # gathered form of the masked conv: one 1x1 contraction over the 12 live taps, in-channel = tap * M + c
        w1 = torch.stack([wt[:, :, dy, dx] for dy, dx in live], dim=1).reshape(wt.shape[0], len(live) * m, 1, 1).contiguous()
        tap_off = torch.tensor([dy * wp + dx for dy, dx in live], device=dev)
        bias = self.context_prediction.bias.detach()
        steps = ec.ar_schedule(h, w, wavefront)
        params = params.reshape(n, h * w, cp)
        sym = torch.zeros(n, h * w, m, device=dev, dtype=torch.int32)
        idx = torch.zeros(n, h * w, m, device=dev, dtype=torch.int32)
        mean_full = torch.zeros(n, h * w, m, device=dev)
        out_full = torch.zeros(n, h * w, m, device=dev)
        if y is not None:
            if tuple(y.shape) != (n, h, w, m):
                raise L.IcadvError(f"autoregressive coding needs image sides that are multiples of 64 (latent "
                                   f"{tuple(y.shape[1:3])} vs hyper-decoder output {(h, w)}); pad the input as "
                                   "attack_TIC.py:95-104 does")
            y = y.reshape(n, h * w, m)
        for st in steps:
            st = st.to(dev)
            pos = st.long()
            base = (pos // w) * wp + (pos % w)
            r = n * pos.numel()
            rows = (r + 15) // 16 * 16
            taps = torch.zeros(rows, len(live) * m, device=dev)
            taps[:r] = y_pad[:, base[:, None] + tap_off[None, :]].reshape(r, -1)
            ctx = Fn.Contraction.apply(taps.view(1, rows // 16, 16, -1).permute(0, 3, 1, 2), w1, bias, 1, 1, False,
                                       L.ACT_NONE, False)
            feat = torch.zeros(rows, cp + ctx.shape[1], device=dev)
            feat[:r, :cp] = params[:, pos].reshape(r, cp)
            feat[:, cp:] = Fn.to_nhwc(ctx).reshape(rows, -1)
            gp = Fn.to_nhwc(self.entropy_parameters(feat.view(1, rows // 16, 16, -1).permute(0, 3, 1, 2)))
            gp = gp.reshape(rows, -1)[:r]
            scales, means = gp[:, :m].contiguous(), gp[:, m:].contiguous()
            ind = gc.build_indexes(scales).view(n, -1, m)
            if decoder is None:
                q = ops.unary(y[:, pos].reshape(r, m) - means, 3)
                sym[:, pos] = q.int().view(n, -1, m)
                idx[:, pos] = ind
                y_hat = q + means
            else:
                idx[:, pos] = ind
                mean_full[:, pos] = means.view(n, -1, m)
                decoder.step(idx, mean_full, out_full, st)
                y_hat = out_full[:, pos].reshape(r, m)
            y_pad[:, base + 2 * wp + 2] = y_hat.view(n, -1, m)
        y_hat = y_pad.view(n, hp, wp, m)[:, 2:-2, 2:-2].contiguous()
        return y_hat, sym.view(n, h, w, m), idx.view(n, h, w, m), torch.cat(steps)

    @torch.no_grad()
    def compress(self, x, wavefront=True):
        gc = self.gaussian_conditional
        y = self.g_a(x)
        z = self.h_a(y)
        z_strings = self.entropy_bottleneck.compress(z)
        z_hat = self.entropy_bottleneck.decompress(z_strings, z.size()[-2:])
        params = Fn.to_nhwc(self.h_s(z_hat)).contiguous()
        y_hat, sym, idx, order = self._ar_pass(Fn.to_nhwc(y).contiguous(), params, None, wavefront)
        y_strings = ec.rans_encode(sym, idx, gc._coder_tables(), mode=1, order=order, lanes=gc.lanes)
        self._coded_y_hat = y_hat          # what the decoder must reproduce (tests)
        return {"strings": [y_strings, z_strings], "shape": z.size()[-2:]}

    @torch.no_grad()
    def decompress(self, strings, shape, wavefront=True):
        assert isinstance(strings, list) and len(strings) == 2
        gc = self.gaussian_conditional
        z_hat = self.entropy_bottleneck.decompress(strings[1], shape)
        params = Fn.to_nhwc(self.h_s(z_hat)).contiguous()
        dec = ec.StepDecoder(strings[0], gc._coder_tables(), self.context_prediction.in_channels, params.device,
                             lanes=gc.lanes)
        y_hat, _, _, _ = self._ar_pass(None, params, dec, wavefront)
        self._decoded_y_hat = y_hat
        return {"x_hat": self.g_s(Fn.to_nchw(y_hat)).clamp_(0, 1)}


class Cheng2020Anchor(JointAutoregressiveHierarchicalPriors):
    """compressai cheng2020_anchor (anchors/model.py:77; widths pinned by InvCompress/ours.py:33-55): residual blocks,
    sub-pixel convolutions, no attention, single Gaussian with mean."""

    def __init__(self, N):
        super().__init__(N, N)
        self.g_a = nn.Sequential(
            ResidualBlockWithStride(3, N, 2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), ResidualBlock(N, N), conv3x3(N, N, 2))
        self.h_a = nn.Sequential(conv3x3(N, N), LeakyReLU(), conv3x3(N, N), LeakyReLU(), conv3x3(N, N, 2), LeakyReLU(),
                                 conv3x3(N, N), LeakyReLU(), conv3x3(N, N, 2))
        self.h_s = nn.Sequential(conv3x3(N, N), LeakyReLU(), subpel_conv3x3(N, N, 2), LeakyReLU(),
                                 conv3x3(N, N * 3 // 2), LeakyReLU(),
                                 subpel_conv3x3(N * 3 // 2, N * 3 // 2, 2), LeakyReLU(),
                                 conv3x3(N * 3 // 2, N * 2))
        self.g_s = nn.Sequential(
            ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
            ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))


class ae_onelayer(MeanScaleHyperprior):
    """The reference's ``debug`` model (anchors/model.py:9-35): one 3x3 stride-1 conv as ``g_a``, one 3x3 stride-1
    transposed conv as ``g_s``; its forward decodes the UNQUANTISED latent (``x_hat = g_s(y)``, :31-32) while the
    likelihoods come from the mean-scale hyperprior."""

    def __init__(self, N, M, **kwargs):
        super().__init__(N, M)
        self.g_a = CodecStack(conv(3, M, kernel_size=3, stride=1))
        self.g_s = CodecStack(deconv(M, 3, kernel_size=3, stride=1))

    def forward(self, x):
        with _param_grads_on(self.training):
            y = self.g_a(x)
            z = self.h_a(y)
            z_hat, z_lik = self.entropy_bottleneck(z)
            gp = self.h_s(z_hat)
            half = gp.shape[1] // 2
            scales_hat, means_hat = Fn.NarrowFn.apply(gp, 0, half), Fn.NarrowFn.apply(gp, half, half)
            _, y_lik = self.gaussian_conditional(y, scales_hat, means=means_hat)
            return {"x_hat": self.g_s(y), "likelihoods": {"y": y_lik, "z": z_lik}}


class ResidualUnit(nn.Module):
    """compressai.layers.AttentionBlock.ResidualUnit: 1x1 (N -> N/2), ReLU, 3x3, ReLU, 1x1 (N/2 -> N), + x, ReLU."""

    def __init__(self, N):
        super().__init__()
        self.conv = nn.Sequential(conv1x1(N, N // 2), ReLU(), conv3x3(N // 2, N // 2), ReLU(), conv1x1(N // 2, N))

    def forward(self, x):
        return Fn.ActFn.apply(Fn.AddFn.apply(self.conv(x), x), L.ACT_RELU)


class AttentionBlock(nn.Module):
    """compressai.layers.AttentionBlock (the simplified, non-local-free attention of cheng2020_attn):
    ``conv_a(x) * sigmoid(conv_b(x)) + x`` with three residual units per branch and a closing 1x1 on the gate."""

    def __init__(self, N):
        super().__init__()
        self.conv_a = nn.Sequential(ResidualUnit(N), ResidualUnit(N), ResidualUnit(N))
        self.conv_b = nn.Sequential(ResidualUnit(N), ResidualUnit(N), ResidualUnit(N), conv1x1(N, N))

    def forward(self, x):
        return Fn.GateFn.apply(self.conv_a(x), self.conv_b(x), x)


class Cheng2020Attention(Cheng2020Anchor):
    """compressai cheng2020_attn (BASELINE config 4 names it): the anchor with an AttentionBlock after
    the second strided block and at the end of ``g_a``, mirrored in ``g_s``."""

    def __init__(self, N):
        super().__init__(N)
        self.g_a = nn.Sequential(
            ResidualBlockWithStride(3, N, 2), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), AttentionBlock(N), ResidualBlock(N, N),
            ResidualBlockWithStride(N, N, 2), ResidualBlock(N, N), conv3x3(N, N, 2), AttentionBlock(N))
        self.g_s = nn.Sequential(
            AttentionBlock(N), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2), ResidualBlock(N, N),
            ResidualBlockUpsample(N, N, 2), AttentionBlock(N), ResidualBlock(N, N), ResidualBlockUpsample(N, N, 2),
            ResidualBlock(N, N), subpel_conv3x3(N, 3, 2))


def _build(model, quality):
    cfg = ZOO[model][quality]
    if model == "factorized":
        return FactorizedPrior(*cfg)
    if model == "hyper":
        return ScaleHyperprior(*cfg)
    if model == "context":
        return JointAutoregressiveHierarchicalPriors(*cfg)
    if model == "cheng2020":
        return Cheng2020Anchor(*cfg)
    if model == "cheng2020_attn":
        return Cheng2020Attention(*cfg)
    if model == "debug":
        return ae_onelayer(*cfg)
    raise L.IcadvError(f"unknown model family '{model}'")


def _no_zoo(pretrained):
    if pretrained:
        raise L.IcadvError("no network in this environment: pass pretrained=False (--new) and load a state_dict")


def bmshj2018_factorized(quality, metric="mse", pretrained=False, **kw):
    _no_zoo(pretrained)
    return _build("factorized", quality)


def bmshj2018_hyperprior(quality, metric="mse", pretrained=False, **kw):
    _no_zoo(pretrained)
    return _build("hyper", quality)


def mbt2018(quality, metric="mse", pretrained=False, **kw):
    _no_zoo(pretrained)
    return _build("context", quality)


def cheng2020_anchor(quality, metric="mse", pretrained=False, **kw):
    _no_zoo(pretrained)
    return _build("cheng2020", quality)


def cheng2020_attn(quality, metric="mse", pretrained=False, **kw):
    _no_zoo(pretrained)
    return _build("cheng2020_attn", quality)


def init_model(MODEL, quality, metric="mse", pretrained=False):
    """Same dispatch as anchors/model.py:60-78 (+ ``cheng2020_attn``, the zoo entry the reference leaves commented
    out next to ``cheng2020_anchor``)."""
    if MODEL == "debug":                      # anchors/model.py:61-68: no zoo entry, --new only
        if pretrained:
            raise L.IcadvError("No download-able model available!")
        return _build("debug", quality)
    table = {"factorized": bmshj2018_factorized, "hyper": bmshj2018_hyperprior, "context": mbt2018,
             "cheng2020": cheng2020_anchor, "cheng2020_attn": cheng2020_attn}
    if MODEL not in table:
        raise L.IcadvError(f"unknown model '{MODEL}'")
    return table[MODEL](quality=quality, metric=metric, pretrained=pretrained)
