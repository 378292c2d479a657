"""MS-SSIM at the reference's call surface, on the fused level kernel (csrc/icadv_msssim.cu).

``ms_ssim`` / ``MS_SSIM`` mirror ``pytorch_msssim`` (attack_rd.py:19; train.py:18,44) -- variant 1;
``MS_SSIM_v2`` mirrors ``utils/torch_msssim.MS_SSIM`` (utils/torch_msssim.py:18-76) -- variant 2.
All three are differentiable with respect to BOTH images (autograd Functions on the one-pass value-and-gradient level
kernel for the 11-tap window, on ``ssim_level_backward_kernel`` otherwise): the unmodified ``attack_our`` ms-ssim branches (``1 - ms_ssim(im_s, im_in)``,
``ms_ssim(output_, output_s)``, attack_rd.py:336,362) and the ms-ssim RD loss (train.py:44,88) differentiate through them.
"""
import ctypes as C
import math

import torch

from . import _lib as L
from .ops import _p, _stream

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)
_CONST = {}


def _weights_tensor(device, weights):
    """[levels, 1, 1] device tensor of the level weights, created once per device (a host -> device copy from pageable
    memory is not allowed while a CUDA graph is being captured; the attack loop captures this composition)."""
    key = (str(device), tuple(weights))
    t = _CONST.get(key)
    if t is None:
        t = torch.tensor(weights, device=device, dtype=torch.float32).view(-1, 1, 1)
        _CONST[key] = t
    return t


def _taps(size, sigma):
    g = [math.exp(-((i - size // 2) ** 2) / (2.0 * sigma * sigma)) for i in range(size)]
    s = sum(g)
    return [v / s for v in g]


def _level(X, Y, taps, same_pad, c1, c2):
    planes, h, w = X.shape[0] * X.shape[1], X.shape[2], X.shape[3]
    n = L.lib().icadv_ssim_workspace_floats(planes, h, w, len(taps), 1 if same_pad else 0)
    ws = torch.empty(max(n, 1), device=X.device, dtype=torch.float32)
    ss = torch.empty(planes, device=X.device, dtype=torch.float32)
    cs = torch.empty(planes, device=X.device, dtype=torch.float32)
    arr = (C.c_float * len(taps))(*taps)
    L.call("icadv_ssim_level", _p(X), _p(Y), _p(ws), _p(ss), _p(cs), planes, h, w, arr, len(taps),
           1 if same_pad else 0, float(c1), float(c2), _stream())
    oh, ow = (h, w) if same_pad else (h - len(taps) + 1, w - len(taps) + 1)
    return ss.view(X.shape[0], X.shape[1]) / (oh * ow), cs.view(X.shape[0], X.shape[1]) / (oh * ow)


def _pool(X, ph, pw):
    n, c, h, w = X.shape
    oh, ow = (h + 2 * ph - 2) // 2 + 1, (w + 2 * pw - 2) // 2 + 1
    out = torch.empty(n, c, oh, ow, device=X.device, dtype=torch.float32)
    L.call("icadv_avgpool2", _p(X), _p(out), n * c, h, w, ph, pw, _stream())
    return out


class _MsSsimFn(torch.autograd.Function):
    """Variant 1 with gradients to both images.  The per-image value is linear in nothing but its own upstream weight,
    so the forward runs the one-pass value-and-gradient composition once per differentiable argument with UNIT upstream
    and keeps that gradient; the backward only scales it (the reference's autograd walks the pyramid twice).  SSIM is
    symmetric in (X, Y): the gradient with respect to Y is the same composition with the images swapped."""

    @staticmethod
    def forward(ctx, X, Y, size_average, kw):
        B = X.shape[0]
        ones = torch.ones(B, device=X.device, dtype=torch.float32)
        Xd, Yd = X.detach(), Y.detach()
        val = gX = gY = None
        if X.requires_grad:
            val, gX = ms_ssim_value_and_grad(Xd, Yd, ones, **kw)
        if Y.requires_grad:
            val, gY = ms_ssim_value_and_grad(Yd, Xd, ones, **kw)
        ctx.size_average, ctx.has = size_average, (gX is not None, gY is not None)
        ctx.save_for_backward(*[g for g in (gX, gY) if g is not None])
        return val.mean() if size_average else val

    @staticmethod
    def backward(ctx, g):
        saved = list(ctx.saved_tensors)
        gX = saved.pop(0) if ctx.has[0] else None
        gY = saved.pop(0) if ctx.has[1] else None
        B = (gX if gX is not None else gY).shape[0]
        up = (g.reshape(1).expand(B) / B) if ctx.size_average else g.reshape(B)
        up = up.to(torch.float32).view(B, 1, 1, 1)
        return (gX * up if gX is not None else None), (gY * up if gY is not None else None), None, None


def ms_ssim(X, Y, data_range=1.0, size_average=True, win_size=11, win_sigma=1.5, weights=WEIGHTS, K=(0.01, 0.03)):
    """pytorch_msssim.ms_ssim (variant 1); differentiable w.r.t. X and Y."""
    if X.shape != Y.shape or X.dim() != 4:
        raise ValueError("Input images should be 4-d tensors of the same shape")
    assert min(X.shape[-2:]) > (win_size - 1) * 2 ** 4, \
        "Image size should be larger than %d due to the 4 downsamplings in ms-ssim" % ((win_size - 1) * 2 ** 4)
    kw = dict(data_range=data_range, win_size=win_size, win_sigma=win_sigma, weights=tuple(weights), K=tuple(K))
    if torch.is_grad_enabled() and (X.requires_grad or Y.requires_grad):
        return _MsSsimFn.apply(X, Y, size_average, kw)
    return _ms_ssim_forward(X, Y, size_average=size_average, **kw)


def _ms_ssim_forward(X, Y, data_range=1.0, size_average=True, win_size=11, win_sigma=1.5, weights=WEIGHTS, K=(0.01, 0.03)):
    X, Y = X.detach().contiguous().float(), Y.detach().contiguous().float()
    taps = _taps(win_size, win_sigma)
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    vals = []
    for lvl in range(len(weights)):
        ss, cs = _level(X, Y, taps, False, c1, c2)
        if lvl < len(weights) - 1:
            vals.append(torch.relu(cs))
            ph, pw = X.shape[2] % 2, X.shape[3] % 2
            X, Y = _pool(X, ph, pw), _pool(Y, ph, pw)
    vals.append(torch.relu(ss))
    w = _weights_tensor(X.device, weights)
    val = torch.prod(torch.stack(vals, 0) ** w, dim=0)
    return val.mean() if size_average else val.mean(1)


def _level_bwd(X, Y, coef_cs, coef_ss, dXnext, pads, taps, c1, c2, same_pad=False):
    n, c, h, w = X.shape
    dX = torch.empty_like(X)
    nh, nw = (dXnext.shape[2], dXnext.shape[3]) if dXnext is not None else (0, 0)
    arr = (C.c_float * len(taps))(*taps)
    L.call("icadv_ssim_level_backward", _p(X), _p(Y), _p(coef_cs), _p(coef_ss), _p(dXnext), _p(dX), n * c, h, w, nh, nw,
           pads[0], pads[1], arr, len(taps), 1 if same_pad else 0, float(c1), float(c2), _stream())
    return dX


def _level_value_grad(X, Y, taps, same_pad, c1, c2, last):
    """One level through the row-marching kernel: (mean ssim, mean cs) per plane and the unit gradient U of the level's
    summed map (cs, or ssim at the last level) with respect to X."""
    planes, h, w = X.shape[0] * X.shape[1], X.shape[2], X.shape[3]
    n = L.lib().icadv_ssim_vg_workspace_floats(planes, h, w, 1 if same_pad else 0)
    ws = torch.empty(max(n, 1), device=X.device, dtype=torch.float32)
    ss = torch.empty(planes, device=X.device, dtype=torch.float32)
    cs = torch.empty(planes, device=X.device, dtype=torch.float32)
    U = torch.empty_like(X)
    arr = (C.c_float * len(taps))(*taps)
    L.call("icadv_ssim_level_value_grad", _p(X), _p(Y), _p(U), _p(ws), _p(ss), _p(cs), planes, h, w, arr, len(taps),
           1 if same_pad else 0, float(c1), float(c2), 1 if last else 0, _stream())
    oh, ow = (h, w) if same_pad else (h - len(taps) + 1, w - len(taps) + 1)
    return ss.view(X.shape[0], X.shape[1]) / (oh * ow), cs.view(X.shape[0], X.shape[1]) / (oh * ow), U


def _combine(U, coef, Dnext, pads):
    planes, h, w = U.shape[0] * U.shape[1], U.shape[2], U.shape[3]
    nh, nw = (Dnext.shape[2], Dnext.shape[3]) if Dnext is not None else (0, 0)
    L.call("icadv_ssim_combine", _p(U), _p(coef), _p(Dnext), planes, h, w, nh, nw, pads[0], pads[1], _stream())
    return U


FUSED_VALUE_GRAD = True   # False: the two-pass tile kernels (level forward, then level backward) -- kept for A/B tests


def ms_ssim_value_and_grad(X, Y, upstream, data_range=1.0, win_size=11, win_sigma=1.5, weights=WEIGHTS, K=(0.01, 0.03),
                           y_pyramid=None):
    """Per-image MS-SSIM (variant 1, mean over channels) and the gradient of ``sum_b upstream[b] * value[b]`` with
    respect to X.  This is what autograd of ``ms_ssim(X, Y)`` delivers in attack_rd.py:336,362, one image per row.

    11-tap window (the reference's): ONE pass per level produces the level value and the unit gradient of its map
    (``icadv_ssim_level_value_grad``); the per-plane chain-rule weights and the pooling chain are applied afterwards,
    coarse to fine, by ``icadv_ssim_combine``."""
    assert X.shape == Y.shape and X.dim() == 4 and min(X.shape[-2:]) > (win_size - 1) * 2 ** 4
    X, Y = X.detach().contiguous().float(), Y.detach().contiguous().float()
    if win_size == 11 and FUSED_VALUE_GRAD:
        return _value_and_grad_fused(X, Y, upstream, data_range, win_size, win_sigma, weights, K, y_pyramid)
    B, Cc = X.shape[0], X.shape[1]
    taps = _taps(win_size, win_sigma)
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    nl = len(weights)
    xs, ys, pads, vals, npx = [X], [Y], [], [], []
    for lvl in range(nl):
        ss, cs = _level(xs[-1], ys[-1], taps, False, c1, c2)
        vals.append(torch.relu(cs if lvl < nl - 1 else ss))
        npx.append((xs[-1].shape[2] - win_size + 1) * (xs[-1].shape[3] - win_size + 1))
        if lvl < nl - 1:
            ph, pw = xs[-1].shape[2] % 2, xs[-1].shape[3] % 2
            pads.append((ph, pw))
            xs.append(_pool(xs[-1], ph, pw))
            ys.append(_pool(ys[-1], ph, pw))
    w = _weights_tensor(X.device, weights)
    V = torch.stack(vals, 0)                       # [levels, B, C]
    P = torch.prod(V ** w, dim=0)                  # [B, C]
    value = P.mean(1)
    dP = (upstream.view(B, 1).to(torch.float32) / Cc).expand(B, Cc)
    dV = torch.where(V > 0, dP.unsqueeze(0) * w * P.unsqueeze(0) / V.clamp(min=1e-30), torch.zeros_like(V))
    zero = torch.zeros(B * Cc, device=X.device, dtype=torch.float32)
    dnext = None
    for lvl in range(nl - 1, -1, -1):
        coef = (dV[lvl] / npx[lvl]).reshape(-1).contiguous()
        last = lvl == nl - 1
        dnext = _level_bwd(xs[lvl], ys[lvl], zero if last else coef, coef if last else zero, dnext,
                           pads[lvl] if lvl < nl - 1 else (0, 0), taps, c1, c2)
    return value, dnext


def _pool_into(X, out, ph, pw):
    n, c, h, w = X.shape
    L.call("icadv_avgpool2", _p(X), _p(out), n * c, h, w, ph, pw, _stream())
    return out


def reference_pyramid(Y, levels=len(WEIGHTS), out=None):
    """The 2x2-average-pool pyramid of an image batch (variant 1 padding): pass it as ``y_pyramid`` to
    ``ms_ssim_value_and_grad`` when the same reference image meets many first arguments (the attack loop's im_s and
    output_s are fixed for all iterations), so its four pool launches run once per attack instead of once per iteration.
    ``out`` (a pyramid this function returned for the SAME tensor ``Y``): refilled in place -- a captured CUDA graph that
    reads the pyramid keeps seeing the same addresses."""
    assert Y.is_contiguous() and Y.dtype == torch.float32
    if out is not None:
        assert out[0] is Y and len(out) == levels
        for lvl in range(1, levels):
            _pool_into(out[lvl - 1], out[lvl], out[lvl - 1].shape[2] % 2, out[lvl - 1].shape[3] % 2)
        return out
    pyr = [Y]
    for _ in range(levels - 1):
        ph, pw = pyr[-1].shape[2] % 2, pyr[-1].shape[3] % 2
        pyr.append(_pool(pyr[-1], ph, pw))
    return pyr


def _value_and_grad_fused(X, Y, upstream, data_range, win_size, win_sigma, weights, K, y_pyramid=None):
    """Launches: per level one value-and-gradient kernel (block partials stay in one workspace) and the pool of X;
    ONE kernel for all the scalar work (level sums, relu, the weighted product, its gradient: icadv_msssim_coefficients);
    per level one combine kernel, coarse to fine."""
    B, Cc = X.shape[0], X.shape[1]
    planes = B * Cc
    taps = _taps(win_size, win_sigma)
    c1, c2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    nl = len(weights)
    assert nl <= 8
    lib = L.lib()
    # geometry of the whole pyramid first: one workspace for the block partials of every level
    dims, pads = [(X.shape[2], X.shape[3])], []
    for _ in range(nl - 1):
        h, w = dims[-1]
        pads.append((h % 2, w % 2))
        dims.append(((h + 2 * pads[-1][0] - 2) // 2 + 1, (w + 2 * pads[-1][1] - 2) // 2 + 1))
    sizes = [lib.icadv_ssim_vg_workspace_floats(planes, h, w, 0) for h, w in dims]
    offs = [sum(sizes[:i]) for i in range(nl)]
    ws = torch.empty(max(sum(sizes), 1), device=X.device, dtype=torch.float32)
    arr = (C.c_float * len(taps))(*taps)
    x, us = X, []
    for lvl in range(nl):
        y = y_pyramid[lvl] if y_pyramid is not None else (Y if lvl == 0 else _pool(y, *pads[lvl - 1]))
        assert y.shape == x.shape, (y.shape, x.shape)
        U = torch.empty_like(x)
        L.call("icadv_ssim_level_value_grad", _p(x), _p(y), _p(U), ws.data_ptr() + 4 * offs[lvl], None, None, planes,
               dims[lvl][0], dims[lvl][1], arr, len(taps), 0, float(c1), float(c2), 1 if lvl == nl - 1 else 0, _stream())
        us.append(U)
        if lvl < nl - 1:
            x = _pool(x, *pads[lvl])
    value = torch.empty(B, device=X.device, dtype=torch.float32)
    coef = torch.empty(nl, planes, device=X.device, dtype=torch.float32)
    npx = [float((h - win_size + 1) * (w - win_size + 1)) for h, w in dims]
    up = upstream.to(torch.float32).contiguous()
    L.call("icadv_msssim_coefficients", _p(ws), (C.c_int * nl)(*offs), (C.c_int * nl)(*[s // (2 * planes) for s in sizes]),
           (C.c_float * nl)(*npx), (C.c_float * nl)(*[float(v) for v in weights]), nl, _p(up), _p(value), _p(coef), B, Cc,
           _stream())
    dnext = None
    for lvl in range(nl - 1, -1, -1):
        dnext = _combine(us[lvl], coef[lvl], dnext, pads[lvl] if lvl < nl - 1 else (0, 0))
    return value, dnext


class MS_SSIM(torch.nn.Module):
    """pytorch_msssim.MS_SSIM(data_range=1, size_average=True, channel=3)."""

    def __init__(self, data_range=1.0, size_average=True, channel=3, win_size=11, win_sigma=1.5):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size, win_sigma=win_sigma)

    def forward(self, X, Y):
        return ms_ssim(X, Y, **self.kw)


def _v2_levels(X, Y, max_val, levels):
    """Variant 2 pyramid: per level (X, Y, taps, global mean ssim, global mean cs)."""
    c1, c2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
    out = []
    for _ in range(levels):
        wsz = min(X.shape[2], X.shape[3], 11)
        taps = _taps(wsz, 1.5 * wsz / 11)
        ss, cs = _level(X, Y, taps, True, c1, c2)
        out.append((X, Y, taps, ss.mean(), cs.mean()))   # global mean over N, C, H, W (planes have equal size)
        X, Y = _pool(X, 0, 0), _pool(Y, 0, 0)
    return out, c1, c2


def _v2_value(lv, levels):
    w = _weights_tensor(lv[0][0].device, WEIGHTS).view(-1)
    mcs_t, ms_t = torch.stack([l[4] for l in lv]), torch.stack([l[3] for l in lv])
    return torch.prod(mcs_t[:levels - 1] ** w[:levels - 1]) * (ms_t[levels - 1] ** w[levels - 1])


class _MsSsimV2Fn(torch.autograd.Function):
    """utils/torch_msssim.MS_SSIM.forward with gradients to both images (the 2-D window is an outer product of the
    1-D taps, so the separable backward kernel applies with the "same" padding origin)."""

    @staticmethod
    def forward(ctx, X, Y, max_val, levels):
        ctx.save_for_backward(X.detach(), Y.detach())
        ctx.cfg = (max_val, levels)
        lv, _, _ = _v2_levels(X.detach().contiguous().float(), Y.detach().contiguous().float(), max_val, levels)
        return _v2_value(lv, levels)

    @staticmethod
    def _grad(A, Bimg, g, max_val, levels):
        lv, c1, c2 = _v2_levels(A, Bimg, max_val, levels)
        value = _v2_value(lv, levels)
        planes = A.shape[0] * A.shape[1]
        dnext = None
        zero = torch.zeros(planes, device=A.device, dtype=torch.float32)
        for l in range(levels - 1, -1, -1):
            Xl, Yl, taps, ms_l, mcs_l = lv[l]
            last = l == levels - 1
            base = ms_l if last else mcs_l
            # d value / d (global mean of this level's map), spread over the planes * H * W positions of the mean
            coef = (g * WEIGHTS[l] * value / base / (planes * Xl.shape[2] * Xl.shape[3])).reshape(1).expand(planes).contiguous()
            dnext = _level_bwd(Xl, Yl, zero if last else coef, coef if last else zero, dnext, (0, 0), taps, c1, c2,
                               same_pad=True)
        return dnext

    @staticmethod
    def backward(ctx, g):
        X, Y = ctx.saved_tensors
        max_val, levels = ctx.cfg
        Xc, Yc = X.contiguous().float(), Y.contiguous().float()
        g = g.to(torch.float32)
        gX = _MsSsimV2Fn._grad(Xc, Yc, g, max_val, levels) if ctx.needs_input_grad[0] else None
        gY = _MsSsimV2Fn._grad(Yc, Xc, g, max_val, levels) if ctx.needs_input_grad[1] else None
        return gX, gY, None, None


class MS_SSIM_v2(torch.nn.Module):
    """utils/torch_msssim.MS_SSIM(size_average=True, max_val=255); differentiable w.r.t. both images."""

    def __init__(self, size_average=True, max_val=255, device_id=0):
        super().__init__()
        self.max_val = max_val

    def forward(self, img1, img2, levels=5):
        if torch.is_grad_enabled() and (img1.requires_grad or img2.requires_grad):
            return _MsSsimV2Fn.apply(img1, img2, self.max_val, levels)
        lv, _, _ = _v2_levels(img1.detach().contiguous().float(), img2.detach().contiguous().float(), self.max_val, levels)
        return _v2_value(lv, levels)
