"""Autograd-aware operators at the reference's operator surface (NCHW tensors in, NCHW out).

Every op here is a ``torch.autograd.Function`` whose forward and backward call the C ABI
(``ops.py``); tensors are kept ``channels_last`` so the NCHW <-> channels-last hop is a view.
Mirrors: nn.Conv2d / nn.ConvTranspose2d (anchors/utils.py:112-130), compressai GDN
(utils/ops.py:58-97), ops.Low_bound / ops.Up_bound (utils/ops.py:28-56).
"""
import weakref

import torch

from . import _lib as L
from . import ops
from . import precision

_PACK_CACHE = {}
_REC = None   # tape.Recorder while a stack is being traced into a static launch program (tape.py)


def invalidate_pack_cache():
    """Drop every packed weight (call after parameters were updated behind autograd's version counters)."""
    _PACK_CACHE.clear()


def to_nhwc(x):
    """NCHW (any memory format) -> contiguous [N,H,W,C] view/copy."""
    return x.contiguous(memory_format=torch.channels_last).permute(0, 2, 3, 1)


def to_nchw(x_nhwc):
    """[N,H,W,C] contiguous -> NCHW-shaped channels_last view (no copy)."""
    return x_nhwc.permute(0, 3, 1, 2)


def tc_shape(k_ch, n_ch):
    """Shapes the tensor path takes (mirrors icadv_conv_tc_supported for the linear epilogue)."""
    if k_ch % 32 or n_ch % 32 or k_ch < 32 or n_ch < 32:
        return False
    return n_ch <= 256 or any(n_ch % t == 0 for t in range(256, 31, -32))   # wider outputs are tiled over N


def packed(weight, kind, split_geom=None, pad_to=None):
    """Packed copy of a layer weight, cached on (storage, version) so it is rebuilt only after an update.
    Weights of tensor-path contractions are rounded to TF32 (nearest) here; with ``split_geom = (Ks, G)`` (parity mode)
    they are returned as G slices of [Whi | Whi | Wlo] instead (csrc/icadv_split.cu); with ``pad_to = (n32, k32)`` rounded
    and zero-padded to channel counts the tensor path takes."""
    key = (id(weight), kind, split_geom, pad_to)
    hit = _PACK_CACHE.get(key)
    # the entry must belong to THIS tensor object (ids and addresses are recycled once a model is freed)
    if hit is not None and hit[3]() is weight and hit[0] == weight._version and hit[2] == tuple(weight.shape) \
            and hit[4] == weight.data_ptr():
        return hit[1]
    if len(_PACK_CACHE) > 512:
        for k in [k for k, v in _PACK_CACHE.items() if v[3]() is None]:
            del _PACK_CACHE[k]
    a, b = weight.shape[0], weight.shape[1]
    n_ch, k_ch = {L.PACK_CONV_FWD: (a, b), L.PACK_CONV_DGRAD: (b, a), L.PACK_CONVT_FWD: (b, a),
                  L.PACK_CONVT_DGRAD: (a, b)}[kind]
    if pad_to is not None:
        pk = ops.pack_weight(weight, kind, round_tf32=True)
        wp = torch.zeros(pk.shape[0], pad_to[0], pad_to[1], device=pk.device, dtype=torch.float32)
        wp[:, :pk.shape[1], :pk.shape[2]].copy_(pk)
    elif split_geom is not None:
        wp = ops.split3_weight(ops.pack_weight(weight, kind, round_tf32=False), *split_geom)
    else:
        wp = ops.pack_weight(weight, kind, round_tf32=tc_shape(k_ch, n_ch) and not precision.split())
    _PACK_CACHE[key] = (weight._version, wp, tuple(weight.shape), weakref.ref(weight), weight.data_ptr())
    return wp


def round_operand(xn, n_ch):
    """The TF32 (round-to-nearest) copy of an activation that feeds tensor-path contractions in the speed mode; the
    tensor itself where no rounding applies (parity mode, shapes on the CUDA-core kernels)."""
    if precision.split() or not tc_shape(xn.shape[-1], n_ch):
        return xn
    return ops.unary(xn, 5)


def contract(xn, weight, kind, bias, *, form, ksize, stride, n_ch, act=L.ACT_NONE, rounded=False):
    """One contraction of the module-by-module path on a channels-last activation, in the current precision mode:
    speed mode -- operand rounded to TF32 (nearest; ``rounded`` = the caller already did), one launch; parity mode --
    operand and weight in the K-sliced three-term split form, one tensor-path launch per slice, partial outputs added in
    fp32.  Shapes the tensor path does not take run on the fp32 CUDA-core kernels in both modes."""
    k_ch = xn.shape[-1]
    if not tc_shape(k_ch, n_ch):
        kp, npad = (k_ch + 31) // 32 * 32, (n_ch + 31) // 32 * 32
        if not precision.split() and tc_shape(kp, npad):
            # speed mode: channel counts the tensor path does not take as they are (the 3x3/2 RGB ends and the 12-channel
            # sub-pixel conv of cheng2020) run on it zero-padded to multiples of 32 -- a few MB of extra traffic against a
            # CUDA-core kernel at 1 - 5 % of its HBM bound (same treatment as tape.TapeProgram, same bits)
            if kp != k_ch:
                xp = torch.zeros(*xn.shape[:-1], kp, device=xn.device, dtype=torch.float32)
                ops.copy_channels(xn.contiguous(), xp, 0, 0, k_ch)
            else:
                xp = xn
            xp = ops.unary(xp, 5)
            bp = None
            if bias is not None:
                bp = torch.zeros(npad, device=xn.device, dtype=torch.float32)
                bp[:n_ch].copy_(bias)
            outp = ops.conv(xp, packed(weight, kind, pad_to=(npad, kp)), bp, form=form, ksize=ksize, stride=stride,
                            n_ch=npad, act=act)
            if npad == n_ch:
                return outp
            out = torch.empty(*outp.shape[:-1], n_ch, device=xn.device, dtype=torch.float32)
            return ops.copy_channels(outp, out, 0, 0, n_ch)
        return ops.conv(xn, packed(weight, kind), bias, form=form, ksize=ksize, stride=stride, n_ch=n_ch, act=act,
                        path="simt" if precision.split() else "auto")
    if precision.split():
        geom = ops.split_geom(k_ch, form, ksize, stride)
        return ops.conv_sliced(ops.split3(xn.contiguous(), *geom), packed(weight, kind, geom), bias, form=form,
                               ksize=ksize, stride=stride, n_ch=n_ch, act=act)
    return ops.conv(xn if rounded else ops.unary(xn, 5), packed(weight, kind), bias, form=form, ksize=ksize,
                    stride=stride, n_ch=n_ch, act=act)


class Contraction(torch.autograd.Function):
    """Conv2d (transposed=False) or ConvTranspose2d (transposed=True), padding k//2, output_padding s-1."""

    @staticmethod
    def forward(ctx, x, weight, bias, ksize, stride, transposed, act, param_grads):
        xn = to_nhwc(x)
        kind = L.PACK_CONVT_FWD if transposed else L.PACK_CONV_FWD
        form = L.FORM_TCONV if transposed else L.FORM_SCONV
        n_ch = weight.shape[1] if transposed else weight.shape[0]
        # speed mode: the operand is rounded to TF32 once; that copy feeds this contraction and, saved, the tensor-core
        # weight gradient (both read their operands as TF32: rounding where produced keeps them unbiased)
        xr = round_operand(xn, n_ch)
        out = contract(xr, weight, kind, bias.detach() if bias is not None else None, form=form, ksize=ksize,
                       stride=stride, n_ch=n_ch, act=act, rounded=True)
        if param_grads and xr is xn and not precision.split() and \
                ops.conv_wgrad_tc_supported(xn.shape[-1], n_ch, ksize, stride, form, xn.shape[1:3]):
            xr = ops.unary(xn, 5)          # RGB end-layer forms: rounded copy for the weight gradient only
        ctx.save_for_backward(xr if param_grads else xn, weight, out if act != L.ACT_NONE else None)
        ctx.cfg = (ksize, stride, transposed, act, param_grads, bias is not None)
        if _REC is not None:
            _REC.note("conv", [xn], out, weight=weight, bias=bias, ksize=ksize, stride=stride, transposed=transposed,
                      act=act, n_ch=n_ch)
        return to_nchw(out)

    @staticmethod
    def backward(ctx, g):
        xn, weight, out = ctx.saved_tensors
        ksize, stride, transposed, act, param_grads, has_bias = ctx.cfg
        gn = to_nhwc(g)
        if act != L.ACT_NONE:
            gn = ops.act_backward(out, gn, act)
        gn = gn.contiguous()
        gx = gw = gb = None
        want_w = param_grads and (ctx.needs_input_grad[1] or (has_bias and ctx.needs_input_grad[2]))
        gr = round_operand(gn, xn.shape[-1]) if (ctx.needs_input_grad[0] or want_w) else gn
        if ctx.needs_input_grad[0]:
            kind = L.PACK_CONVT_DGRAD if transposed else L.PACK_CONV_DGRAD
            form = L.FORM_SCONV if transposed else L.FORM_TCONV
            gx = to_nchw(contract(gr, weight, kind, None, form=form, ksize=ksize, stride=stride, n_ch=xn.shape[-1],
                                  rounded=True))
        if want_w:
            form = L.FORM_TCONV if transposed else L.FORM_SCONV
            n_ch = weight.shape[1] if transposed else weight.shape[0]
            # K6: tcgen05 weight gradient on the TF32-rounded operands in the speed mode; fp32 CUDA-core kernel in the
            # parity mode and for the 3-channel end layers
            tc = not precision.split() and ops.conv_wgrad_tc_supported(xn.shape[-1], n_ch, ksize, stride, form,
                                                                       xn.shape[1:3])
            if tc and gr is gn:
                gr = ops.unary(gn, 5)
            dwp, db = ops.conv_wgrad(xn, gr if tc else gn, form=form, ksize=ksize, stride=stride, n_ch=n_ch,
                                     want_bias=has_bias, path="tc" if tc else "simt")
            gw = ops.unpack_weight_grad(dwp, weight, L.PACK_CONVT_FWD if transposed else L.PACK_CONV_FWD)
            gb = db
        return gx, gw, gb, None, None, None, None, None


class GdnFn(torch.autograd.Function):
    """y = x * (beta + gamma x^2)^(-1/2)  (inverse: ^(+1/2)); beta/gamma already reparametrised."""

    @staticmethod
    def forward(ctx, x, beta_eff, gamma_eff, inverse):
        xn = to_nhwc(x)
        C = xn.shape[-1]
        ctx.split = precision.split()
        if ctx.split:
            y, sc = ops.gdn_forward_split(xn.contiguous(), beta_eff, gamma_eff, inverse)
        else:
            y, sc = ops.conv(xn, None, None, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C,
                             epi=L.EPI_IGDN_FWD if inverse else L.EPI_GDN_FWD, gmat=gamma_eff.contiguous(),
                             beta=beta_eff.contiguous(), acc_from_in=True, path="tc")
        ctx.save_for_backward(y, sc, gamma_eff)
        ctx.inverse = inverse
        if _REC is not None:
            _REC.note("gdn", [xn], y, module=_REC.current_gdn, inverse=inverse)
        return to_nchw(y)

    @staticmethod
    def backward(ctx, g):
        y, sc, gamma_eff = ctx.saved_tensors
        C = y.shape[-1]
        if ctx.split:
            return to_nchw(ops.gdn_backward_split(to_nhwc(g).contiguous(), y, sc, gamma_eff, ctx.inverse)), None, None, None
        gx = ops.conv(to_nhwc(g), None, None, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C,
                      epi=L.EPI_IGDN_BWD if ctx.inverse else L.EPI_GDN_BWD, gmat=gamma_eff.t().contiguous(),
                      y_prev=y, sc_prev=sc, acc_from_in=True, path="tc")
        return to_nchw(gx), None, None, None


class BoundFn(torch.autograd.Function):
    """ops.Low_bound / ops.Up_bound (utils/ops.py:28-56): clamp forward, pass-through-if backward."""

    @staticmethod
    def forward(ctx, x, bound, upper):
        xc = x.contiguous()
        ctx.save_for_backward(xc)
        ctx.cfg = (float(bound), bool(upper))
        return ops.bound_forward(xc.view(-1), float(bound), upper).view_as(xc)

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        bound, upper = ctx.cfg
        return ops.bound_backward(xc.view(-1), g.contiguous().view(-1), bound, upper).view_as(xc), None, None


class Low_bound:
    """Drop-in for ``ops.Low_bound`` (``.apply(x, b)``)."""

    @staticmethod
    def apply(x, lower_bound=1e-6):
        return BoundFn.apply(x, lower_bound, False)


class Up_bound:
    """Drop-in for ``ops.Up_bound`` (``.apply(x, b)``)."""

    @staticmethod
    def apply(x, up_bound=1.0):
        return BoundFn.apply(x, up_bound, True)


class Round_STE(torch.autograd.Function):
    """``ops.Round_STE`` (utils/ops.py:8-15): round to nearest (ties to even, as ``torch.round``) forward, identity backward."""

    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous()
        return ops.unary(xc.view(-1), 3).view_as(xc)

    @staticmethod
    def backward(ctx, g):
        return g


class UniverseQuant(torch.autograd.Function):
    """``ops.UniverseQuant`` (utils/ops.py:17-25): ``round(x + u) - u`` with ``u ~ U(-1/2, 1/2)`` per element, identity
    backward.  The sample comes from the library's Philox kernel on torch's CUDA generator (the reference draws on the CPU
    and copies: a different stream of the same distribution)."""

    @staticmethod
    def forward(ctx, x):
        xc = x.contiguous()
        u = ops.uniform_noise_like(xc)
        q = ops.unary(ops.unary(xc.view(-1), 4, u.view(-1)), 3)          # round(x + u)
        u.neg_()
        return ops.unary(q, 4, u.view(-1)).view_as(xc)                   # - u

    @staticmethod
    def backward(ctx, g):
        return g


class ActFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, act):
        xc = x.contiguous(memory_format=torch.channels_last) if x.dim() == 4 else x.contiguous()
        y = ops.unary(xc, {L.ACT_ABS: 0, L.ACT_RELU: 1, L.ACT_LEAKY: 2}[act])
        ctx.save_for_backward(xc)
        ctx.act = act
        if _REC is not None and xc.dim() == 4:
            _REC.note("act", [to_nhwc(xc)], to_nhwc(y), act=act)
        return y

    @staticmethod
    def backward(ctx, g):
        (xc,) = ctx.saved_tensors
        gc = g.contiguous(memory_format=torch.channels_last) if g.dim() == 4 else g.contiguous()
        return ops.act_backward(xc, gc, ctx.act), None


class AddFn(torch.autograd.Function):
    """a + b (residual connections of the cheng2020 blocks); both operands channels_last, same shape."""

    @staticmethod
    def forward(ctx, a, b):
        ac = a.contiguous(memory_format=torch.channels_last)
        bc = b.contiguous(memory_format=torch.channels_last)
        y = ops.unary(ac, 4, bc)
        if _REC is not None:
            _REC.note("add", [to_nhwc(ac), to_nhwc(bc)], to_nhwc(y))
        return y

    @staticmethod
    def backward(ctx, g):
        return g, g


class GateFn(torch.autograd.Function):
    """a * sigmoid(b) + x: the gate of compressai.layers.AttentionBlock (cheng2020_attn)."""

    @staticmethod
    def forward(ctx, a, b, x):
        cl = torch.channels_last
        ac, bc, xc = a.contiguous(memory_format=cl), b.contiguous(memory_format=cl), x.contiguous(memory_format=cl)
        y = ops.attention_gate(ac, bc, xc)
        ctx.save_for_backward(ac, bc)
        if _REC is not None:
            _REC.note("gate", [to_nhwc(ac), to_nhwc(bc), to_nhwc(xc)], to_nhwc(y))
        return y

    @staticmethod
    def backward(ctx, g):
        ac, bc = ctx.saved_tensors
        gc = g.contiguous(memory_format=torch.channels_last)
        ga, gb = ops.attention_gate_backward(ac, bc, gc)
        return ga, gb, g


class PixelShuffleFn(torch.autograd.Function):
    """nn.PixelShuffle(r) (compressai subpel_conv3x3)."""

    @staticmethod
    def forward(ctx, x, r):
        ctx.r = r
        xn = to_nhwc(x).contiguous()
        y = ops.pixel_shuffle(xn, r)
        if _REC is not None:
            _REC.note("shuffle", [xn], y, r=r)
        return to_nchw(y)

    @staticmethod
    def backward(ctx, g):
        return to_nchw(ops.pixel_shuffle(to_nhwc(g).contiguous(), ctx.r, inverse=True)), None


class CatFn(torch.autograd.Function):
    """torch.cat((a, b), dim=1) on channels_last tensors (anchors/model.py:104)."""

    @staticmethod
    def forward(ctx, a, b):
        an, bn = to_nhwc(a).contiguous(), to_nhwc(b).contiguous()
        ca, cb = an.shape[-1], bn.shape[-1]
        out = torch.empty(*an.shape[:-1], ca + cb, device=a.device, dtype=torch.float32)
        ops.copy_channels(an, out, 0, 0, ca)
        ops.copy_channels(bn, out, 0, ca, cb)
        ctx.split = (ca, cb)
        return to_nchw(out)

    @staticmethod
    def backward(ctx, g):
        ca, cb = ctx.split
        gn = to_nhwc(g).contiguous()
        ga = torch.empty(*gn.shape[:-1], ca, device=g.device, dtype=torch.float32)
        gb = torch.empty(*gn.shape[:-1], cb, device=g.device, dtype=torch.float32)
        ops.copy_channels(gn, ga, 0, 0, ca)
        ops.copy_channels(gn, gb, ca, 0, cb)
        return to_nchw(ga), to_nchw(gb)


class NarrowFn(torch.autograd.Function):
    """x[:, start:start+count] (one half of ``chunk(2, 1)``, anchors/model.py:105) as a dense channels_last tensor."""

    @staticmethod
    def forward(ctx, x, start, count):
        xn = to_nhwc(x).contiguous()
        out = torch.empty(*xn.shape[:-1], count, device=x.device, dtype=torch.float32)
        ops.copy_channels(xn, out, start, 0, count)
        ctx.cfg = (start, count, xn.shape[-1])
        return to_nchw(out)

    @staticmethod
    def backward(ctx, g):
        start, count, c = ctx.cfg
        gn = to_nhwc(g).contiguous()
        gx = torch.zeros(*gn.shape[:-1], c, device=g.device, dtype=torch.float32)
        ops.copy_channels(gn, gx, 0, start, count)
        return to_nchw(gx), None, None


# ------------------------------------------------------------------ codec update of adversarial training (train.py --adv)
class GdnTrainFn(torch.autograd.Function):
    """GDN / IGDN with gradients to the RAW beta / gamma parameters (compressai GDN under train.py:359)."""

    @staticmethod
    def forward(ctx, x, beta_raw, gamma_raw, inverse, beta_bound, beta_ped, gamma_bound, gamma_ped):
        be = ops.gdn_reparam(beta_raw, beta_bound, beta_ped)
        ctx.split = precision.split()
        ga = ops.gdn_reparam(gamma_raw, gamma_bound, gamma_ped, round_tf32=not ctx.split)
        xn = to_nhwc(x)
        C = xn.shape[-1]
        if ctx.split:
            y, sc = ops.gdn_forward_split(xn.contiguous(), be, ga, inverse)
        else:
            y, sc = ops.conv(xn, None, None, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C,
                             epi=L.EPI_IGDN_FWD if inverse else L.EPI_GDN_FWD, gmat=ga.contiguous(), beta=be.contiguous(),
                             acc_from_in=True, path="tc")
        ctx.save_for_backward(y, sc, ga, beta_raw, gamma_raw)
        ctx.cfg = (inverse, beta_bound, gamma_bound)
        return to_nchw(y)

    @staticmethod
    def backward(ctx, g):
        y, sc, ga, beta_raw, gamma_raw = ctx.saved_tensors
        inverse, beta_bound, gamma_bound = ctx.cfg
        C = y.shape[-1]
        gn = to_nhwc(g).contiguous()
        if ctx.split:
            gx = ops.gdn_backward_split(gn, y, sc, ga, inverse)
        else:
            gx = ops.conv(gn, None, None, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C,
                          epi=L.EPI_IGDN_BWD if inverse else L.EPI_GDN_BWD, gmat=ga.t().contiguous(), y_prev=y, sc_prev=sc,
                          acc_from_in=True, path="tc")
        gb, gg = ops.gdn_param_grad(gn, y, sc, beta_raw, gamma_raw, inverse=inverse, beta_bound=beta_bound,
                                    gamma_bound=gamma_bound)
        return to_nchw(gx), gb, gg, None, None, None, None, None


class EbTrainFn(torch.autograd.Function):
    """Train-mode EntropyBottleneck (x_hat = x + noise, factorised likelihood) with gradients to x and to the
    matrices / biases / factors (anchors/model.py:87-89,93 under train.py:353-359)."""

    @staticmethod
    def forward(ctx, x, noise_nhwc, medians, lik_bound, *params):
        ms, bs, fs = params[0:5], params[5:10], params[10:14]
        xn = to_nhwc(x).contiguous()
        C = xn.shape[-1]
        table = ops.eb_prepare(ms, bs, fs, C)
        if noise_nhwc is None:
            noise_nhwc = ops.uniform_noise_like(xn)
        x_hat, lik, _ = ops.eb_forward(xn, table, medians, training=True, noise=noise_nhwc, lik_bound=lik_bound)
        ctx.save_for_backward(x_hat, table, *params)
        ctx.lik_bound = lik_bound
        return to_nchw(x_hat), to_nchw(lik)

    @staticmethod
    def backward(ctx, g_xhat, g_lik):
        x_hat, table = ctx.saved_tensors[0:2]
        params = ctx.saved_tensors[2:]
        ms, fs = params[0:5], params[10:14]
        C = x_hat.shape[-1]
        gl = to_nhwc(g_lik).contiguous() if g_lik is not None else torch.zeros_like(x_hat)
        g_x, g_raw = ops.eb_backward(x_hat, gl, table, ms, fs, lik_bound=ctx.lik_bound)
        gx = to_nchw(g_x)
        if g_xhat is not None:
            gx = ops.unary(gx.contiguous(memory_format=torch.channels_last), 4,
                           g_xhat.contiguous(memory_format=torch.channels_last))
        rows = lambda a, b, shape: g_raw[a:b].t().reshape(C, *shape).contiguous()
        gm = [rows(0, 3, (3, 1))] + [rows(9 + 15 * i, 18 + 15 * i, (3, 3)) for i in range(3)] + [rows(54, 57, (1, 3))]
        gb = [rows(3, 6, (3, 1))] + [rows(18 + 15 * i, 21 + 15 * i, (3, 1)) for i in range(3)] + [rows(57, 58, (1, 1))]
        gf = [rows(6, 9, (3, 1))] + [rows(21 + 15 * i, 24 + 15 * i, (3, 1)) for i in range(3)]
        return (gx, None, None, None, *gm, *gb, *gf)


class GcTrainFn(torch.autograd.Function):
    """Train-mode GaussianConditional (y_hat = y + noise, discretised Gaussian likelihood) with gradients to y, the
    scales and the means (anchors/model.py:95,106)."""

    @staticmethod
    def forward(ctx, y, scales, means, noise_nhwc, scale_bound, lik_bound):
        yn, sn = to_nhwc(y).contiguous(), to_nhwc(scales).contiguous()
        mn = to_nhwc(means).contiguous() if means is not None else None
        if noise_nhwc is None:
            noise_nhwc = ops.uniform_noise_like(yn)
        y_hat, lik, _ = ops.gc_forward(yn, sn, mn, training=True, noise=noise_nhwc, scale_bound=scale_bound,
                                       lik_bound=lik_bound)
        ctx.save_for_backward(y_hat, sn, mn)
        ctx.cfg = (scale_bound, lik_bound)
        return to_nchw(y_hat), to_nchw(lik)

    @staticmethod
    def backward(ctx, g_yhat, g_lik):
        y_hat, sn, mn = ctx.saved_tensors
        scale_bound, lik_bound = ctx.cfg
        gl = to_nhwc(g_lik).contiguous() if g_lik is not None else torch.zeros_like(y_hat)
        g_y, g_s, g_m = ops.gc_backward(y_hat, sn, mn, gl, scale_bound=scale_bound, lik_bound=lik_bound)
        gy = to_nchw(g_y)
        if g_yhat is not None:
            gy = ops.unary(gy.contiguous(memory_format=torch.channels_last), 4,
                           g_yhat.contiguous(memory_format=torch.channels_last))
        return gy, to_nchw(g_s), (to_nchw(g_m) if g_m is not None else None), None, None, None


class LogSumFn(torch.autograd.Function):
    """sum log(clamp(lik, min=floor)) (train.py:62-64) -> 1-element tensor; gradient g / lik inside the clamp range."""

    @staticmethod
    def forward(ctx, lik, floor):
        lk = lik if (lik.is_contiguous() or lik.is_contiguous(memory_format=torch.channels_last)) else lik.contiguous()
        ctx.save_for_backward(lk)
        ctx.floor = floor
        return ops.log_sum(lk, floor)

    @staticmethod
    def backward(ctx, g):
        (lk,) = ctx.saved_tensors
        return ops.log_sum_backward(lk, ctx.floor, g.contiguous()), None


class MseFn(torch.autograd.Function):
    """mean((x - t)^2) (train.py:71) -> 1-element tensor; gradient 2 g (x - t) / numel to x."""

    @staticmethod
    def forward(ctx, x, t):
        xc = x.contiguous(memory_format=torch.channels_last)
        tc = t.contiguous(memory_format=torch.channels_last)
        ctx.save_for_backward(xc, tc)
        flat = lambda a: a.permute(0, 2, 3, 1).reshape(1, -1)
        return ops.sum_sqdiff(flat(xc), flat(tc)) * (1.0 / xc.numel())

    @staticmethod
    def backward(ctx, g):
        xc, tc = ctx.saved_tensors
        return ops.scaled_diff(xc, tc, g.contiguous(), 2.0 / xc.numel()), None
