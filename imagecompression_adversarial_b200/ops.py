"""Functional layer over the C ABI: torch tensors in, torch tensors out, every FLOP in libicadv_b200.so.

Internal activation layout is channels-last: tensors of shape [N, H, W, C], contiguous, fp32, CUDA.
(At the operator surface these are the ``permute(0, 2, 3, 1)`` view of a ``channels_last`` NCHW tensor,
so crossing it costs no copy.)  PyTorch only provides memory and the stream here.
"""
import ctypes as C

import torch

from . import _lib as L

_plans = {}


def _stream():
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t, name, dtype=torch.float32):
    if t is None:
        return
    if not (t.is_cuda and t.dtype == dtype and t.is_contiguous()):
        raise L.IcadvError(f"{name}: expected a contiguous CUDA {dtype} tensor, got {t.dtype} "
                           f"{'cuda' if t.is_cuda else 'cpu'} contiguous={t.is_contiguous()}")


def require_device():
    if not torch.cuda.is_available():
        raise L.IcadvError("no CUDA device: imagecompression_adversarial_b200 has no CPU path")
    L.call("icadv_check_device")


def pack_weight(w, kind, round_tf32=False):
    """torch Conv2d [Co,Ci,k,k] / ConvTranspose2d [Ci,Co,k,k] weight -> packed [k*k][n_ch][k_ch]."""
    w = w.detach().contiguous()
    _chk(w, "weight")
    k = w.shape[-1]
    if kind in (L.PACK_CONV_FWD, L.PACK_CONV_DGRAD):
        c_out, c_in = w.shape[0], w.shape[1]
    else:
        c_in, c_out = w.shape[0], w.shape[1]
    n, kk = (c_out, c_in) if kind in (L.PACK_CONV_FWD, L.PACK_CONVT_FWD) else (c_in, c_out)
    out = torch.empty(k * k, n, kk, device=w.device, dtype=torch.float32)
    L.call("icadv_pack_weight", _p(w), _p(out), kind, c_out, c_in, k, 1 if round_tf32 else 0, _stream())
    return out


def pack_weight_rgb(w, round_tf32=True):
    """First-layer form: w[n][3][5][5] -> [5 kh][n][32] (k = kw*4 + c)."""
    w = w.detach().contiguous()
    _chk(w, "weight")
    assert tuple(w.shape[1:]) == (3, 5, 5), w.shape
    out = torch.empty(5, w.shape[0], 32, device=w.device, dtype=torch.float32)
    L.call("icadv_pack_weight_rgb", _p(w), _p(out), w.shape[0], 1 if round_tf32 else 0, _stream())
    return out


def alloc_pad4(n_img, h, w, device):
    """Zeroed padded RGB0 buffer [n, h+4, w+8, 4] (borders stay zero; pad_rgb4 fills the interior)."""
    return torch.zeros(n_img, h + 4, w + 8, 4, device=device, dtype=torch.float32)


def pad_rgb4(src, dst, active=None, n_active=None):
    n, h, w, c = src.shape
    assert c == 3 and tuple(dst.shape) == (n, h + 4, w + 8, 4)
    L.call("icadv_pad_rgb4", _p(src), _p(dst), n, h, w, 1, _p(active), _p(n_active), _stream())
    return dst


def unpack_weight_grad(dwpack, like, kind, accumulate_into=None):
    """packed dW [k*k][n][k] -> gradient in the torch layout of ``like``."""
    k = like.shape[-1]
    if kind in (L.PACK_CONV_FWD, L.PACK_CONV_DGRAD):
        c_out, c_in = like.shape[0], like.shape[1]
    else:
        c_in, c_out = like.shape[0], like.shape[1]
    dst = accumulate_into if accumulate_into is not None else torch.empty_like(like)
    L.call("icadv_unpack_weight", _p(dwpack), _p(dst), kind, c_out, c_in, k, 1 if accumulate_into is not None else 0,
           _stream())
    return dst


def nchw_to_nhwc(x):
    x = x.contiguous()
    _chk(x, "x")
    n, c, h, w = x.shape
    out = torch.empty(n, h, w, c, device=x.device, dtype=torch.float32)
    L.call("icadv_nchw_to_nhwc", _p(x), _p(out), n, c, h, w, _stream())
    return out


def nhwc_to_nchw(x):
    _chk(x, "x")
    n, h, w, c = x.shape
    out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    L.call("icadv_nhwc_to_nchw", _p(x), _p(out), n, c, h, w, _stream())
    return out


def clamp01_nhwc_to_nchw(x):
    """Up_bound(Low_bound(x, 0), 1) of a channels-last image batch, written as planes (one launch)."""
    _chk(x, "x")
    n, h, w, c = x.shape
    out = torch.empty(n, c, h, w, device=x.device, dtype=torch.float32)
    L.call("icadv_clamp01_nhwc_to_nchw", _p(x), _p(out), n, c, h, w, _stream())
    return out


def clamp01_backward_nchw_to_nhwc(g_nchw, x_nhwc, out=None):
    """Gradient of ``clamp01_nhwc_to_nchw`` (utils/ops.py:28-56 rules at the unclamped x), written channels-last."""
    _chk(g_nchw, "g"); _chk(x_nhwc, "x")
    n, h, w, c = x_nhwc.shape
    if out is None:
        out = torch.empty_like(x_nhwc)
    _chk(out, "out")
    L.call("icadv_clamp01_backward_nchw_to_nhwc", _p(g_nchw), _p(x_nhwc), _p(out), n, c, h, w, _stream())
    return out


def pixel_shuffle(x_nhwc, r, inverse=False):
    """nn.PixelShuffle(r) on a channels-last tensor [N,H,W,C*r*r] -> [N,H*r,W*r,C] (inverse: the other way)."""
    _chk(x_nhwc, "x")
    n, h, w, c = x_nhwc.shape
    if inverse:
        lo_h, lo_w, c_out = h // r, w // r, c
        out = torch.empty(n, lo_h, lo_w, c * r * r, device=x_nhwc.device, dtype=torch.float32)
    else:
        lo_h, lo_w, c_out = h, w, c // (r * r)
        out = torch.empty(n, h * r, w * r, c_out, device=x_nhwc.device, dtype=torch.float32)
    L.call("icadv_pixel_shuffle", _p(x_nhwc), _p(out), n, lo_h, lo_w, c_out, r, 1 if inverse else 0, _stream())
    return out


def copy_channels(src, dst, src_off, dst_off, count):
    """dst[..., dst_off:dst_off+count] = src[..., src_off:src_off+count] on channels-last tensors of equal pixel count."""
    _chk(src, "src")
    _chk(dst, "dst")
    n_px = src.numel() // src.shape[-1]
    assert dst.numel() // dst.shape[-1] == n_px
    L.call("icadv_copy_channels", _p(src), _p(dst), n_px, src.shape[-1], dst.shape[-1], src_off, dst_off, count, _stream())
    return dst


def gdn_reparam(raw, bound, pedestal, transpose=False, round_tf32=False):
    raw = raw.detach().contiguous()
    _chk(raw, "raw")
    rows, cols = (raw.shape[0], raw.shape[1]) if raw.dim() == 2 else (1, raw.numel())
    out = torch.empty_like(raw)
    L.call("icadv_gdn_reparam", _p(raw), _p(out), rows, cols, float(bound), float(pedestal), 1 if transpose else 0,
           1 if round_tf32 else 0, _stream())
    return out


def out_hw(form, ksize, stride, h, w):
    if form == L.FORM_SCONV:
        p = ksize // 2
        return (h + 2 * p - ksize) // stride + 1, (w + 2 * p - ksize) // stride + 1
    return h * stride, w * stride


def make_desc(x, wpack, bias, out, *, form, ksize, stride, n_ch, epi=L.EPI_LINEAR, act=L.ACT_NONE, gmat=None,
              beta=None, out_scale=None, y_prev=None, sc_prev=None, acc_from_in=False, active=None, n_active=None,
              round_out=False, in_pad4=False):
    if in_pad4:
        n, h, w, k_ch = x.shape[0], x.shape[1] - 4, x.shape[2] - 8, 3
    else:
        n, h, w, k_ch = x.shape
    d = L.ConvDesc()
    d.form, d.ksize, d.stride, d.n_img, d.in_h, d.in_w, d.k_ch, d.n_ch = form, ksize, stride, n, h, w, k_ch, n_ch
    d.inp, d.wpack, d.bias, d.out = _p(x), _p(wpack), _p(bias), _p(out)
    d.epi, d.act = epi, act
    d.gmat, d.beta, d.out_scale, d.y_prev, d.sc_prev = _p(gmat), _p(beta), _p(out_scale), _p(y_prev), _p(sc_prev)
    d.acc_from_in = 1 if acc_from_in else 0
    d.round_out_tf32 = 1 if round_out else 0
    d.in_pad4 = 1 if in_pad4 else 0
    d.active, d.n_active = _p(active), _p(n_active)
    return d


def conv(x, wpack, bias=None, *, form, ksize, stride, n_ch, epi=L.EPI_LINEAR, act=L.ACT_NONE, gmat=None, beta=None,
         y_prev=None, sc_prev=None, acc_from_in=False, active=None, n_active=None, out=None, out_scale=None,
         path="auto", round_out=False, in_pad4=False):
    """One contraction launch (see include/icadv.h).  Returns ``out`` or ``(out, out_scale)`` for the
    GDN/IGDN forward epilogues.  ``path``: "auto" (tensor path when the shape allows), "tc", "simt"."""
    for t, nm in ((x, "x"), (wpack, "wpack"), (bias, "bias"), (gmat, "gmat"), (beta, "beta"), (y_prev, "y_prev"),
                  (sc_prev, "sc_prev"), (out, "out"), (out_scale, "out_scale")):
        _chk(t, nm)
    _chk(active, "active", torch.int32)
    _chk(n_active, "n_active", torch.int32)
    n, h, w, _ = x.shape
    if in_pad4:
        h, w = h - 4, w - 8
    oh, ow = out_hw(form, ksize, stride, h, w)
    if out is None:
        if form == L.FORM_TCONV and ksize < stride:
            # output pixels no tap reaches (the input gradient of a strided 1x1 conv: 3 of 4 pixels) are not written
            # by the kernels: they are the bias, or zero.  torch.empty() is NOT zero (under the reference's
            # torch.use_deterministic_algorithms(True), self_ensemble.py:31, it is NaN-filled).
            out = torch.zeros(n, oh, ow, n_ch, device=x.device, dtype=torch.float32)
            if bias is not None:
                out += bias
        else:
            out = torch.empty(n, oh, ow, n_ch, device=x.device, dtype=torch.float32)
    fwd_gdn = epi in (L.EPI_GDN_FWD, L.EPI_IGDN_FWD)
    if fwd_gdn and out_scale is None:
        out_scale = torch.empty_like(out)
    d = make_desc(x, wpack, bias, out, form=form, ksize=ksize, stride=stride, n_ch=n_ch, epi=epi, act=act, gmat=gmat,
                  beta=beta, out_scale=out_scale, y_prev=y_prev, sc_prev=sc_prev, acc_from_in=acc_from_in,
                  active=active, n_active=n_active, round_out=round_out, in_pad4=in_pad4)
    use_tc = path == "tc" or (path == "auto" and L.lib().icadv_conv_tc_supported(C.byref(d)) == 1)
    if use_tc:
        L.call("icadv_conv_tc", C.byref(d), _stream())
    else:
        if epi != L.EPI_LINEAR or acc_from_in:
            raise L.IcadvError("GDN epilogues need the tensor path (k_ch, n_ch multiples of 32)")
        L.call("icadv_conv_simt", C.byref(d), _stream())
    return (out, out_scale) if fwd_gdn else out


class ConvPlan:
    """Cached tensor maps + launch geometry for fixed buffers (icadv_conv_plan_*)."""

    def __init__(self, desc, keep):
        self._keep = keep  # tensors whose pointers are baked into the tensor maps
        self.desc = desc
        h = C.c_void_p()
        L.call("icadv_conv_plan_create", C.byref(desc), C.byref(h))
        self._h = h
        self.kernels = L.lib().icadv_conv_plan_num_launches(h)

    def launch(self):
        L.call("icadv_conv_plan_launch", self._h, _stream())

    def __del__(self):
        try:
            if getattr(self, "_h", None):
                L.lib().icadv_conv_plan_destroy(self._h)
        except Exception:
            pass


class SimtLaunch:
    """Same interface as ConvPlan for the CUDA-core path (no cached state beyond the descriptor)."""

    def __init__(self, desc, keep):
        self._keep, self._d = keep, desc
        self.kernels = desc.stride * desc.stride if desc.form == L.FORM_TCONV else 1

    def launch(self):
        L.call("icadv_conv_simt", C.byref(self._d), _stream())


_WGRAD_PATH = {"auto": 0, "simt": 1, "tc": 2}


def conv_wgrad(x, gout, *, form, ksize, stride, n_ch, want_bias=True, path="auto"):
    """dW (packed) and dbias of a contraction whose forward input was ``x`` and output gradient ``gout``.
    ``path``: "tc" = tcgen05 kernel (pixel axis as the reduction axis; operands are read as TF32, so callers round them),
    "simt" = fp32 CUDA-core kernel, "auto" = tensor path where the shape is eligible (channel counts multiples of 32)."""
    _chk(x, "x")
    _chk(gout, "gout")
    k_ch = x.shape[-1]
    dw = torch.empty(ksize * ksize, n_ch, k_ch, device=x.device, dtype=torch.float32)
    db = torch.empty(n_ch, device=x.device, dtype=torch.float32) if want_bias else None
    d = make_desc(x, dw, None, gout, form=form, ksize=ksize, stride=stride, n_ch=n_ch)
    L.call("icadv_conv_wgrad_ex", C.byref(d), _p(gout), _p(dw), _p(db), _WGRAD_PATH[path], _stream())
    return dw, db


def conv_wgrad_tc_supported(k_ch, n_ch, ksize, stride, form=L.FORM_SCONV, in_hw=(2, 2)):
    """Whether icadv_conv_wgrad takes the tcgen05 kernel for this contraction (channel counts multiples of 32, or one of
    the two RGB end-layer forms: 3 -> N conv / N -> 3 transposed conv, 5x5 stride 2)."""
    d = L.ConvDesc()
    d.form, d.k_ch, d.n_ch, d.ksize, d.stride, d.in_h, d.in_w = form, k_ch, n_ch, ksize, stride, in_hw[0], in_hw[1]
    return bool(L.lib().icadv_conv_wgrad_tc_supported(C.byref(d)))


# ------------------------------------------------------------------ perturbation step
class PerturbState:
    """Device-resident per-image loop state (include/icadv.h icadv_perturb_state)."""

    def __init__(self, n_img, device):
        f = lambda: torch.zeros(n_img, device=device, dtype=torch.float32)
        i = lambda n=n_img: torch.zeros(n, device=device, dtype=torch.int32)
        self.n_img = n_img
        self.sum_d2, self.loss_i, self.lr, self.step_size, self.bc2_sqrt = f(), f(), f(), f(), f()
        self.branch, self.active, self.step, self.n_active = i(), i(), i(), i(1)
        self.ws = torch.zeros(n_img * L.RED_BLOCKS, device=device, dtype=torch.float32)
        self.counter = i(1)             # arrival counter of perturb_forward's blocks (reset by the kernel itself)
        s = L.PerturbState()
        for name in ("sum_d2", "loss_i", "branch", "active", "n_active", "step", "lr", "step_size", "bc2_sqrt", "counter"):
            setattr(s, name, _p(getattr(self, name)))
        s.cond_handle = 0               # set by the engine while it captures the loop's IF node
        self.c = s


def perturb_forward(im_s, noise, im_in, st, *, eps, budget, force_branch=-1, lr0=0.01, lr_gamma=0.33, sched_period,
                    beta1=0.9, beta2=0.999, w_in=None, ge_test=False, batch_budget=False):
    """``batch_budget``: one budget test on the batch mean of loss_i (attack_rd.py:333 on a batch, train.py:342)."""
    per_img = im_s[0].numel()
    L.call("icadv_perturb_forward_roi", _p(im_s), _p(noise), _p(im_in), _p(st.ws), C.byref(st.c), st.n_img, per_img,
           float(eps), float(budget), int(force_branch), float(lr0), float(lr_gamma), int(sched_period), float(beta1),
           float(beta2), _p(w_in), (1 if ge_test else 0) | (2 if batch_budget else 0), _stream())


def perturb_update_adam(im_s, noise, g_in, m, v, st, *, eps, beta1=0.9, beta2=0.999, adam_eps=1e-8, gradA_scale,
                        gradB_scale=1.0, g_a_ext=None, w_in=None):
    per_img = im_s[0].numel()
    L.call("icadv_perturb_update_adam_roi", _p(im_s), _p(noise), _p(g_in), _p(g_a_ext), _p(m), _p(v), C.byref(st.c),
           st.n_img, per_img, float(eps), float(beta1), float(beta2), float(adam_eps), float(gradA_scale),
           float(gradB_scale), _p(w_in), _stream())


def output_loss(x, ref, g_x, ws, sum_d2, *, do_clamp, grad_scale, active=None, n_active=None, w_out=None):
    n_img, per_img = x.shape[0], x[0].numel()
    L.call("icadv_output_loss_roi", _p(x), _p(ref), _p(g_x), _p(ws), _p(sum_d2), n_img, per_img, 1 if do_clamp else 0,
           float(grad_scale), _p(active), _p(n_active), _p(w_out), _stream())


def sum_sqdiff(a, b):
    n_img, per_img = a.shape[0], a[0].numel()
    ws = torch.empty(n_img * L.RED_BLOCKS, device=a.device, dtype=torch.float32)
    out = torch.empty(n_img, device=a.device, dtype=torch.float32)
    L.call("icadv_sum_sqdiff", _p(a), _p(b), _p(ws), _p(out), n_img, per_img, _stream())
    return out


def ifgsm_update(im_s, im_adv, g, *, alpha, eps):
    L.call("icadv_ifgsm_update", _p(im_s), _p(im_adv), _p(g), im_s.numel(), float(alpha), float(eps), _stream())


def cw_combine(g_net, im_in, im_s, g_out, c, level, sum_d2):
    """C&W gradient wrt im_in (attack_cw.py:111-140): 2 (im_in - im_s) / P + c_n g_net, c_n zeroed per image once its
    reconstruction error exceeds 1.1 x its target level."""
    n_img, per_img = im_s.shape[0], im_s[0].numel()
    L.call("icadv_cw_combine", _p(g_net), _p(im_in), _p(im_s), _p(g_out), _p(c), _p(level), _p(sum_d2), n_img, per_img,
           _stream())


def mifgsm_update(im_s, im_adv, g, g_mom, ws, l1, *, alpha, eps, mu=1.0):
    n_img, per_img = im_s.shape[0], im_s[0].numel()
    L.call("icadv_mifgsm_update", _p(im_s), _p(im_adv), _p(g), _p(g_mom), _p(ws), _p(l1), n_img, per_img, float(alpha),
           float(eps), float(mu), _stream())


def bound_forward(x, bound, upper):
    y = torch.empty_like(x)
    L.call("icadv_bound_forward", _p(x), _p(y), x.numel(), float(bound), 1 if upper else 0, _stream())
    return y


def bound_backward(x, gy, bound, upper):
    gx = torch.empty_like(x)
    L.call("icadv_bound_backward", _p(x), _p(gy), _p(gx), x.numel(), float(bound), 1 if upper else 0, _stream())
    return gx


# ------------------------------------------------------------------ elementwise helpers
_UNARY = {L.ACT_ABS: 0, L.ACT_RELU: 1, L.ACT_LEAKY: 2}


def unary(x, op, b=None):
    """op 0 abs, 1 relu, 2 leaky(0.01), 3 round, 4 x + b.  Works on the raw storage order of ``x``."""
    y = torch.empty_like(x)
    if b is not None and b.stride() != x.stride():
        raise L.IcadvError("unary: operands must share a memory layout")
    L.call("icadv_unary", _p(x), _p(b), _p(y), x.numel(), int(op), _stream())
    return y


def attention_gate(a, b, x):
    """a * sigmoid(b) + x on three tensors sharing shape and memory layout."""
    if not (a.stride() == b.stride() == x.stride() and a.shape == b.shape == x.shape):
        raise L.IcadvError("attention_gate: operands must share shape and memory layout")
    y = torch.empty_like(a)
    L.call("icadv_attention_gate", _p(a), _p(b), _p(x), _p(y), a.numel(), _stream())
    return y


def attention_gate_backward(a, b, g):
    if not (a.stride() == b.stride() == g.stride() and a.shape == g.shape):
        raise L.IcadvError("attention_gate_backward: operands must share shape and memory layout")
    ga, gb = torch.empty_like(a), torch.empty_like(a)
    L.call("icadv_attention_gate_backward", _p(a), _p(b), _p(g), _p(ga), _p(gb), a.numel(), _stream())
    return ga, gb


def act_backward(x, g, act):
    if x.stride() != g.stride() or x.shape != g.shape:
        raise L.IcadvError("act_backward: operands must share shape and memory layout")
    gx = torch.empty_like(g)
    L.call("icadv_act_backward", _p(x), _p(g), _p(gx), x.numel(), _UNARY[act], _stream())
    return gx


# ------------------------------------------------------------------ entropy models
def eb_prepare(matrices, biases, factors, C):
    """raw EntropyBottleneck parameters -> prepared [58][C] table (softplus / tanh hoisted)."""
    dev = matrices[0].device
    table = torch.empty(58, C, device=dev, dtype=torch.float32)
    keep = [t.detach().contiguous() for t in list(matrices) + list(biases) + list(factors)]
    arr = lambda ts: (L._fp * len(ts))(*[t.data_ptr() for t in ts])
    L.call("icadv_eb_prepare", arr(keep[0:5]), arr(keep[5:10]), arr(keep[10:14]), _p(table), C, _stream())
    return table


def eb_forward(x_nhwc, table, medians, *, training, noise=None, lik_bound=1e-9, bits_floor=0.0):
    """x_nhwc [N,...,C] channels-last.  Returns (x_hat, likelihood, bits[N])."""
    _chk(x_nhwc, "x")
    n_img, C = x_nhwc.shape[0], x_nhwc.shape[-1]
    per_img = x_nhwc[0].numel()
    x_hat, lik = torch.empty_like(x_nhwc), torch.empty_like(x_nhwc)
    ws = torch.empty(n_img * L.RED_BLOCKS, device=x_nhwc.device, dtype=torch.float32)
    bits = torch.empty(n_img, device=x_nhwc.device, dtype=torch.float32)
    L.call("icadv_eb_forward", _p(x_nhwc), _p(noise), _p(table), _p(medians), _p(x_hat), _p(lik), _p(ws), _p(bits),
           n_img, per_img, C, 1 if training else 0, float(lik_bound), float(bits_floor), _stream())
    return x_hat, lik, bits


def gc_forward(y, scales, means=None, *, training, noise=None, scale_bound=0.11, lik_bound=1e-9, bits_floor=0.0):
    """Elementwise over identically laid-out tensors.  Returns (y_hat, likelihood, bits[N])."""
    for t, nm in ((y, "y"), (scales, "scales"), (means, "means"), (noise, "noise")):
        _chk(t, nm)
    n_img, per_img = y.shape[0], y[0].numel()
    y_hat, lik = torch.empty_like(y), torch.empty_like(y)
    ws = torch.empty(n_img * L.RED_BLOCKS, device=y.device, dtype=torch.float32)
    bits = torch.empty(n_img, device=y.device, dtype=torch.float32)
    L.call("icadv_gc_forward", _p(y), _p(scales), _p(means), _p(noise), _p(y_hat), _p(lik), _p(ws), _p(bits), n_img,
           per_img, 1 if training else 0, float(scale_bound), float(lik_bound), float(bits_floor), _stream())
    return y_hat, lik, bits


# ------------------------------------------------------------------ codec update of adversarial training (train.py --adv)
def eb_backward(x_hat, g_lik, table, matrices, factors, *, lik_bound=1e-9):
    """(g_x, g_raw[58][C]) of the train-mode EntropyBottleneck likelihood; x_hat, g_lik channels-last [..., C]."""
    _chk(x_hat, "x_hat")
    _chk(g_lik, "g_lik")
    Cc = x_hat.shape[-1]
    rows = x_hat.numel() // Cc
    g_x = torch.empty_like(x_hat)
    g_raw = torch.empty(58, Cc, device=x_hat.device, dtype=torch.float32)
    ws = torch.empty(L.lib().icadv_eb_backward_workspace_floats(rows, Cc), device=x_hat.device, dtype=torch.float32)
    keep = [t.detach().contiguous() for t in list(matrices) + list(factors)]
    arr = lambda ts: (L._fp * len(ts))(*[t.data_ptr() for t in ts])
    L.call("icadv_eb_backward", _p(x_hat), _p(g_lik), _p(table), arr(keep[0:5]), arr(keep[5:9]), _p(g_x), _p(g_raw),
           _p(ws), rows, Cc, float(lik_bound), _stream())
    return g_x, g_raw


def gc_backward(y_hat, scales, means, g_lik, *, scale_bound=0.11, lik_bound=1e-9):
    for t, nm in ((y_hat, "y_hat"), (scales, "scales"), (means, "means"), (g_lik, "g_lik")):
        _chk(t, nm)
    g_y, g_s = torch.empty_like(y_hat), torch.empty_like(y_hat)
    g_m = torch.empty_like(y_hat) if means is not None else None
    L.call("icadv_gc_backward", _p(y_hat), _p(scales), _p(means), _p(g_lik), _p(g_y), _p(g_s), _p(g_m), y_hat.numel(),
           float(scale_bound), float(lik_bound), _stream())
    return g_y, g_s, g_m


def log_sum(lik, floor):
    """sum log(max(lik, floor)) as a 1-element device tensor (any memory layout: the sum is order-free per block)."""
    _chk_dense(lik)
    ws = torch.empty(L.RED_BLOCKS, device=lik.device, dtype=torch.float32)
    out = torch.empty(1, device=lik.device, dtype=torch.float32)
    L.call("icadv_log_sum", _p(lik), _p(ws), _p(out), lik.numel(), float(floor), _stream())
    return out


def log_sum_backward(lik, floor, scale_dev, scale_host=1.0):
    _chk_dense(lik)
    g = torch.empty_like(lik)
    L.call("icadv_log_sum_backward", _p(lik), _p(g), lik.numel(), float(floor), _p(scale_dev), float(scale_host), _stream())
    return g


def scaled_diff(a, b, scale_dev, scale_host=1.0):
    _chk_dense(a)
    if a.stride() != b.stride() or a.shape != b.shape:
        raise L.IcadvError("scaled_diff: operands must share shape and memory layout")
    out = torch.empty_like(a)
    L.call("icadv_scaled_diff", _p(a), _p(b), _p(out), a.numel(), _p(scale_dev), float(scale_host), _stream())
    return out


def _chk_dense(t):
    if not (t.is_cuda and t.dtype == torch.float32 and
            (t.is_contiguous() or t.is_contiguous(memory_format=torch.channels_last))):
        raise L.IcadvError("expected a dense CUDA float32 tensor")


GDN_PARAM_GRAD_TC = True   # False: the fp32 CUDA-core kernel everywhere (A/B tests)


def gdn_param_grad(g, y, sc, beta_raw, gamma_raw, *, inverse, beta_bound, gamma_bound):
    """(d beta_raw [C], d gamma_raw [C,C]) from g = dL/dy, saved y and scale (channels-last)."""
    for t, nm in ((g, "g"), (y, "y"), (sc, "sc")):
        _chk(t, nm)
    Cc = y.shape[-1]
    gb = torch.empty(Cc, device=y.device, dtype=torch.float32)
    gg = torch.empty(Cc, Cc, device=y.device, dtype=torch.float32)
    from . import precision
    if (GDN_PARAM_GRAD_TC and not precision.split() and y.dim() == 4 and g.shape == y.shape == sc.shape and
            conv_wgrad_tc_supported(Cc, Cc, 1, 1, in_hw=(y.shape[1], y.shape[2]))):
        # tensor path: d gamma = the 1x1 weight gradient of (input x^2, output gradient -+1/2 t), d beta its bias gradient
        T, X2 = torch.empty_like(y), torch.empty_like(y)
        L.call("icadv_gdn_param_operands", _p(g), _p(y), _p(sc), _p(T), _p(X2), y.numel(), 1 if inverse else 0, _stream())
        dw, db = conv_wgrad(X2, T, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=Cc, want_bias=True, path="tc")
        L.call("icadv_gdn_param_grad_finalize", _p(dw), _p(db), _p(beta_raw.detach().contiguous()),
               _p(gamma_raw.detach().contiguous()), _p(gb), _p(gg), Cc, float(beta_bound), float(gamma_bound), _stream())
        return gb, gg
    ws = torch.empty(L.lib().icadv_gdn_param_grad_workspace_floats(Cc), device=y.device, dtype=torch.float32)
    L.call("icadv_gdn_param_grad", _p(g), _p(y), _p(sc), _p(beta_raw.detach().contiguous()),
           _p(gamma_raw.detach().contiguous()), _p(gb), _p(gg), _p(ws), y.numel() // Cc, Cc, 1 if inverse else 0,
           float(beta_bound), float(gamma_bound), _stream())
    return gb, gg


def sumsq(flat):
    ws = torch.empty(L.RED_BLOCKS, device=flat.device, dtype=torch.float32)
    out = torch.empty(1, device=flat.device, dtype=torch.float32)
    L.call("icadv_sumsq", _p(flat), _p(ws), _p(out), flat.numel(), _stream())
    return out


def adam_clip_step(params, grads, m, v, *, sumsq_dev, max_norm, lr, step, grad_scale=1.0, beta1=0.9, beta2=0.999,
                   eps=1e-8):
    L.call("icadv_adam_clip_step", _p(params), _p(grads), _p(m), _p(v), params.numel(), _p(sumsq_dev), float(max_norm),
           float(grad_scale), float(lr), float(beta1), float(beta2), float(eps), int(step), _stream())


# ------------------------------------------------------------------ 3xTF32 parity mode (csrc/icadv_split.cu)
def split_geom(c, form=None, ksize=1, stride=1):
    """(Ks, G) of the K-sliced split form of a C-channel operand feeding a contraction with this geometry: G slices of
    Ks channels, Ks * G >= 3C, with at most ~128 sequential tensor-core accumulations per slice (taps x Ks/8) -- the
    accumulator truncates after every instruction, so the chain length bounds the bias (csrc/icadv_split.cu)."""
    if form == L.FORM_TCONV:
        taps = ((ksize + stride - 1) // stride) ** 2       # taps per output-parity class
    else:
        taps = ksize * ksize
    chunks = max(1, 32 // taps)
    total = (3 * c + 31) // 32                              # 32-channel chunks of the whole split form
    chunks = min(chunks, total)
    g = (total + chunks - 1) // chunks
    return 32 * chunks, g


def split3(x, ks, g, op=0, out=None):
    """[..., C] channels-last activation -> [G, ..., Ks] slices of [hi | lo | hi | 0]; op 1 squares first."""
    _chk(x, "x")
    c = x.shape[-1]
    if out is None:
        out = torch.empty(g, *x.shape[:-1], ks, device=x.device, dtype=torch.float32)
    assert out.numel() == x.numel() // c * ks * g, (out.shape, x.shape, ks, g)
    L.call("icadv_split3", _p(x), _p(out), x.numel() // c, c, ks, g, int(op), 0, _stream())
    return out


def split3_weight(wpack, ks, g, out=None):
    """packed weight [taps][n][k] (or [1][n][k] for a matrix) -> [G][taps][n][Ks] slices of [hi | hi | lo | 0]."""
    _chk(wpack, "wpack")
    k = wpack.shape[-1]
    if out is None:
        out = torch.empty(g, *wpack.shape[:-1], ks, device=wpack.device, dtype=torch.float32)
    L.call("icadv_split3", _p(wpack), _p(out), wpack.numel() // k, k, ks, g, 0, 1, _stream())
    return out


def gdn_bwd_operand_split3(g_out, y, sc, inverse, ks, g, out=None):
    c = y.shape[-1]
    if out is None:
        out = torch.empty(g, *y.shape[:-1], ks, device=y.device, dtype=torch.float32)
    L.call("icadv_gdn_bwd_operand_split3", _p(g_out), _p(y), _p(sc), _p(out), y.numel() // c, c, ks, g,
           1 if inverse else 0, _stream())
    return out


def sum_slices(parts, out):
    """out = parts[0] + parts[1] + ... in round-to-nearest fp32 (parts: [G, ...] contiguous)."""
    L.call("icadv_sum_slices", _p(parts), _p(out), out.numel(), parts.shape[0], _stream())
    return out


def gdn_apply(x, nrm, y, sc, inverse):
    L.call("icadv_gdn_apply", _p(x), _p(nrm), _p(y), _p(sc), x.numel(), 1 if inverse else 0, _stream())


def gdn_bwd_combine(g, y, sc, w, out, inverse):
    L.call("icadv_gdn_bwd_combine", _p(g), _p(y), _p(sc), _p(w), _p(out), y.numel(), 1 if inverse else 0, _stream())


def conv_sliced(xs, ws, bias, *, form, ksize, stride, n_ch, act=L.ACT_NONE):
    """Parity-mode contraction from prepared slices: xs [G, N, H, W, Ks], ws [G, taps, n_ch, Ks].  One tensor-path
    launch per slice (the bias rides in slice 0), partial outputs added in fp32, activation applied after the sum."""
    g = xs.shape[0]
    n, h, w = xs.shape[1], xs.shape[2], xs.shape[3]
    oh, ow = out_hw(form, ksize, stride, h, w)
    if form == L.FORM_TCONV and ksize < stride:
        parts = torch.zeros(g, n, oh, ow, n_ch, device=xs.device, dtype=torch.float32)
        if bias is not None:
            parts[0] += bias
    else:
        parts = torch.empty(g, n, oh, ow, n_ch, device=xs.device, dtype=torch.float32)
    for i in range(g):
        conv(xs[i], ws[i], bias if i == 0 else None, form=form, ksize=ksize, stride=stride, n_ch=n_ch, out=parts[i],
             path="tc")
    out = parts[0] if g == 1 else sum_slices(parts, torch.empty_like(parts[0]))
    if act != L.ACT_NONE:
        out = unary(out, _UNARY[act])
    return out


def gdn_forward_split(xn, beta_eff, gamma_eff, inverse):
    """Unfused GDN / IGDN in the parity mode: (y, sc).  gamma_eff [C][C] unrounded."""
    c = xn.shape[-1]
    ks, g = split_geom(c)
    sq = split3(xn, ks, g, op=1)
    nrm = conv_sliced(sq, split3_weight(gamma_eff.contiguous().view(1, c, c), ks, g), beta_eff.contiguous(),
                      form=L.FORM_SCONV, ksize=1, stride=1, n_ch=c)
    y, sc = torch.empty_like(xn), torch.empty_like(xn)
    gdn_apply(xn, nrm, y, sc, inverse)
    return y, sc


def gdn_backward_split(gn, y, sc, gamma_eff, inverse):
    c = y.shape[-1]
    ks, g = split_geom(c)
    t = gdn_bwd_operand_split3(gn, y, sc, inverse, ks, g)
    w = conv_sliced(t, split3_weight(gamma_eff.t().contiguous().view(1, c, c), ks, g), None, form=L.FORM_SCONV, ksize=1,
                    stride=1, n_ch=c)
    out = torch.empty_like(y)
    gdn_bwd_combine(gn, y, sc, w, out, inverse)
    return out


def uniform_noise_like(x, lo=-0.5, hi=0.5):
    """U[lo, hi) sample shaped (and laid out) like ``x`` from the library's Philox kernel.  Seed and counter are those
    of torch's CUDA generator of the device (``torch.manual_seed`` restarts the stream, every draw advances it), so
    runs reproduce exactly like runs that draw from torch."""
    out = torch.empty_like(x)
    n = out.numel()
    gen = torch.cuda.default_generators[x.device.index if x.device.index is not None else torch.cuda.current_device()]
    seed, offset = gen.initial_seed() & ((1 << 64) - 1), gen.get_offset()
    L.call("icadv_uniform_noise", _p(out), n, C.c_uint64(seed), C.c_uint64(offset // 4), float(lo), float(hi), _stream())
    gen.set_offset(offset + 4 * ((n + 3) // 4))      # torch keeps the Philox offset in units of 4 (one 128-bit block)
    return out


def probe_tf32_peak(iters=4000, n=256, reps=3):
    """Measured kind::tf32 tensor-core rate of this GPU in TFLOP/s (csrc/icadv_probe.cu): best of ``reps`` launches of a
    bare MMA loop, each timed with CUDA events.  ``iters`` sets the launch length (4000 K-blocks ~ 1 ms)."""
    flops = C.c_double(0.0)
    best = None
    for _ in range(reps + 1):                       # first launch: warm-up
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        L.call("icadv_probe_tf32_peak", int(iters), int(n), C.byref(flops), _stream())
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)
        best = ms if best is None else min(best, ms)
    return flops.value / (best * 1e-3) / 1e12
