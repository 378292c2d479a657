"""Drop-in for the reference's ``attack_`` (attack_rd.py:381-575) and ``self_ensemble.eval``
(self_ensemble.py:173-252) on the fused device-resident loop.

``attack_(im_s, net, args)`` keeps the reference's signature and return tuple.  Differences, all stated:
  * a batch is attacked with PER-IMAGE budget tests by default (each image behaves like an N=1 call = the CLI path,
    which attacks one image at a time); ``args.budget_scope = "batch"`` selects the reference's literal batch semantics
    (loss_i / loss_o are means over the batch, one shared branch per iteration: attack_rd.py:333-364 on a batch), which
    is what ``training.adv_train_step`` uses because train.py:342 hands ``attack_`` a whole training batch;
  * no host synchronisation inside the loop (the reference syncs every step at attack_rd.py:334).
"""
import math

import torch

from . import metrics
from . import _lib as L
from . import ops
from . import precision
from .engine import AttackEngine, GenericAttackEngine, IfgsmEngine, RoiSpec, TapeAttackEngine
from .program import parse_stack

_ENGINES = {}


# The scalar metrics of attack_rd.py:402-419 and self_ensemble.py:173-252 through the library's fixed-order reductions
# (no torch arithmetic on image-sized tensors; what remains in Python is scalar bookkeeping).
def _clamp01(x):
    return ops.unary(x.contiguous(memory_format=torch.channels_last), 6)


def _mse(a, b):
    """mean((a - b)^2) over everything, as a Python float (these values are printed / returned as floats)."""
    ac = a.contiguous(memory_format=torch.channels_last)
    bc = b.contiguous(memory_format=torch.channels_last)
    flat = lambda t: t.permute(0, 2, 3, 1).reshape(1, -1)
    return float(ops.sum_sqdiff(flat(ac), flat(bc))[0]) / ac.numel()


def _bpp(likelihoods, num_pixels):
    """sum_k sum log(lik_k) / (-ln 2 * num_pixels)  (attack_rd.py:419, self_ensemble.py:222), 0-dim device tensor."""
    total = None
    for lik in likelihoods.values():
        lk = lik if (lik.is_contiguous() or lik.is_contiguous(memory_format=torch.channels_last)) else lik.contiguous()
        t = ops.log_sum(lk, 0.0)
        total = t if total is None else ops.unary(total, 4, t)
    return (total * (1.0 / (-math.log(2) * num_pixels))).reshape(())


def _fused_stacks(net):
    try:
        parse_stack(net.g_a)
        parse_stack(net.g_s)
        return True
    except L.IcadvError:
        return False


def _engine_for(net, im_s, args, roi=None):
    n, _, h, w = im_s.shape
    force = getattr(args, "force_branch", -1)
    scope = getattr(args, "budget_scope", "image")
    key = (id(net), n, h, w, args.steps, float(args.epsilon), float(args.noise), float(args.lr_attack),
           bool(args.clamp), args.att_metric, force, roi.key() if roi is not None else None, precision.get(), scope)
    eng = _ENGINES.get(key)
    if eng is None:
        # plain conv/GDN stacks: fused launch programs; anything else: traced into a static launch program (speed mode) or
        # walked module by module through autograd (parity mode: every contraction K-sliced and split, precision.py)
        cls = AttackEngine if _fused_stacks(net) else (GenericAttackEngine if precision.split() else TapeAttackEngine)
        eng = cls(net, n, h, w, steps=args.steps, epsilon=args.epsilon, noise_budget=args.noise,
                  lr_attack=args.lr_attack, clamp=args.clamp, att_metric=args.att_metric, force_branch=force, roi=roi,
                  budget_scope=scope)
        _ENGINES.clear()  # one live engine: its buffers are sized for the batch
        _ENGINES[key] = eng
    else:
        eng.refresh_parameters()
    return eng


@torch.no_grad()
def clean_pass(im_s, net, args):
    """attack_rd.py:402-419."""
    net.eval()
    result = net(im_s)
    output_s = _clamp01(result["x_hat"]) if args.clamp else result["x_hat"]
    num_pixels = im_s.shape[2] * im_s.shape[3]
    bpp_ori = _bpp(result["likelihoods"], num_pixels)
    return output_s, bpp_ori


@torch.no_grad()
def eval(im_adv, im_s, output_s, net, args):  # noqa: A001 (the reference shadows the builtin too)
    """self_ensemble.py:173-252 (defence branches out of scope)."""
    net.eval()
    im_ = _clamp01(im_adv) if args.clamp else im_adv
    result = net(im_)
    x_hat = result["x_hat"]
    mse_in = _mse(im_, im_s)
    output_ = _clamp01(x_hat) if args.clamp else x_hat
    num_pixels = im_adv.shape[2] * im_adv.shape[3]
    bpp = _bpp(result["likelihoods"], num_pixels)
    mse_out = _mse(output_, output_s)
    mse_results = {"mse_in": mse_in, "mse_out": mse_out}
    vi_results = {"vi": None, "vi_msim": None}
    msim_out = metrics.ms_ssim(output_, output_s, data_range=1.0).item()
    msim_in = metrics.ms_ssim(im_, im_s, data_range=1.0).item()
    if mse_in > 1e-20 and mse_out > 1e-20:
        vi_results["vi"] = 10.0 * math.log10(mse_out / mse_in)
        if not getattr(args, "adv", False) and msim_in < 0.9999:
            vi_results["vi_msim"] = 10.0 * math.log10((1 - msim_out) / (1 - msim_in))
    else:
        print(f"[!] Warning: mse_in ({mse_in}) or mse_out {mse_out} is zero")
    return im_, output_, bpp, mse_results, vi_results


def attack_(im_s, net, args, record=None, noise_init=None, im_t=None, metrics_pass=True):
    """Same contract as attack_rd.attack_: returns
    (im_adv, output_adv, output_s, bpp_ori, bpp, mse_results, vi_results).
    ``metrics_pass=False`` (used by ``training.adv_train_step``, which like train.py:342 consumes element [0] only) skips
    the final evaluation pass of self_ensemble.eval and its host reads: the call then returns without synchronising, so
    the caller's next launches queue up behind the attack, and elements [1], [4], [5], [6] of the tuple are None.
    ``im_t`` (the ``-t`` target image, same shape as ``im_s``) switches to the targeted / ROI loss with ``args.mask_loc``,
    ``args.lamb_bkg_in``, ``args.lamb_bkg_out``, ``args.lamb_tar`` (semantics: oracle/attack.py attack_our_roi; the
    reference's own path for these flags is dead code)."""
    output_s, bpp_ori = clean_pass(im_s, net, args)
    roi = output_t = None
    if im_t is not None:
        output_t, _ = clean_pass(im_t, net, args)
        roi = RoiSpec(getattr(args, "mask_loc", None), getattr(args, "lamb_bkg_in", 1.0),
                      getattr(args, "lamb_bkg_out", 1.0), getattr(args, "lamb_tar", 1.0))
    if noise_init is None and getattr(args, "random", 1) > 1:
        noise_init = torch.empty_like(im_s).uniform_(-1e-2, 1e-2)          # attack_rd.py:498-499
    net.train()                                                            # attack_rd.py:504
    eng = _engine_for(net, im_s, args, roi)
    eng.load(im_s, output_s, noise_init, output_t)
    eng.run(args.steps, record=record)
    im_in = eng.im_in_nchw().contiguous()
    if not metrics_pass:
        im_adv = _clamp01(im_in) if args.clamp else im_in                  # what eval returns as im_ (self_ensemble.py:186)
        return im_adv, None, output_s, bpp_ori, None, None, None
    im_adv, output_adv, bpp, mse_results, vi_results = eval(im_in, im_s, output_s, net, args)
    return im_adv, output_adv, output_s, bpp_ori, bpp, mse_results, vi_results


def attack_ifgsm(im_s, net, args, random_start=False, multi_start=1, momentum=False, record=None, start=None):
    """Same contract as attack_ifgsm.attack_ifgsm (attack_ifgsm.py:364-438), single start:
    returns (im_adv, output_adv, output_s, bpp_ori, bpp, mse_in, mse_out, vi)."""
    output_s, bpp_ori = clean_pass(im_s, net, args)
    eps = args.epsilon / 255.0
    if start is None and random_start:
        start = torch.clamp(im_s + torch.empty_like(im_s).uniform_(-eps, eps), 0, 1)    # attack_ifgsm.py:381-383
    net.train()
    n, _, h, w = im_s.shape
    eng = IfgsmEngine(net, n, h, w, steps=args.steps, epsilon=args.epsilon, momentum=momentum)
    eng.load(im_s, output_s, start)
    eng.run(args.steps, record=record)
    im_adv = eng.im_adv_nchw().contiguous()
    im_, output_adv, bpp, mse_results, vi_results = eval(im_adv, im_s, output_s, net, args)
    return im_, output_adv, output_s, bpp_ori, bpp, mse_results["mse_in"], mse_results["mse_out"], vi_results["vi"]


def attack_cw(im_s, net, args, record=None):
    """Same contract as attack_cw.attack_ (attack_cw.py:194-263; inner search :142-192, loss :111-140): a bisection on
    the reconstruction-error level wrapped around a bisection on the loss weight c, each probe ``args.steps`` Adam
    iterations through the network.  Returns (im_adv, output_adv, output_s, bpp_ori, bpp, mse_in, mse_out, vi).

    The reference handles one image per call; here every image of the batch carries its own (level, c) search state
    (host arrays), the iterations of all images run as one fused batch, and one device->host read of 2 floats per image
    closes each block of ``args.steps`` iterations (the reference synchronises every iteration).  An image whose outer
    search has converged (:248-249) keeps the adversarial input it had at that point."""
    import numpy as np
    from .engine import CwEngine
    output_s, bpp_ori = clean_pass(im_s, net, args)
    net.train()                                                            # attack_cw.py:231
    n, _, h, w = im_s.shape
    key = ("cw", id(net), n, h, w, float(args.epsilon), float(args.lr_attack), bool(args.clamp), precision.get())
    eng = _ENGINES.get(key)
    if eng is None:
        eng = CwEngine(net, n, h, w, epsilon=args.epsilon, lr_attack=args.lr_attack, clamp=args.clamp)
        _ENGINES.clear()
        _ENGINES[key] = eng
    else:
        eng.refresh_parameters()
    per_img = float(eng.per_img)
    min_noise = np.full(n, float(args.noise)); max_noise = np.full(n, 0.1)
    level = max_noise.copy()
    loss_i = np.zeros(n)
    done = np.zeros(n, dtype=bool)
    final_in = torch.zeros_like(im_s)
    for _ in range(args.search_steps):                                     # attack_cw.py:239
        loss_i_old = loss_i
        # ---- search_noise (attack_cw.py:142-192): fresh perturbation and optimiser, bisection on c
        eng.load(im_s, output_s)
        c_r = np.full(n, float(args.lamb_attack)); c_l = np.zeros(n); c = c_r.copy()
        eng.level.copy_(torch.as_tensor(level, dtype=torch.float32))
        for _ in range(args.search_steps):
            eng.c.copy_(torch.as_tensor(c, dtype=torch.float32))
            eng.run(args.steps)
            got = torch.stack((eng.st.loss_i, eng.loss_o_sum / per_img)).cpu().double().numpy()   # one D2H per block
            loss_i, mse_o = got[0], got[1]
            if record is not None:
                record.append((level.copy(), c.copy(), loss_i.copy(), mse_o.copy()))
            low = mse_o < 0.99 * level                                     # attack_cw.py:186-189
            c_l = np.where(low, c, c_l); c_r = np.where(low, c_r, c)
            c = (c_r + c_l) / 2
        im_in = eng.im_in_nchw()
        conv = (np.abs(loss_i - loss_i_old) < args.noise * 0.01) & (np.abs(loss_i - args.noise) < args.noise * 0.1)
        newly = ~done                                                      # images still searching take this result
        final_in[torch.as_tensor(newly, device=im_s.device)] = im_in[torch.as_tensor(newly, device=im_s.device)]
        done |= conv                                                       # attack_cw.py:248-249 (break)
        if done.all():
            break
        over = loss_i > args.noise                                         # attack_cw.py:251-255
        max_noise = np.where(over & ~done, level, max_noise)
        min_noise = np.where(~over & ~done, level, min_noise)
        level = np.where(done, level, (min_noise + max_noise) / 2)
    im_adv, output_adv, bpp, mse_results, vi_results = eval(final_in.contiguous(), im_s, output_s, net, args)
    return im_adv, output_adv, output_s, bpp_ori, bpp, mse_results["mse_in"], mse_results["mse_out"], vi_results["vi"]


@torch.no_grad()
def recompression(im_s, net, args, repeat_times=None):
    """Repeated coding of the same images (recompression.py:21-61; ``coder.code`` :154-164): each round decodes, clamps to
    [0, 1] and goes through the 8-bit PNG lattice (``write_image`` / ``read_image``, coder.py:20-48: round(x * 255) / 255)
    before it is coded again.  The whole batch stays on the device between rounds (the reference round-trips a PNG file per
    image and round).  Returns, for the LAST round as the reference reports it (:46-50): (x_hat, bpp, psnr, ms_ssim) with
    bpp / psnr / ms_ssim per batch (scalars) against the original images."""
    n_rounds = int(repeat_times if repeat_times is not None else args.steps)
    net.eval()
    x = im_s
    result = None
    for _ in range(n_rounds):
        result = net(x)
        x = torch.round(_clamp01(result["x_hat"]) * 255.0) / 255.0        # the PNG lattice between rounds
    x_hat = _clamp01(result["x_hat"])
    num_pixels = im_s.shape[0] * im_s.shape[2] * im_s.shape[3]            # RateDistortionLoss: N * H * W (train.py:60-64)
    bpp = float(_bpp(result["likelihoods"], num_pixels))
    mse = _mse(x_hat, im_s)
    psnr = -10.0 * math.log10(mse) if mse > 0 else float("inf")
    msim = float(metrics.ms_ssim(x_hat, im_s, data_range=1.0, size_average=True))
    return x_hat, bpp, psnr, msim
