"""Oracle restatement of the reference's perturbation loop.  Test infrastructure only.

Follows, line by line where it matters:
  * ``attack_our``  -- /root/reference/attack_rd.py:332-379
  * ``attack_``     -- /root/reference/attack_rd.py:381-575
  * ``eval``        -- /root/reference/self_ensemble.py:173-252
  * sign update     -- /root/reference/attack_ifgsm.py:348-438
  * RD loss / optimisers / --adv step -- /root/reference/train.py:37-96, 335-366; coder.py:50-86

Differences from the reference text, all forced by it being un-runnable off-GPU: the perturbation
is created on ``im_s.device`` instead of ``.cuda()`` (attack_rd.py:501), and ms_ssim comes from
``oracle.msssim`` (variant 1) instead of the missing ``pytorch_msssim``.
"""
import math
from types import SimpleNamespace

import torch

from .layers import low_bound, up_bound
from .msssim import MS_SSIM, ms_ssim


def default_args(**kw):
    """Defaults of ``coder.config()`` (coder.py:166-220) for the flags the hot path reads."""
    a = dict(model="hyper", metric="ms-ssim", quality=3, steps=1001, random=1, lamb_attack=0.2,
             noise=1e-4, lr_attack=0.01, att_metric="L2", epsilon=16.0, pad=None, debug=False,
             clamp=True, defend=False, method="ensemble", adv=False, lr_train=1e-4, search_steps=20,
             target=None, mask_loc=None, lamb_bkg_in=1.0, lamb_bkg_out=1.0, lamb_tar=1.0)
    a.update(kw)
    return SimpleNamespace(**a)


def bpp_from_likelihoods(likelihoods, num_pixels):
    """attack_rd.py:419 / self_ensemble.py:222."""
    return sum(torch.log(l).sum() / (-math.log(2) * num_pixels) for l in likelihoods.values())


def attack_our(im_s, output_s, im_in, net, args):
    """attack_rd.py:332-379."""
    loss_i = torch.mean((im_s - im_in) ** 2)
    force = getattr(args, "force_branch", -1)                          # test hook: 1 = always the network branch
    if (loss_i > args.noise) if force < 0 else (force == 0):           # :334 (host sync)
        if args.att_metric == "ms-ssim":
            loss = 1.0 - ms_ssim(im_s, im_in, data_range=1.0, size_average=True)
        if args.att_metric == "L2":
            loss = loss_i
        loss_o = torch.zeros(1)
        branch = "A"
    else:
        y_main = net.g_a(im_in)                                        # :344 no quantiser
        x_ = net.g_s(y_main)                                           # :349
        output_ = up_bound(low_bound(x_, 0.0), 1.0) if args.clamp else x_   # :353-356
        if args.att_metric == "ms-ssim":
            loss_o = ms_ssim(output_, output_s, data_range=1.0, size_average=True)
        if args.att_metric == "L2":
            loss_o = 1.0 - torch.mean((output_s - output_) * (output_s - output_))  # :364
        loss = loss_o
        branch = "B"
    return loss, loss_i, loss_o, branch


def roi_masks(im_s, mask_loc):
    """attack_cv.py:149-163: ``mask_bkg`` is 1 outside ``[y0:y1, x0:x1]`` (``mask_loc = x0 x1 y0 y1``, W then H) and 0
    inside; without ``--mask_loc`` the whole image is the target area.  Returns (mask_bkg, mask_tar), [1,C,H,W]."""
    _, C, H, W = im_s.shape
    mask = torch.zeros(1, C, H, W, device=im_s.device)
    if mask_loc is not None:
        mask = torch.ones(1, C, H, W, device=im_s.device)
        x0, x1, y0, y1 = mask_loc
        mask[:, :, y0:y1, x0:x1] = 0.0
    return mask, 1.0 - mask


def attack_our_roi(im_s, output_s, output_t, im_in, net, args, mask_bkg, mask_tar):
    """Targeted / ROI loss.  The reference's own call path for these flags is dead code
    (attack_rd.py never reads ``-t`` / ``--mask_loc`` in the loss); the formulas are those of attack_cv.py:149-163 and
    attack_data.py:204-221, with the per-pixel masks INSIDE the means (the reference text multiplies a scalar mean by
    a mask tensor, which would not even be a scalar loss): THIS restatement is the definition parity is measured on.
      loss_i = mean(d_in^2 mask_tar) + lamb_bkg_in mean(d_in^2 mask_bkg);   switch on loss_i >= noise (attack_data.py:219)
      loss_o = lamb_tar mean((out - output_t)^2 mask_tar) + lamb_bkg_out mean((out - output_s)^2 mask_bkg)  (minimised)
    """
    d2 = (im_s - im_in) ** 2
    loss_i = torch.mean(d2 * mask_tar) + args.lamb_bkg_in * torch.mean(d2 * mask_bkg)
    if loss_i >= args.noise:
        return loss_i, loss_i, torch.zeros(1), "A"
    x_ = net.g_s(net.g_a(im_in))
    output_ = up_bound(low_bound(x_, 0.0), 1.0) if args.clamp else x_
    loss_o = (args.lamb_tar * torch.mean((output_ - output_t) ** 2 * mask_tar) +
              args.lamb_bkg_out * torch.mean((output_ - output_s) ** 2 * mask_bkg))
    return loss_o, loss_i, loss_o, "B"


@torch.no_grad()
def eval_metrics(im_adv, im_s, output_s, net, args):
    """self_ensemble.py:173-252 (defence branches are out of scope)."""
    net.eval()
    im_ = torch.clamp(im_adv, min=0.0, max=1.0) if args.clamp else im_adv
    result = net(im_)
    x_hat = result["x_hat"]
    mse_in = torch.mean((im_ - im_s) ** 2)
    output_ = torch.clamp(x_hat, min=0.0, max=1.0) if args.clamp else x_hat
    num_pixels = im_adv.shape[2] * im_adv.shape[3]
    bpp = bpp_from_likelihoods(result["likelihoods"], num_pixels)
    mse_out = torch.mean((output_ - output_s) ** 2)
    small = min(im_.shape[-2:]) <= 160  # pytorch_msssim asserts; oracle reports None instead
    msim_out = None if small else ms_ssim(output_, output_s, data_range=1.0).item()
    msim_in = None if small else ms_ssim(im_, im_s, data_range=1.0).item()
    mse_results = {"mse_in": mse_in.item(), "mse_out": mse_out.item()}
    vi_results = {"vi": None, "vi_msim": None}
    if mse_in > 1e-20 and mse_out > 1e-20:
        vi_results["vi"] = 10.0 * math.log10(mse_out.item() / mse_in.item())
        if not args.adv and msim_in is not None and msim_in < 0.9999:
            vi_results["vi_msim"] = 10.0 * math.log10((1 - msim_out) / (1 - msim_in))
    return im_, output_, bpp, mse_results, vi_results


def clean_pass(im_s, net, args):
    """attack_rd.py:402-419: eval-mode full forward -> output_s, bpp_ori."""
    H, W = im_s.shape[2], im_s.shape[3]
    with torch.no_grad():
        net.eval()
        result = net(im_s)
        output_s = torch.clamp(result["x_hat"], 0.0, 1.0) if args.clamp else result["x_hat"]
        bpp_ori = bpp_from_likelihoods(result["likelihoods"], H * W)
    return output_s, bpp_ori, result


def attack_(im_s, net, args, record=None, noise_init=None, im_t=None):
    """attack_rd.py:381-575.  ``record`` (a list) receives per-step
    ``(branch, loss, loss_i)``; ``noise_init`` lets a test share the ``-random>1`` start; ``im_t`` (the ``-t`` image)
    switches to the targeted / ROI loss ``attack_our_roi``."""
    output_s, bpp_ori, _ = clean_pass(im_s, net, args)
    roi = im_t is not None
    if roi:
        output_t = clean_pass(im_t, net, args)[0]
        mask_bkg, mask_tar = roi_masks(im_s, args.mask_loc)
    noise_range = args.epsilon / 255.0                                  # :490
    if noise_init is not None:
        noise = noise_init.clone()
    elif args.random > 1:
        noise = torch.empty_like(im_s).uniform_(-1e-2, 1e-2)            # :498-499
    else:
        noise = torch.zeros_like(im_s)                                  # :496
    noise = noise.to(im_s.device).requires_grad_(True)
    optimizer = torch.optim.Adam([noise], lr=args.lr_attack)            # :502
    sched = torch.optim.lr_scheduler.MultiStepLR(optimizer, [1, 2, 3], gamma=0.33)  # :503
    net.train()                                                         # :504
    for i in range(args.steps):
        noise_clipped = up_bound(low_bound(noise, -noise_range), noise_range)      # :507
        im_in = up_bound(low_bound(im_s + noise_clipped, 0.0), 1.0)                # :517
        if roi:
            loss, loss_i, loss_o, branch = attack_our_roi(im_s, output_s, output_t, im_in, net, args, mask_bkg, mask_tar)
        else:
            loss, loss_i, loss_o, branch = attack_our(im_s, output_s, im_in, net, args)
        optimizer.zero_grad()
        loss.backward()
        optimizer.step()                                                # :546-548
        if record is not None:
            record.append((branch, float(loss.detach()), float(loss_i.detach())))
        if i % (args.steps // 3) == 0:                                  # :553 (ZeroDivisionError if steps<3)
            sched.step()
    im_adv, output_adv, bpp, mse_results, vi_results = eval_metrics(
        im_in.detach(), im_s, output_s, net, args)                      # :573
    return im_adv, output_adv, output_s, bpp_ori, bpp, mse_results, vi_results


def attack_cw(im_s, net, args, record=None):
    """attack_cw.py:111-263 for ONE image (the reference's unit): outer bisection on the reconstruction-error level
    (:233-257), inner bisection on the loss weight c (:142-192), loss = loss_i + c (1 - MSE_o) with c zeroed once
    MSE_o > 1.1 level (:111-140).  ``args.search_steps`` bounds both searches (coder.py flag).  ``record`` receives one
    ``(level, c, loss_i, mse_o)`` per block of ``args.steps`` iterations.  Returns the reference's tuple
    (im_adv, output_adv, output_s, bpp_ori, bpp, mse_in, mse_out, vi)."""
    output_s, bpp_ori, _ = clean_pass(im_s, net, args)                      # :216-230
    net.train()                                                             # :231
    noise_range = args.epsilon / 255.0

    def search_noise(noise_level):                                          # :142-192
        noise = torch.zeros_like(im_s).requires_grad_(True)                 # :145-148
        optimizer = torch.optim.Adam([noise], lr=args.lr_attack)            # :149
        c_r, c_l = args.lamb_attack, 0.0
        c = c_r
        for _ in range(args.search_steps):                                  # :158
            for _ in range(args.steps):                                     # :159
                noise_clipped = up_bound(low_bound(noise, -noise_range), noise_range)
                im_in = up_bound(low_bound(im_s + noise_clipped, 0.0), 1.0)
                loss_i = torch.mean((im_s - im_in) ** 2)                    # :114
                out = up_bound(low_bound(net.g_s(net.g_a(im_in)), 0.0), 1.0)
                loss_o = 1.0 - torch.mean((output_s - out) * (output_s - out))   # :137
                cc = 0.0 if float(1.0 - loss_o) > noise_level * 1.1 else c        # :138-139
                loss = loss_i + cc * loss_o
                optimizer.zero_grad()
                loss.backward()
                optimizer.step()
            if record is not None:
                record.append((noise_level, c, float(loss_i.detach()), float(1.0 - loss_o.detach())))
            if float(1.0 - loss_o) < 0.99 * noise_level:                    # :186
                c_l = c
            else:
                c_r = c
            c = (c_r + c_l) / 2                                             # :190
        return float(loss_i.detach()), im_in.detach()

    min_noise, max_noise = args.noise, 0.1                                  # :232-234
    noise_level = max_noise
    loss_i = 0.0
    search_cnt = 0
    im_in = im_s
    while search_cnt < args.search_steps:                                   # :239
        loss_i_old = loss_i
        loss_i, im_in = search_noise(noise_level)
        if abs(loss_i - loss_i_old) < args.noise * 0.01 and abs(loss_i - args.noise) < args.noise * 0.1:   # :248
            break
        if loss_i > args.noise:
            max_noise = noise_level
        else:
            min_noise = noise_level
        noise_level = (min_noise + max_noise) / 2
        search_cnt += 1
    im_adv, output_adv, bpp, mse_results, vi_results = eval_metrics(im_in, im_s, output_s, net, args)
    return im_adv, output_adv, output_s, bpp_ori, bpp, mse_results["mse_in"], mse_results["mse_out"], vi_results["vi"]


def lr_schedule(steps, lr0=0.01, gamma=0.33):
    """LR used at iteration i under MultiStepLR([1,2,3]) stepped when i % (steps//3) == 0."""
    out, lr, n = [], lr0, 0
    for i in range(steps):
        out.append(lr)
        if i % (steps // 3) == 0:
            n += 1
            if n in (1, 2, 3):
                lr = lr * gamma
    return out


def attack_ifgsm(im_s, net, args, momentum=False, record=None, start=None):
    """attack_ifgsm.py:364-438 (single start).  ``start`` injects the PGD random start."""
    output_s, bpp_ori, _ = clean_pass(im_s, net, args)
    eps = args.epsilon / 255.0
    im_adv = (im_s if start is None else torch.clamp(start, 0, 1)).detach().requires_grad_(True)
    net.train()
    g, alpha = 0, eps / args.steps
    for i in range(args.steps):
        output_ = net.g_s(net.g_a(im_adv))
        loss_o = torch.mean((output_s - output_) * (output_s - output_))
        net.zero_grad()
        grad, = torch.autograd.grad(loss_o, im_adv)
        if momentum:                                                    # :348-362
            g = 1.0 * g + grad / torch.norm(grad, p=1)
            im_new = torch.clamp(im_adv + alpha * torch.sign(g), 0, 1)
        else:
            im_new = im_adv + eps / args.steps * grad.sign()            # :409-415
        im_new = torch.where(im_new > im_s + eps, im_s + eps, im_new)   # :417-418
        im_new = torch.where(im_new < im_s - eps, im_s - eps, im_new)
        im_adv = im_new.detach().requires_grad_(True)
        if record is not None:
            record.append(("B", float(loss_o), float(torch.mean((im_adv - im_s) ** 2))))
    im_, output_adv, bpp, mse_results, vi_results = eval_metrics(im_adv.detach(), im_s, output_s, net, args)
    return im_, output_adv, output_s, bpp_ori, bpp, mse_results, vi_results


# ---------------------------------------------------------------- train.py --adv pieces
LAMBDA_MSE = {1: 0.0018, 2: 0.0035, 3: 0.0067, 4: 0.0130, 5: 0.0250, 6: 0.0483, 7: 0.0932, 8: 0.1800}
LAMBDA_MSSSIM = {1: 2.40, 2: 4.58, 3: 8.73, 4: 16.64, 5: 31.73, 6: 60.50, 7: 115.37, 8: 220.00}


class RateDistortionLoss(torch.nn.Module):
    """train.py:37-96 (training branch; lpips is out of scope)."""

    def __init__(self, metric="mse", lmbda=1e-2):
        super().__init__()
        self.metric, self.lmbda = metric, lmbda
        self.mssim = MS_SSIM(data_range=1, size_average=True, channel=3)

    def forward(self, output, target):
        N, _, H, W = target.shape
        num_pixels = N * H * W
        out = {}
        bpp = 0.0
        for lik in output["likelihoods"].values():
            lik = torch.clamp(lik, min=1.0 / 65536)                     # :62
            bpp = bpp + torch.log(lik).sum() / (-math.log(2) * num_pixels)
        out["bpp_loss"] = bpp
        if self.metric == "mse":
            out["distortion_loss"] = torch.mean((output["x_hat"] - target) ** 2)
            out["loss"] = self.lmbda * 255 ** 2 * out["distortion_loss"] + out["bpp_loss"]
        else:
            out["distortion_loss"] = self.mssim(output["x_hat"], target)
            out["loss"] = self.lmbda * (1 - out["distortion_loss"]) + out["bpp_loss"]
        return out


def configure_optimizers(net, lr_train):
    """coder.py:50-86."""
    params = dict(net.named_parameters())
    main = sorted(n for n in params if not n.endswith(".quantiles"))
    aux = sorted(n for n in params if n.endswith(".quantiles"))
    return (torch.optim.Adam((params[n] for n in main), lr=lr_train),
            torch.optim.Adam((params[n] for n in aux), lr=1e-3))


def adv_train_step(batch_x, net, args, criterion, optimizer, aux_optimizer):
    """One iteration of train.py:335-366 (N_ADV = 0)."""
    batch_adv = attack_(batch_x, net, args)[0].detach()                 # :342
    net.train()
    batch_x = batch_adv.clone()                                         # :347 (in-place overwrite)
    result = net(batch_x)
    out = criterion(result, batch_x)                                    # :353
    optimizer.zero_grad()
    aux_optimizer.zero_grad()
    out["loss"].backward()
    torch.nn.utils.clip_grad_norm_(net.parameters(), 1.0)               # :360
    optimizer.step()
    aux = net.aux_loss()
    aux.backward()
    aux_optimizer.step()
    return out, aux


def test_epoch(test_dataloader, model, criterion, args):
    """train.py:196-242: mean VI over attacked batches under ``--adv`` (budget forced to 1e-4, :212-216), else the mean
    RD loss.  Returns (value, dict of the other averages)."""
    model.eval()
    device = next(model.parameters()).device
    vi, loss, bpp, dist, aux, n = 0.0, 0.0, 0.0, 0.0, 0.0, 0
    for d in test_dataloader:
        d = d.to(device)
        if args.adv:
            noise, args.noise = args.noise, 0.0001
            vi += attack_(d, model, args)[-1]["vi"]
            args.noise = noise
        else:
            with torch.no_grad():
                out = criterion(model(d.detach()), d)
                aux += float(model.aux_loss())
                loss += float(out["loss"]); bpp += float(out["bpp_loss"]); dist += float(out["distortion_loss"])
        n += 1
    avg = {"loss": loss / n, "bpp_loss": bpp / n, "distortion_loss": dist / n, "aux_loss": aux / n, "vi": vi / n}
    return (avg["vi"] if args.adv else avg["loss"]), avg


@torch.no_grad()
def recompression(im_s, net, args, repeat_times):
    """recompression.py:21-61 without the PNG files: decode, clamp, 8-bit lattice (coder.py:20-48), code again; the last
    round's bpp (RateDistortionLoss convention, train.py:60-64), PSNR and MS-SSIM against the original."""
    net.eval()
    x = im_s
    for _ in range(repeat_times):
        result = net(x)
        x = torch.round(torch.clamp(result["x_hat"], 0.0, 1.0) * 255.0) / 255.0
    x_hat = torch.clamp(result["x_hat"], 0.0, 1.0)
    n, _, h, w = im_s.shape
    bpp = float(sum(torch.log(l).sum() / (-math.log(2) * n * h * w) for l in result["likelihoods"].values()))
    psnr = -10.0 * math.log10(float(torch.mean((x_hat - im_s) ** 2)))
    msim = float(ms_ssim(x_hat, im_s, data_range=1.0, size_average=True))
    return x_hat, bpp, psnr, msim


@torch.no_grad()
def batch_test(images, net):
    """test.py:28-60 + the evaluation branch of RateDistortionLoss (train.py:50-72): per image bpp (likelihoods clamped at
    1/65536, N*H*W pixels), PSNR and MS-SSIM of the clamped reconstruction; returns the four averages of the AVG: line."""
    net.eval()
    sums = [0.0, 0.0, 0.0, 0.0]
    for im in images:
        im = im if im.dim() == 4 else im.unsqueeze(0)
        result = net(im)
        n, _, h, w = im.shape
        bpp = float(sum(torch.log(torch.clamp(l, min=1.0 / 65536)).sum() / (-math.log(2) * n * h * w)
                        for l in result["likelihoods"].values()))
        x_hat = torch.clamp(result["x_hat"], 0.0, 1.0)
        mse = float(torch.mean((x_hat - im) ** 2))
        msim = float(ms_ssim(x_hat, im, data_range=1.0, size_average=True))
        for i, v in enumerate((bpp, -10.0 * math.log10(mse), msim, -10.0 * math.log10(1.0 - msim))):
            sums[i] += v
    return tuple(v / len(images) for v in sums)


# ---------------------------------------------------------------- synthetic inputs
def synthetic_image(i, H=512, W=768, device="cpu"):
    """Seeded Kodak-like image on the k/255 lattice: U[0,1) field -> separable Gaussian blur
    sigma=3 -> affine rescale to [0.05, 0.95] -> round(.*255)/255.  Returns [1,3,H,W] fp32."""
    g = torch.Generator().manual_seed(1234 + i)
    x = torch.rand(1, 3, H, W, generator=g)
    r = 9
    k = torch.exp(-(torch.arange(-r, r + 1, dtype=torch.float32) ** 2) / (2 * 3.0 ** 2))
    k = k / k.sum()
    xp = torch.nn.functional.pad(x, (r, r, r, r), mode="reflect")
    xp = torch.nn.functional.conv2d(xp, k.view(1, 1, 1, -1).expand(3, 1, 1, -1), groups=3)
    xp = torch.nn.functional.conv2d(xp, k.view(1, 1, -1, 1).expand(3, 1, -1, 1), groups=3)
    lo, hi = xp.amin(), xp.amax()
    xp = 0.05 + 0.9 * (xp - lo) / (hi - lo)
    return (torch.round(xp * 255.0) / 255.0).to(device)
