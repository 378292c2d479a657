"""The codec update of adversarial training (``train.py --adv``, train.py:335-366) on the sm_100a kernels.

Mirrors, with the reference's names and argument meaning:
  * ``RateDistortionLoss``      train.py:37-96  (training branch; rate term with the 1/65536 likelihood clamp)
  * ``configure_optimizers``    coder.py:50-86  (main parameters vs ``*.quantiles``; Adam(lr_train) / Adam(1e-3))
  * ``adv_train_step``          train.py:335-366 (attack_ on the batch, train-mode forward, RD loss, backward with
                                weight gradients, clip_grad_norm_(1.0), Adam, aux loss on the quantiles)

Design for one process per GPU: all main parameters live in ONE flat fp32 buffer (parameters are views into it) and so do
their gradients, so the data-parallel exchange is a single NCCL all-reduce of 20 MB (hyperprior N=128/M=192) and
clip + Adam is one kernel over the flat buffer with the clip coefficient formed on the device (no host sync).
The unmodified reference ``train.py`` also works on these modules with its own torch optimisers.
"""
import math

import torch
import torch.distributed as dist

from . import functional as Fn
from . import metrics
from . import ops

LAMBDA_MSE = {1: 0.0018, 2: 0.0035, 3: 0.0067, 4: 0.0130, 5: 0.0250, 6: 0.0483, 7: 0.0932, 8: 0.1800}
LAMBDA_MSSSIM = {1: 2.40, 2: 4.58, 3: 8.73, 4: 16.64, 5: 31.73, 6: 60.50, 7: 115.37, 8: 220.00}


class _MsssimFn(torch.autograd.Function):
    """pytorch_msssim.MS_SSIM(data_range=1, size_average=True) with its gradient to the first argument."""

    @staticmethod
    def forward(ctx, x, target):
        xc, tc = x.detach().contiguous(), target.detach().contiguous()
        up = torch.ones(xc.shape[0], device=xc.device)
        val, grad = metrics.ms_ssim_value_and_grad(xc, tc, up)
        ctx.save_for_backward(grad)
        ctx.n = xc.shape[0]
        return val.mean().reshape(1)

    @staticmethod
    def backward(ctx, g):
        (grad,) = ctx.saved_tensors
        return ops.scaled_diff(grad, torch.zeros_like(grad), g.contiguous(), 1.0 / ctx.n), None


class RateDistortionLoss(torch.nn.Module):
    """train.py:37-96: the training branch and the ``training=False`` evaluation metrics (``lpips`` is out of scope)."""

    def __init__(self, metric="mse", lmbda=1e-2):
        super().__init__()
        self.metric, self.lmbda = metric, lmbda

    def forward(self, output, target, training=True):
        N, _, H, W = target.shape
        num_pixels = N * H * W
        out = {}
        bpp = None
        for lik in output["likelihoods"].values():                       # train.py:61-64
            term = Fn.LogSumFn.apply(lik, 1.0 / 65536) * (1.0 / (-math.log(2) * num_pixels))
            bpp = term if bpp is None else bpp + term
        out["bpp_loss"] = bpp.reshape(())
        if not training:                                                 # train.py:67-72 (evaluation metrics)
            from . import metrics
            with torch.no_grad():
                x_hat = torch.clamp(output["x_hat"], 0.0, 1.0)
                output["x_hat"] = x_hat
                out["mse_loss"] = Fn.MseFn.apply(x_hat, target).reshape(())
                out["msim_loss"] = metrics.ms_ssim(x_hat, target, data_range=1.0, size_average=True)
                out["psnr"] = -10.0 * math.log10(float(out["mse_loss"]))
                out["msim_dB"] = -10.0 * math.log10(1.0 - float(out["msim_loss"]))
            return out
        lamb_r = 0 if self.lmbda == 100 else 1                           # train.py:77-80
        if self.metric == "mse":
            out["distortion_loss"] = Fn.MseFn.apply(output["x_hat"], target).reshape(())
            out["loss"] = self.lmbda * 255 ** 2 * out["distortion_loss"] + lamb_r * out["bpp_loss"]
        elif self.metric == "ms-ssim":
            out["distortion_loss"] = _MsssimFn.apply(output["x_hat"], target).reshape(())
            out["loss"] = self.lmbda * (1 - out["distortion_loss"]) + lamb_r * out["bpp_loss"]
        else:
            raise ValueError(f"metric {self.metric} is not built")
        return out


def exchange_gradients(flat_grad, group=None):
    """The one exchange step of data-parallel adversarial training: SUM all-reduce of the flat gradient buffer (NCCL over
    NVLink on GPUs; gloo in the CPU tests).  Returns the factor the optimiser must scale gradients by (1/world)."""
    if not (dist.is_available() and dist.is_initialized()):
        return 1.0
    world = dist.get_world_size(group)
    if world == 1:
        return 1.0
    dist.all_reduce(flat_grad, op=dist.ReduceOp.SUM, group=group)
    return 1.0 / world


class FusedAdamClip(torch.optim.Optimizer):
    """Adam over ONE flat parameter buffer with ``clip_grad_norm_(max_norm)`` folded in (train.py:360-361).

    ``params`` keep their identity (``state_dict`` keys, shapes) but their storage becomes a view of ``self.flat`` and
    their ``.grad`` a view of ``self.flat_grad``; ``zero_grad`` zeroes the flat buffer in place.

    A ``torch.optim.Optimizer``: ``param_groups[0]["lr"]`` is the live learning rate (schedulers such as
    ``ReduceLROnPlateau`` attach to it, coder.py:109-116 reads it back on resume), and ``state_dict()`` /
    ``load_state_dict()`` use ``torch.optim.Adam``'s layout (per-parameter ``step`` / ``exp_avg`` / ``exp_avg_sq``), so
    the reference's checkpoint format (train.py:443-454: ``optimizer`` / ``aux_optimizer`` entries) round-trips and a
    resumed run continues with the same moments."""

    def __init__(self, params, lr, max_norm=None, betas=(0.9, 0.999), eps=1e-8):
        params = [p for p in params]
        assert params, "no parameters"
        super().__init__(params, dict(lr=lr, betas=betas, eps=eps, weight_decay=0, amsgrad=False))
        self.params = params
        dev = self.params[0].device
        total = sum(p.numel() for p in self.params)
        self.flat = torch.empty(total, device=dev, dtype=torch.float32)
        self.flat_grad = torch.zeros(total, device=dev, dtype=torch.float32)
        self.m, self.v = torch.zeros_like(self.flat), torch.zeros_like(self.flat)
        off = 0
        self._spans = []
        for p in self.params:
            n = p.numel()
            self.flat[off:off + n].copy_(p.detach().reshape(-1))
            p.data = self.flat[off:off + n].view(p.shape)
            p.grad = self.flat_grad[off:off + n].view(p.shape)
            self._spans.append((off, n))
            off += n
        self.max_norm = max_norm
        self.steps = 0
        self.last_sumsq = None

    @property
    def lr(self):
        return self.param_groups[0]["lr"]

    @lr.setter
    def lr(self, value):
        self.param_groups[0]["lr"] = value

    def zero_grad(self, set_to_none=False):
        self.flat_grad.zero_()
        for p, (off, n) in zip(self.params, self._spans):      # autograd may have replaced a view: re-attach
            if p.grad is None or p.grad.data_ptr() != self.flat_grad[off:off + n].data_ptr():
                p.grad = self.flat_grad[off:off + n].view(p.shape)

    def step(self, group=None, exchange=True):
        """``exchange=False`` skips the data-parallel all-reduce (the auxiliary optimiser: ``aux_loss`` depends on the
        parameters only, so its gradient is already identical on every rank)."""
        for p, (off, n) in zip(self.params, self._spans):      # a gradient that autograd re-allocated is copied back
            if p.grad is not None and p.grad.data_ptr() != self.flat_grad[off:off + n].data_ptr():
                self.flat_grad[off:off + n].copy_(p.grad.reshape(-1))
        scale = exchange_gradients(self.flat_grad, group) if exchange else 1.0
        self.steps += 1
        sumsq = ops.sumsq(self.flat_grad) if self.max_norm is not None else None
        self.last_sumsq = sumsq
        g = self.param_groups[0]
        ops.adam_clip_step(self.flat, self.flat_grad, self.m, self.v, sumsq_dev=sumsq,
                           max_norm=self.max_norm if self.max_norm is not None else 0.0, lr=g["lr"], step=self.steps,
                           grad_scale=scale, beta1=g["betas"][0], beta2=g["betas"][1], eps=g["eps"])
        Fn.invalidate_pack_cache()   # the kernel updated weights behind autograd's version counters

    def state_dict(self):
        """torch.optim.Adam layout: {"state": {i: {"step", "exp_avg", "exp_avg_sq"}}, "param_groups": [...]}."""
        state = {}
        if self.steps > 0:
            for i, (p, (off, n)) in enumerate(zip(self.params, self._spans)):
                state[i] = {"step": torch.tensor(float(self.steps)),
                            "exp_avg": self.m[off:off + n].view(p.shape).clone(),
                            "exp_avg_sq": self.v[off:off + n].view(p.shape).clone()}
        group = {k: v for k, v in self.param_groups[0].items() if k != "params"}
        group["params"] = list(range(len(self.params)))
        return {"state": state, "param_groups": [group]}

    def load_state_dict(self, state_dict):
        groups = state_dict["param_groups"]
        if len(groups) != 1 or len(groups[0]["params"]) != len(self.params):
            raise ValueError("loaded state dict does not match this optimiser's parameters")
        for k, v in groups[0].items():
            if k != "params":
                self.param_groups[0][k] = v
        self.m.zero_(); self.v.zero_()
        steps = 0
        for i, st in state_dict["state"].items():
            p, (off, n) = self.params[int(i)], self._spans[int(i)]
            if tuple(st["exp_avg"].shape) != tuple(p.shape):
                raise ValueError(f"optimizer state {i}: shape {tuple(st['exp_avg'].shape)} != {tuple(p.shape)}")
            self.m[off:off + n].copy_(st["exp_avg"].reshape(-1))
            self.v[off:off + n].copy_(st["exp_avg_sq"].reshape(-1))
            steps = max(steps, int(float(st["step"])))
        self.steps = steps


def configure_optimizers(net, args):
    """coder.py:50-86 on the fused optimiser: (main, aux)."""
    named = dict(net.named_parameters())
    main = sorted(n for n, p in named.items() if not n.endswith(".quantiles") and p.requires_grad)
    aux = sorted(n for n, p in named.items() if n.endswith(".quantiles") and p.requires_grad)
    assert not (set(main) & set(aux))
    return (FusedAdamClip([named[n] for n in main], lr=args.lr_train, max_norm=1.0),
            FusedAdamClip([named[n] for n in aux], lr=1e-3, max_norm=None))


def adv_train_step(batch_x, net, args, criterion, optimizer, aux_optimizer, group=None, timings=None,
                   budget_scope="batch"):
    """One iteration of train.py:335-366 with N_ADV = 0.  Returns (out_criterion, aux_loss).
    ``budget_scope``: "batch" (default) = the reference's semantics on this path -- train.py:342 passes the whole batch
    to ``attack_``, whose ``attack_our`` tests the budget on the batch mean (attack_rd.py:333-334) and takes one branch
    for all images; "image" = every image tests its own budget (what the sharded CLI attack does).
    ``timings`` (a dict of lists, benchmark use): receives the CUDA-event times in ms of the attack part
    (``attack_ms``), the codec update (``update_ms``) and, inside it, the gradient all-reduce (``allreduce_ms``);
    measuring them synchronises the stream three times per step."""
    from .attack import attack_
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(5)] if timings is not None else None
    if ev:
        ev[0].record()
    if getattr(args, "budget_scope", None) != budget_scope:
        import copy
        args = copy.copy(args)
        args.budget_scope = budget_scope
    batch_adv = attack_(batch_x, net, args, metrics_pass=False)[0].detach()   # :342-343 (only [0] is used: no eval pass,
                                                                           # no host sync between the attack and the update)
    if ev:
        ev[1].record()
    net.train()                                                          # :346
    batch = batch_adv.clone()                                            # :347
    result = net(batch)                                                  # :351
    out = criterion(result, batch)                                       # :353
    optimizer.zero_grad()
    aux_optimizer.zero_grad()
    out["loss"].backward()                                               # :359
    if ev:
        ev[2].record()
    optimizer.step(group)                                                # all-reduce + clip_grad_norm_(1.0) + Adam (:360-361)
    if ev:
        ev[3].record()
    aux = net.aux_loss()                                                 # :363 (parameter-only math: identical on all ranks)
    aux.backward()
    aux_optimizer.step(exchange=False)
    if ev:
        ev[4].record()
        torch.cuda.synchronize()
        timings.setdefault("attack_ms", []).append(ev[0].elapsed_time(ev[1]))
        timings.setdefault("update_ms", []).append(ev[1].elapsed_time(ev[4]))
        timings.setdefault("allreduce_ms", []).append(ev[2].elapsed_time(ev[3]))   # all-reduce + sumsq + clip/Adam kernel
    return out, aux


def test_epoch(epoch, test_dataloader, model, criterion, log_dir, args, group=None):
    """train.py:196-242.  ``--adv``: every batch is attacked with the input budget forced to 1e-4 (:212-216) and the
    mean VI (10 log10(mse_out / mse_in)) is returned; otherwise the RD loss, its terms and the auxiliary loss are
    averaged over the batches (``training=True`` in the reference's criterion call only selects the loss form, :220).
    One log line is appended to ``log_dir`` (a file path in the reference, :233-235).  Multi-GPU (one process per GPU,
    each with its shard of the loader): the batch means and counts are summed over ``group`` before the averages form."""
    from .attack import attack_
    model.eval()
    device = next(model.parameters()).device
    sums = {"loss": 0.0, "bpp": 0.0, "aux": 0.0, "d": 0.0, "vi": 0.0}
    count = 0
    for d in test_dataloader:
        d = d.to(device)
        if args.adv:
            noise = args.noise
            args.noise = 0.0001                                           # :214 force the input perturbation level
            try:
                vi_results = attack_(d, model, args)[-1]
            finally:
                args.noise = noise
            sums["vi"] += float(vi_results["vi"])
        else:
            with torch.no_grad():
                out_net = model(d.detach())
                out = criterion(out_net, d)
                sums["aux"] += float(model.aux_loss())
                sums["bpp"] += float(out["bpp_loss"])
                sums["loss"] += float(out["loss"])
                sums["d"] += float(out["distortion_loss"])
        count += 1
    if group is not None or (torch.distributed.is_available() and torch.distributed.is_initialized()):
        t = torch.tensor([sums[k] for k in ("loss", "bpp", "aux", "d", "vi")] + [float(count)], device=device,
                         dtype=torch.float64)
        torch.distributed.all_reduce(t, group=group)
        vals = t.tolist()
        sums = dict(zip(("loss", "bpp", "aux", "d", "vi"), vals[:5]))
        count = int(vals[5])
    n = max(count, 1)
    log = (f"Test epoch {epoch}: Average losses:\tLoss: {sums['loss'] / n:.4f} |\tMSE loss: {sums['d'] / n:.6f} |"
           f"\tBpp loss: {sums['bpp'] / n:.3f} |\tRecomp loss: {0.0:.3f}\n")
    if log_dir:
        with open(log_dir, "a") as f:
            f.write(log)
    if args.adv:
        print(f"Test Loss (VI): {sums['vi'] / n:.4f}")
        return sums["vi"] / n
    print(log)
    return sums["loss"] / n


@torch.no_grad()
def batch_test(images, net, criterion=None):
    """test.py:28-60 without the file IO: eval-mode ``net(x)`` per image (``coder.code``, coder.py:154-164), the
    evaluation branch of the criterion, and the four averages of the ``AVG:`` line: (bpp, psnr, ms_ssim, ms_ssim_dB).
    ``images``: an iterable of [1, 3, H, W] tensors (or one [N, 3, H, W] tensor, taken image by image like the reference)."""
    criterion = criterion or RateDistortionLoss()
    net.eval()
    sums = [0.0, 0.0, 0.0, 0.0]
    count = 0
    for im in images:
        im = im if im.dim() == 4 else im.unsqueeze(0)
        result = net(im)
        m = criterion(result, im, training=False)
        for i, k in enumerate(("bpp_loss", "psnr", "msim_loss", "msim_dB")):
            sums[i] += float(m[k])
        count += 1
    avg = [v / max(count, 1) for v in sums]
    print("AVG:", *avg)
    return tuple(avg)
