"""Developer check: train-mode forward + RD-loss backward of every family; reports non-finite parameter gradients.
NaN-filled torch.empty (deterministic mode) exposes output regions no kernel writes."""
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200 import training as ptr  # noqa: E402

dev = torch.device("cuda:0")
for det in (False, True):
    torch.use_deterministic_algorithms(det, warn_only=True)
    if det:
        torch.utils.deterministic.fill_uninitialized_memory = True
    for model, q in (("hyper", 1), ("context", 1), ("cheng2020", 1)):
        torch.manual_seed(0)
        net = pm.init_model(model, q, "mse", pretrained=False).to(dev).train()
        x = torch.rand(2, 3, 128, 128, device=dev)
        out = net(x)
        crit = ptr.RateDistortionLoss("mse", ptr.LAMBDA_MSE[q])
        res = crit(out, x)
        net.zero_grad()
        res["loss"].backward()
        bad = [n for n, p in net.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
        none = [n for n, p in net.named_parameters() if p.grad is None and not n.endswith("quantiles")]
        print(f"det={det} {model}: loss {float(res['loss']):.5f} finite_out={bool(torch.isfinite(out['x_hat']).all())} "
              f"non-finite grads: {len(bad)} {bad[:6]}  no grad: {none[:6]}", flush=True)
        if bad:
            # which direction: distortion-only and rate-only
            for term in ("distortion_loss", "bpp_loss"):
                net.zero_grad()
                crit(net(x), x)[term].backward()
                b2 = [n for n, p in net.named_parameters() if p.grad is not None and not bool(torch.isfinite(p.grad).all())]
                print(f"    {term} alone: {len(b2)} non-finite {b2[:8]}", flush=True)
