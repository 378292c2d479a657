"""Developer check: with torch.use_deterministic_algorithms(True) torch.empty() returns NaN-filled memory (the reference
runs in that mode, self_ensemble.py:31), so any kernel that leaves part of its output unwritten shows up as NaN.
Walks the codec module by module (forward outputs and input gradients) for every model family."""
import sys
import torch
sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm

torch.use_deterministic_algorithms(True, warn_only=True)
dev = torch.device("cuda:0")
bad = 0
for model, q, hw in (("cheng2020", 1, (192, 192)), ("cheng2020", 6, (64, 128)), ("context", 4, (128, 192)), ("hyper", 3, (192, 256)),
                     ("factorized", 1, (64, 96))):
    torch.manual_seed(0)
    net = pm.init_model(model, q, "mse", pretrained=False).to(dev).train()
    for p in net.parameters():
        p.requires_grad_(False)
    x = torch.rand(2, 3, *hw, device=dev).requires_grad_(True)
    names = {m: n for n, m in net.named_modules()}

    def fwd_hook(mod, inp, out):
        global bad
        outs = out if isinstance(out, (tuple, list)) else (out,)
        for o in outs:
            if torch.is_tensor(o) and o.is_floating_point() and not torch.isfinite(o).all():
                bad += 1
                print(f"[{model} q{q}] forward NaN after {names[mod]} ({type(mod).__name__}) shape {tuple(o.shape)} "
                      f"frac {float((~torch.isfinite(o)).float().mean()):.4f}")
            if torch.is_tensor(o) and o.requires_grad:
                o.register_hook(lambda g, mod=mod: bwd_check(g, mod))

    def bwd_check(g, mod):
        global bad
        if not torch.isfinite(g).all():
            bad += 1
            print(f"[{model} q{q}] backward NaN in grad wrt OUTPUT of {names[mod]} ({type(mod).__name__}) shape {tuple(g.shape)} "
                  f"frac {float((~torch.isfinite(g)).float().mean()):.4f}")

    hooks = [m.register_forward_hook(fwd_hook) for m in net.modules() if len(list(m.children())) == 0]
    out = net.g_s(net.g_a(x))
    out.backward(torch.rand_like(out))
    if not torch.isfinite(x.grad).all():
        bad += 1
        print(f"[{model} q{q}] input gradient has NaN: frac {float((~torch.isfinite(x.grad)).float().mean()):.4f}")
    full = net(x.detach())
    for k, v in [("x_hat", full["x_hat"])] + list(full["likelihoods"].items()):
        if not torch.isfinite(v).all():
            bad += 1
            print(f"[{model} q{q}] net(x)[{k}] has NaN")
    for h in hooks:
        h.remove()
print("non-finite findings:", bad)
