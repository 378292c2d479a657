"""Kernel-level time of ONE codec update (train.py:351-365) at the config-5 shape: hyper q1, 8 x 256x256."""
import sys
import time
from types import SimpleNamespace

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200 import training as ptr  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 1, "mse", pretrained=False).to(dev)
args = SimpleNamespace(lr_train=1e-5)
crit = ptr.RateDistortionLoss("mse", ptr.LAMBDA_MSE[1])
opt, aux = ptr.configure_optimizers(net, args)
x = torch.rand(8, 3, 256, 256, device=dev)


def step():
    net.train()
    out = crit(net(x), x)
    opt.zero_grad(); aux.zero_grad()
    out["loss"].backward()
    opt.step()
    a = net.aux_loss(); a.backward(); aux.step(None)


for _ in range(3):
    step()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(5):
    step()
torch.cuda.synchronize()
print(f"update step: {1e3 * (time.perf_counter() - t0) / 5:.1f} ms wall")
from torch.profiler import ProfilerActivity, profile  # noqa: E402
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step()
    torch.cuda.synchronize()
rows = [(e.key, e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total, e.count) for e in prof.key_averages()]
rows = sorted((r for r in rows if r[1] > 0), key=lambda r: -r[1])
tot = sum(r[1] for r in rows)
print(f"device time total {tot / 1e3:.1f} ms")
for k, t, c in rows[:14]:
    print(f"  {t / 1e3:8.2f} ms  x{c:<4d} {k[:90]}")
print("host side (self CPU time):")
crow = sorted(((e.key, e.self_cpu_time_total, e.count) for e in prof.key_averages()), key=lambda r: -r[1])
for k, t, c in crow[:14]:
    print(f"  {t / 1e3:8.2f} ms  x{c:<4d} {k[:90]}")
import cProfile, pstats
pr = cProfile.Profile(); pr.enable(); step(); torch.cuda.synchronize(); pr.disable()
pstats.Stats(pr).sort_stats("cumulative").print_stats(28)
