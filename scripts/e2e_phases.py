"""Where the end-to-end attack_() call spends its time (64 x 512x768, hyper q3, 60 forced network iterations)."""
import argparse
import sys
import time

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import attack as patk  # noqa: E402
from imagecompression_adversarial_b200 import models as pm  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev)
host = (torch.randint(0, 256, (n, 3, 512, 768), dtype=torch.uint8).float() / 255.0).pin_memory()
a = argparse.Namespace(model="hyper", quality=3, metric="mse", steps=60, random=1, noise=1e-4, lr_attack=0.01,
                       att_metric="L2", epsilon=16.0, clamp=True, adv=False, force_branch=1)


def tick(label, t0):
    torch.cuda.synchronize()
    t1 = time.perf_counter()
    print(f"  {label:28s} {1e3 * (t1 - t0):8.1f} ms")
    return t1


for rep in range(2):
    print("call", rep)
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    x = host.to(dev, non_blocking=True)
    t = tick("H2D", t0)
    output_s, bpp_ori = patk.clean_pass(x, net, a)
    t = tick("clean_pass", t)
    net.train()
    eng = patk._engine_for(net, x, a)
    t = tick("engine (build/refresh)", t)
    eng.load(x, output_s, None)
    t = tick("load", t)
    eng.run(a.steps)
    t = tick("run 60 iterations", t)
    im_in = eng.im_in_nchw().contiguous()
    res = patk.eval(im_in, x, output_s, net, a)
    t = tick("eval", t)
    out = res[0].to("cpu")
    t = tick("D2H", t)
    print(f"  total {1e3 * (t - t0):.1f} ms -> {n * a.steps / (t - t0):.1f} image-iterations/s")
