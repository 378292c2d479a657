"""Driver for the ncu capture of the bandwidth-bound kernels (perturb_forward / perturb_update_adam / output_loss /
pad_rgb4 / eb_forward / gc_forward / ssim_level / ssim_level_backward): 16 images 512x768, hyper q3.
Usage (under gpurun):  python scripts/profile_aux_kernels.py  &&  ncu --set full --clock-control none \
    -k regex:'perturb_|output_loss|pad_rgb4|eb_forward|gc_forward|ssim_level' -c 16 -o gpurun_out/aux python scripts/profile_aux_kernels.py"""
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import metrics  # noqa: E402
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200.engine import AttackEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev)
x = torch.rand(n, 3, 512, 768, device=dev)
ref = (x + 0.02 * torch.randn_like(x)).clamp(0, 1)
net.train()
eng = AttackEngine(net, n, 512, 768, steps=1001, force_branch=1, use_graph=False)
eng.load(x, ref)
eng.run(2)                      # perturb_forward, pad_rgb4, output_loss, perturb_update_adam (+ the contractions)
net.eval()
with torch.no_grad():
    net(x)                      # eb_forward, gc_forward (eval mode)
up = torch.ones(n, device=dev)
metrics.ms_ssim_value_and_grad(x, ref, up)     # ssim_level, avgpool2, ssim_level_backward (5 levels)
torch.cuda.synchronize()
print("ok")
