"""Short eager (no CUDA graph) run of the fused attack iteration for ncu: 8 images 512x768, hyper q3,
forced branch B, 1 warm-up + 2 iterations."""
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200.engine import AttackEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 3
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev).train()
x = torch.rand(n, 3, 512, 768, device=dev)
ref = torch.rand(n, 3, 512, 768, device=dev)
eng = AttackEngine(net, n, 512, 768, steps=1001, force_branch=1, use_graph=False)
eng.load(x, ref)
eng.run(iters)
torch.cuda.synchronize()
print("kernels/iteration", eng.kernels_per_iteration(), "loss_i", eng.st.loss_i[:2].tolist())
