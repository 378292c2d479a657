"""Developer check: is the fused loop bit-reproducible across runs IN one process and ACROSS processes?
Prints a checksum of the perturbation after 3 forced-network iterations (hyper q3), for two image sizes."""
import sys
import hashlib
import torch
sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm
from imagecompression_adversarial_b200.engine import AttackEngine

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev).train()
for (n, h, w) in ((1, 64, 64), (2, 192, 256)):
    g = torch.Generator(device=dev).manual_seed(1)
    x = torch.rand(n, 3, h, w, device=dev, generator=g)
    ref = torch.rand(n, 3, h, w, device=dev, generator=g)
    sums = []
    for rep in range(3):
        eng = AttackEngine(net, n, h, w, steps=6, force_branch=1, use_graph=False)
        eng.load(x, ref)
        eng.run(3)
        torch.cuda.synchronize()
        sums.append(hashlib.md5(eng.noise.cpu().numpy().tobytes()).hexdigest()[:10])
        del eng
    print(f"{n}x{h}x{w}:", " ".join(sums))
