"""Developer check: traced launch program vs the module walk, block by block (forward and input gradient)."""
import sys

import torch
import torch.nn as nn

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200 import tape  # noqa: E402

dev = torch.device("cuda:0")
rel = lambda a, b: float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp(min=1e-30))
N = 128
cases = {
    "conv3x3": (nn.Sequential(pm.conv3x3(N, N)), N),
    "conv3x3 s2": (nn.Sequential(pm.conv3x3(N, N, 2)), N),
    "conv1x1 s2": (nn.Sequential(pm.conv1x1(N, N, 2)), N),
    "conv+leaky+conv": (nn.Sequential(pm.conv3x3(N, N), pm.LeakyReLU(), pm.conv3x3(N, N)), N),
    "ResidualBlock": (nn.Sequential(pm.ResidualBlock(N, N)), N),
    "RBS(N,N)": (nn.Sequential(pm.ResidualBlockWithStride(N, N, 2)), N),
    "RBS(3,N)": (nn.Sequential(pm.ResidualBlockWithStride(3, N, 2)), 3),
    "subpel": (nn.Sequential(pm.subpel_conv3x3(N, N, 2)), N),
    "subpel->3": (nn.Sequential(pm.subpel_conv3x3(N, 3, 2)), N),
    "RBU": (nn.Sequential(pm.ResidualBlockUpsample(N, N, 2)), N),
    "RB+RBS": (nn.Sequential(pm.ResidualBlock(N, N), pm.ResidualBlockWithStride(N, N, 2)), N),
}
for name, (stack, cin) in cases.items():
    torch.manual_seed(1)
    stack = stack.to(dev).train()
    n, h, w = 2, 16, 24
    x = torch.randn(n, cin, h, w, device=dev)
    xi = x.clone().requires_grad_(True)
    out = stack(xi)
    gout = torch.randn_like(out)
    out.backward(gout)
    prog = tape.TapeProgram(stack, n, h, w, dev)
    prog.x_in.copy_(x.permute(0, 2, 3, 1))
    prog.forward()
    prog.g_out.copy_(gout.permute(0, 2, 3, 1))
    prog.backward()
    print(f"{name:18s} fwd {rel(prog.out.permute(0, 3, 1, 2), out.detach()):.2e}  bwd {rel(prog.g_in.permute(0, 3, 1, 2), xi.grad):.2e}  "
          f"nodes {[nd['kind'] + ('*' if nd.get('act') else '') for nd in prog.nodes]}", flush=True)
