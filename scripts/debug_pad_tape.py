"""Developer check: channel-padded contractions of the traced program (forward + input gradient) against the module walk,
on a two-layer stack (well conditioned) and on a tamed cheng2020."""
import sys
import torch
import torch.nn as nn
sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm
from imagecompression_adversarial_b200.tape import TapeProgram

dev = torch.device("cuda:0")
rel = lambda a, b: float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt())


def check(stack, n, h, w, name):
    x = torch.rand(n, 3, h, w, device=dev)
    xi = x.clone().requires_grad_(True)
    out = stack(xi)
    gout = torch.randn_like(out)
    out.backward(gout)
    tp = TapeProgram(stack, n, h, w, dev)
    tp.x_in.copy_(x.permute(0, 2, 3, 1))
    tp.forward()
    tp.g_out.copy_(gout.permute(0, 2, 3, 1))
    tp.backward()
    print(name, "fwd", rel(tp.out.permute(0, 3, 1, 2), out.detach()), "dgrad", rel(tp.g_in.permute(0, 3, 1, 2), xi.grad),
          "| out rms", float(out.pow(2).mean().sqrt()), "grad rms", float(xi.grad.pow(2).mean().sqrt()))


torch.manual_seed(0)
s1 = nn.Sequential(pm.conv3x3(3, 64, 2), pm.subpel_conv3x3(64, 3, 2)).to(dev).train()
check(s1, 2, 64, 96, "two-layer (3->64 /2, 64->12 + shuffle)")
s2 = nn.Sequential(pm.ResidualBlockWithStride(3, 64, 2), pm.ResidualBlockUpsample(64, 64, 2), pm.subpel_conv3x3(64, 3, 1)).to(dev).train()
check(s2, 2, 64, 96, "residual blocks with GDN / IGDN")
for scale in (1.0, 0.6):
    torch.manual_seed(0)
    net = pm.init_model("cheng2020", 1, "mse", pretrained=False).to(dev).train()
    with torch.no_grad():
        for name, m in net.named_modules():
            if hasattr(m, "weight") and m.weight is not None and m.weight.dim() == 4 and name.startswith(("g_a", "g_s")):
                m.weight.mul_(scale)
    full = nn.Sequential(net.g_a, net.g_s)
    check(net.g_a, 2, 64, 96, f"cheng2020 g_a, weights x{scale}")

# tape and module walk against the fp32 oracle (cuDNN, TF32 off) on the tamed cheng2020 g_a
from oracle import models as om
torch.backends.cudnn.allow_tf32 = False
torch.backends.cuda.matmul.allow_tf32 = False
onet = om.init_model("cheng2020", 1, seed=0).to(dev).train()
with torch.no_grad():
    for name, m in onet.named_modules():
        if isinstance(m, nn.Conv2d) and name.startswith(("g_a", "g_s")):
            m.weight.mul_(0.6)
pnet = pm.init_model("cheng2020", 1, "mse", pretrained=False).to(dev).train()
pnet.load_state_dict(onet.state_dict())
n, h, w = 2, 64, 96
x = torch.rand(n, 3, h, w, device=dev)
gout = None
res = {}
for nm, net in (("oracle", onet), ("walk", pnet)):
    xi = x.clone().requires_grad_(True)
    out = net.g_a(xi)
    if gout is None:
        gout = torch.randn_like(out)
    out.backward(gout)
    res[nm] = (out.detach(), xi.grad.detach())
tp = TapeProgram(pnet.g_a, n, h, w, dev)
tp.x_in.copy_(x.permute(0, 2, 3, 1)); tp.forward(); tp.g_out.copy_(gout.permute(0, 2, 3, 1)); tp.backward()
res["tape"] = (tp.out.permute(0, 3, 1, 2), tp.g_in.permute(0, 3, 1, 2))
for nm in ("walk", "tape"):
    print(nm, "vs oracle: fwd", rel(res[nm][0], res["oracle"][0]), "dgrad", rel(res[nm][1], res["oracle"][1]))
print("tape vs walk: fwd", rel(res["tape"][0], res["walk"][0]), "dgrad", rel(res["tape"][1], res["walk"][1]))
