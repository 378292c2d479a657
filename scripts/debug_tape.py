"""Developer check: per-node difference between a traced launch program and the module walk it was traced from."""
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import functional as Fn  # noqa: E402
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200 import tape  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("cheng2020", 1, "mse", pretrained=False).to(dev).train()
n, h, w = 2, 64, 96
x = torch.rand(n, 3, h, w, device=dev).contiguous(memory_format=torch.channels_last)
# eager walk with the recorder on the REAL input: keeps every intermediate
rec = tape.Recorder()
in_id = rec.tid(Fn.to_nhwc(x))
Fn._REC = rec
with torch.no_grad():
    out = net.g_a(x)
Fn._REC = None
eager = {i: t for t, i in zip(rec.keep, range(len(rec.keep)))}
prog = tape.TapeProgram(net.g_a, n, h, w, dev)
prog.x_in.copy_(x.permute(0, 2, 3, 1))
prog.forward()
rel = lambda a, b: float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp(min=1e-30))
# ids are assigned in first-seen order in both traces (same module walk): compare buffer by buffer
_, alias = tape.TapeProgram._fuse_activations([dict(nd) for nd in rec.nodes])
for nd in rec.nodes:
    tid = alias.get(nd["out"], nd["out"])
    if nd["out"] in alias:
        continue
    if tid in prog.buf:
        # a fused activation: the program's conv buffer holds the post-activation value
        want = eager[nd["out"]]
        fused_act = [k for k, v in alias.items() if v == nd["out"]]
        if fused_act:
            want = eager[fused_act[0]]
        print(f"{nd['kind']:8s} out={nd['out']:3d} shape={tuple(want.shape)} rel={rel(prog.buf[tid], want):.3e}"
              f"{' (fused act)' if fused_act else ''}")
