"""Developer measurement: cost of an iteration in which NO image takes the network branch (all images on the budget branch:
every network launch is replayed from the graph and exits at once) against a forced network-branch iteration."""
import sys
import torch
sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import models as pm
from imagecompression_adversarial_b200.engine import AttackEngine

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev).train()
x = torch.rand(n, 3, 512, 768, device=dev)
ref = torch.rand(n, 3, 512, 768, device=dev)
for force, name in ((1, "network branch forced (B)"), (0, "budget branch forced (A): network launches exit at once")):
    eng = AttackEngine(net, n, 512, 768, steps=1001, force_branch=force, use_graph=True)
    eng.load(x, ref)
    eng.run(5)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(50); e1.record()
    torch.cuda.synchronize()
    print(f"{n} images, {name}: {e0.elapsed_time(e1) / 50:.3f} ms per iteration")

# un-forced loop: the graph's IF node skips the network section of budget-branch iterations (csrc/icadv_graph.cu)
import os
for flag in ("0", "1"):
    os.environ["ICADV_GRAPH_IF"] = flag
    eng = AttackEngine(net, n, 512, 768, steps=1001, use_graph=True)
    eng.load(x, ref)
    eng.run(20)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); eng.run(200); e1.record()
    torch.cuda.synchronize()
    print(f"{n} images, un-forced loop (iterations 20..219 of a 1001-step schedule), ICADV_GRAPH_IF={flag}: "
          f"{e0.elapsed_time(e1) / 200:.3f} ms per iteration")
