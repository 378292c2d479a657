"""Per-launch timing of one fused attack iteration with CUDA events (hyper q3, n x 512x768, forced branch B).
Each launch is timed alone (L2-cold by size at n >= 16), so times add up to slightly more than the graph step.
Usage: python scripts/step_breakdown.py [n_img] [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200.engine import AttackEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev).train()
x = torch.rand(n, 3, 512, 768, device=dev)
ref = torch.rand(n, 3, 512, 768, device=dev)
eng = AttackEngine(net, n, 512, 768, steps=1001, force_branch=1, use_graph=False)
eng.load(x, ref)
eng.run(2)
torch.cuda.synchronize()


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def describe(prog, j, direction):
    u = prog.units[j]
    px_in = prog.hw[j][0] * prog.hw[j][1]
    macs = (px_in * u.k * u.k * u.cin * u.cout) // (u.s * u.s) if not u.transposed else px_in * u.k * u.k * u.cin * u.cout
    if direction == "fwd" and u.gdn is not None:
        macs += prog.hw[j + 1][0] * prog.hw[j + 1][1] * u.cout * u.cout
    if direction == "bwd" and j > 0 and prog.units[j - 1].gdn is not None:
        macs += prog.hw[j][0] * prog.hw[j][1] * u.cin * u.cin
    return macs


rows = []
total = 0.0
for name, prog in (("g_a", eng.ga), ("g_s", eng.gs)):
    for direction, launches in (("fwd", prog.fwd), ("bwd", prog.bwd)):
        order = list(range(len(prog.units))) if direction == "fwd" else list(range(len(prog.units) - 1, -1, -1))
        conv_i = 0
        for p in launches:
            ms = timed(p.launch)
            kind = type(p).__name__
            row = {"stack": name, "dir": direction, "kind": kind, "ms": round(ms, 4)}
            if kind != "PadLaunch":
                j = order[conv_i]
                conv_i += 1
                u = prog.units[j]
                gflop = 2.0 * describe(prog, j, direction) * n / 1e9
                row.update(unit=j, layer=f"{'deconv' if u.transposed else 'conv'} {u.cin}->{u.cout}" +
                           (" +gdn" if (direction == "fwd" and u.gdn is not None) else "") +
                           (" +gdn_bwd" if (direction == "bwd" and j > 0 and prog.units[j - 1].gdn is not None) else ""),
                           gflop=round(gflop, 1), tflops=round(gflop / ms, 1), kernels=p.kernels)
            rows.append(row)
            total += ms
            print(json.dumps(row), flush=True)
print(json.dumps({"n_img": n, "sum_ms": round(total, 3)}))
if len(sys.argv) > 2:
    json.dump(rows, open(sys.argv[2], "w"), indent=1)
