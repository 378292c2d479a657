"""Per-launch roofline table of one fused attack iteration (the table bench.py embeds), for A/B runs of kernel variants:
   ICADV_TC_STREAM_BWD=0 python scripts/launch_table.py 64
Usage: python scripts/launch_table.py [n_img] [out.json]"""
import json
import sys

import torch

sys.path.insert(0, ".")
import bench  # noqa: E402
from imagecompression_adversarial_b200 import models as pm  # noqa: E402
from imagecompression_adversarial_b200.engine import AttackEngine  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 64
dev = torch.device("cuda:0")
torch.manual_seed(0)
net = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev).train()
x = torch.rand(n, 3, 512, 768, device=dev)
ref = torch.rand(n, 3, 512, 768, device=dev)
eng = AttackEngine(net, n, 512, 768, steps=1001, force_branch=1, use_graph=True)
eng.load(x, ref)
eng.run(5)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
eng.run(20)
e1.record()
torch.cuda.synchronize()
step_ms = e0.elapsed_time(e1) / 20
pk, _ = bench.peaks()
burst, sustained = bench.measured_tf32_peak()
rows = bench.launch_table(eng, pk, burst)
for r in rows:
    print(f"{r['name']:46s} {r['ms']:8.4f} ms  bound {r['bound']:6s} {r['bound_ms']:7.4f}  frac {r['frac']:5.3f}  "
          f"{r['tflops']:7.1f} TF/s {r['gbs']:7.1f} GB/s")
print(json.dumps({"n_img": n, "step_ms": step_ms, "sum_ms": sum(r["ms"] for r in rows), "tf32_burst": burst,
                  "tf32_sustained": sustained}))
if len(sys.argv) > 2:
    json.dump({"step_ms": step_ms, "rows": rows}, open(sys.argv[2], "w"), indent=1)
