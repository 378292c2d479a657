"""Developer tool: run the g_s.6 input-gradient + IGDN-backward launch (rgb_in form) a few times, for ncu captures.
Usage: python scripts/one_bwd_layer.py [n_img] [rgb|tconv|sconv]"""
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
which = sys.argv[2] if len(sys.argv) > 2 else "rgb"
H, W, Cc = 512, 768, 128
gen = lambda s: torch.Generator(device=dev).manual_seed(s)
gm = (0.1 * torch.eye(Cc, device=dev)).contiguous()
yp = torch.randn(n, H // 2, W // 2, Cc, device=dev, generator=gen(3))
sp = 0.5 + torch.rand(n, H // 2, W // 2, Cc, device=dev, generator=gen(4))
gin = torch.empty_like(yp)
if which == "rgb":
    gx = torch.randn(n, H, W, 3, device=dev, generator=gen(7))
    padg = ops.pad_rgb4(gx, ops.alloc_pad4(n, H, W, dev))
    wr = ops.pack_weight_rgb(torch.randn(Cc, 3, 5, 5, device=dev, generator=gen(6)) / 9)
    d = ops.make_desc(padg, wr, None, gin, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_BWD, gmat=gm,
                      y_prev=yp, sc_prev=sp, in_pad4=True, round_out=True)
    keep = (padg, wr, gin, yp, sp)
else:
    x = torch.randn(n, H // 4, W // 4, Cc, device=dev, generator=gen(1))
    w = torch.randn(25, Cc, Cc, device=dev, generator=gen(2)) / 56
    d = ops.make_desc(x, w, None, gin, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_GDN_BWD, gmat=gm,
                      y_prev=yp, sc_prev=sp, round_out=True)
    keep = (x, w, gin, yp, sp)
plan = ops.ConvPlan(d, keep)
for _ in range(3):
    plan.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5):
    plan.launch()
e1.record()
torch.cuda.synchronize()
print(f"{which} n={n}: {e0.elapsed_time(e1) / 5:.4f} ms per launch")
