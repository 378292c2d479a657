"""Developer tool: launch ONE hyper-q3 layer shape a few times (for `ncu --set full -k regex:conv_tc`), no debug counters.
Usage: python scripts/one_layer.py <gs4_fwd|ga2_bwd|ga0_fwd|ga2_fwd|gs4_bwd|gs6_bwd|gs6_fwd> [n_img] [launches]"""
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import ops  # noqa: E402

which = sys.argv[1]
n = int(sys.argv[2]) if len(sys.argv) > 2 else 16
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 3
dev = torch.device("cuda:0")
H, W, Cc = 512, 768, 128
gen = lambda s: torch.Generator(device=dev).manual_seed(s)
gm = (0.1 * torch.eye(Cc, device=dev)).contiguous()
beta = torch.ones(Cc, device=dev)
w = torch.randn(25, Cc, Cc, device=dev, generator=gen(2)) / 56
big = lambda s: torch.randn(n, H // 2, W // 2, Cc, device=dev, generator=gen(s))
small = lambda s: torch.randn(n, H // 4, W // 4, Cc, device=dev, generator=gen(s))
if which == "gs4_fwd":     # deconv 128->128 + IGDN forward, 128x192 -> 256x384
    x, out, sc = small(1), big(7), big(8)
    d = ops.make_desc(x, w, beta, out, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_FWD, gmat=gm, beta=beta, out_scale=sc)
elif which == "gs4_lin":   # deconv 128->128, linear epilogue (developer experiments)
    x, out = small(1), big(7)
    d = ops.make_desc(x, w, beta, out, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_LINEAR)
elif which == "ga2_bwd":   # dgrad of conv 128->128 (a transposed conv) + GDN backward of the layer below
    x, yp, sp, out = small(1), big(3), 0.5 + torch.rand(n, H // 2, W // 2, Cc, device=dev, generator=gen(4)), big(9)
    d = ops.make_desc(x, w, None, out, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_GDN_BWD, gmat=gm, y_prev=yp, sc_prev=sp)
elif which == "ga2_fwd":   # conv 128->128 5x5/2 + GDN forward, 256x384 -> 128x192
    x, out, sc = big(1), small(7), small(8)
    d = ops.make_desc(x, w, beta, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_GDN_FWD, gmat=gm, beta=beta, out_scale=sc)
elif which == "gs4_bwd":   # dgrad of deconv 128->128 (a stride-2 conv) + IGDN backward of the layer below
    x, yp, sp, out = big(1), small(3), 0.5 + torch.rand(n, H // 4, W // 4, Cc, device=dev, generator=gen(4)), small(9)
    d = ops.make_desc(x, w, None, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_BWD, gmat=gm, y_prev=yp, sc_prev=sp)
elif which in ("ga0_fwd", "gs6_bwd"):   # RGB first-layer form: conv 3->128 5x5/2 (+ GDN fwd | + IGDN bwd)
    xi = torch.rand(n, H, W, 3, device=dev, generator=gen(5))
    pad = ops.pad_rgb4(xi, ops.alloc_pad4(n, H, W, dev))
    wr = ops.pack_weight_rgb(torch.randn(Cc, 3, 5, 5, device=dev, generator=gen(6)) / 9)
    out, sc = big(7), 0.5 + torch.rand(n, H // 2, W // 2, Cc, device=dev, generator=gen(4))
    if which == "ga0_fwd":
        d = ops.make_desc(pad, wr, beta, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_GDN_FWD, gmat=gm, beta=beta, out_scale=sc, in_pad4=True)
    else:
        yp = big(3)
        d = ops.make_desc(pad, wr, None, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_BWD, gmat=gm, y_prev=yp, sc_prev=sc, in_pad4=True)
elif which == "gs6_fwd":   # col2im: deconv 128->3
    x = big(1)
    w3 = torch.randn(25, 3, Cc, device=dev, generator=gen(2)) / 56
    out = torch.empty(n, H, W, 3, device=dev)
    d = ops.make_desc(x, w3, None, out, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=3, epi=L.EPI_LINEAR)
else:
    raise SystemExit(f"unknown layer {which}")
plan = ops.ConvPlan(d, None)
plan.launch()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    plan.launch()
e1.record()
torch.cuda.synchronize()
print(f"{which} n={n}: {e0.elapsed_time(e1) / reps:.3f} ms per launch ({plan.kernels} kernel(s))")
