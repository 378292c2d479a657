"""Developer tool: where a persistent-kernel CTA spends its cycles (per-role counters), for the hyper-q3 layer shapes.
Columns (median over CTAs, microseconds at 1.9 GHz): whole main-loop span of the MMA warp, of which blocked on the
TMA rings / on a TMEM buffer still owned by an epilogue group; per epilogue group: waiting for an accumulator,
pass 1 (+ normalisation wait), pass 2."""
import math
import sys
import ctypes as C

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
n = int(sys.argv[1]) if len(sys.argv) > 1 else 16
H, W, Cc = 512, 768, 128
gen = lambda s: torch.Generator(device=dev).manual_seed(s)
gm = (0.1 * torch.eye(Cc, device=dev)).contiguous()
beta = torch.ones(Cc, device=dev)


def run(name, d, keep):
    plan = ops.ConvPlan(d, keep)
    dbg = torch.zeros(148 * 16, dtype=torch.int64, device=dev)
    L.call("icadv_conv_plan_set_debug", plan._h, C.c_void_p(dbg.data_ptr()))
    plan.launch()
    torch.cuda.synchronize()
    dbg.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); plan.launch(); e1.record()
    torch.cuda.synchronize()
    t = dbg.view(148, 16).double().cpu()
    t = t[t[:, 7] > 0]
    med = t.median(0).values / 1.9e3
    items = float(t[:, 7].median())
    print(f"{name}: {e0.elapsed_time(e1):.3f} ms, {items:.0f} items/CTA")
    print(f"   MMA warp: main loops {med[0]:8.1f} us (rings {med[1]:7.1f}, TMEM-buffer wait {med[2]:7.1f}), whole {med[8]:8.1f} us")
    print(f"   of the ring wait: {med[4]:.1f} us on the patch ring")
    for g in range(2):
        print(f"   epilogue group {g}: wait acc {med[9 + 3 * g]:8.1f}  pass1(+norm wait) {med[10 + 3 * g]:8.1f}  pass2 {med[11 + 3 * g]:8.1f} us")
    print(f"   group 0 waiting for the last normalisation MMA: {med[3]:.1f} us;  blocked on the saved-tensor ring (streaming "
          f"backward kernel): group 0 {med[5]:.1f} us, group 1 {med[6]:.1f} us; store drain before the first operand write: "
          f"group 0 {med[15]:.1f} us")


# g_s.4: deconv 128->128 + IGDN forward at 128x192 -> 256x384
x = torch.randn(n, H // 4, W // 4, Cc, device=dev, generator=gen(1))
w = torch.randn(25, Cc, Cc, device=dev, generator=gen(2)) / 56
out = torch.empty(n, H // 2, W // 2, Cc, device=dev); sc = torch.empty_like(out)
d = ops.make_desc(x, w, beta, out, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_FWD, gmat=gm, beta=beta, out_scale=sc)
run("g_s.4 deconv + IGDN fwd (persistent)", d, (x, w, out, sc))
# g_a.2 dgrad: TCONV + GDN backward
yp = torch.randn(n, H // 2, W // 2, Cc, device=dev, generator=gen(3)); sp = 0.5 + torch.rand(n, H // 2, W // 2, Cc, device=dev, generator=gen(4))
gin = torch.empty_like(yp)
d = ops.make_desc(x, w, None, gin, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_GDN_BWD, gmat=gm, y_prev=yp, sc_prev=sp)
run("g_a.2 dgrad + GDN bwd (persistent)", d, (x, w, gin, yp, sp))
# g_a.0: rgb_in + GDN forward
xi = torch.rand(n, H, W, 3, device=dev, generator=gen(5))
pad = ops.pad_rgb4(xi, ops.alloc_pad4(n, H, W, dev))
wr = ops.pack_weight_rgb(torch.randn(Cc, 3, 5, 5, device=dev, generator=gen(6)) / 9)
d = ops.make_desc(pad, wr, beta, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_GDN_FWD, gmat=gm, beta=beta, out_scale=sc, in_pad4=True)
run("g_a.0 rgb_in + GDN fwd (persistent)", d, (pad, wr, out, sc))
# g_s.6 dgrad: rgb_in + IGDN backward (the HBM-bound launch the streaming backward kernel exists for)
gx = torch.randn(n, H, W, 3, device=dev, generator=gen(7))
padg = ops.pad_rgb4(gx, ops.alloc_pad4(n, H, W, dev))
d = ops.make_desc(padg, wr, None, gin, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_BWD, gmat=gm, y_prev=yp, sc_prev=sp, in_pad4=True)
run("g_s.6 dgrad rgb_in + IGDN bwd", d, (padg, wr, gin, yp, sp))
# g_s.4 dgrad: conv 5x5/2 + IGDN backward
g4 = torch.randn(n, H // 2, W // 2, Cc, device=dev, generator=gen(8))
yq = torch.randn(n, H // 4, W // 4, Cc, device=dev, generator=gen(9)); sq = 0.5 + torch.rand(n, H // 4, W // 4, Cc, device=dev, generator=gen(10))
go = torch.empty_like(yq)
d = ops.make_desc(g4, w, None, go, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=Cc, epi=L.EPI_IGDN_BWD, gmat=gm, y_prev=yq, sc_prev=sq)
run("g_s.4 dgrad conv + IGDN bwd", d, (g4, w, go, yq, sq))
# g_s.6 fwd: deconv 128->3 5x5/2 (col2im epilogue) at 256x384 -> 512x768
x6 = torch.randn(n, H // 2, W // 2, Cc, device=dev, generator=gen(11))
w6 = torch.randn(25, 3, Cc, device=dev, generator=gen(12)) / 56
b6 = torch.zeros(3, device=dev)
o6 = torch.empty(n, H, W, 3, device=dev)
d = ops.make_desc(x6, w6, b6, o6, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=3)
run("g_s.6 deconv 128->3 fwd (col2im, persistent)", d, (x6, w6, b6, o6))
