"""Developer tool: per-CTA phase timeline of conv_tc_kernel (clock64 stamps) for a few layer shapes."""
import math
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import ops  # noqa: E402
import ctypes as C  # noqa: E402

dev = torch.device("cuda:0")
n = 8
names = ["start", "setup", "first_full", "main_issued", "gdn_issued", "epi_acc1", "pass1_done", "epi_acc2",
         "pass2_done", "stores_done", "end", "mma_wait_w", "mma_wait_p", "prod_wait_empty", "prod_loop", "c0_loaded", "c0_bar1", "c0_written", "c0_bar2", "c0_issued", "c2_loaded",
         "c2_storewait", "c2_bar2", "c2_issued", "-", "-", "-", "-", "acc_q0", "acc_q1", "acc_q2", "acc_q3", "ld_q0", "ld_q1", "ld_q2",
         "ld_q3"]


ONLY = sys.argv[1] if len(sys.argv) > 1 else ""


def run(name, d, keep, grid):
    if ONLY and ONLY not in name:
        return
    plan = ops.ConvPlan(d, keep)
    dbg = torch.zeros(2 * grid * 32, dtype=torch.int64, device=dev)
    L.call("icadv_conv_plan_set_debug", plan._h, C.c_void_p(dbg.data_ptr()))
    for _ in range(2):
        plan.launch()
    torch.cuda.synchronize()
    t = dbg.view(-1, 32).double().cpu()
    t = t[t[:, 0] > 0]
    rel = (t[:, :32] - t[:, :1])
    med = rel.median(0).values
    print(name, "CTAs", t.shape[0])
    print("   " + "  ".join(f"{nm}={med[i]/1.9e3:6.2f}us" for i, nm in enumerate(names[:15])))


H, W = 512, 768
# first layer: rgb_in + GDN
x = torch.rand(n, H, W, 3, device=dev)
pad = ops.pad_rgb4(x, ops.alloc_pad4(n, H, W, dev))
w = torch.randn(128, 3, 5, 5, device=dev) / 9
wp = ops.pack_weight_rgb(w)
gm = (0.1 * torch.eye(128, device=dev)).contiguous()
beta = torch.ones(128, device=dev)
bias = torch.zeros(128, device=dev)
out = torch.empty(n, H // 2, W // 2, 128, device=dev)
sc = torch.empty_like(out)
d = ops.make_desc(pad, wp, bias, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=128, epi=L.EPI_GDN_FWD, gmat=gm,
                  beta=beta, out_scale=sc, in_pad4=True)
run("g_a.0 rgb_in + GDN", d, (pad, wp, out, sc, gm, beta, bias), 768 * n)
d = ops.make_desc(pad, wp, bias, out, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=128, in_pad4=True)
run("g_a.0 rgb_in linear", d, (pad, wp, out, bias), 768 * n)
# g_a.2
x2 = out
w2 = torch.randn(25, 128, 128, device=dev) / 56
out2 = torch.empty(n, H // 4, W // 4, 128, device=dev)
sc2 = torch.empty_like(out2)
d = ops.make_desc(x2, w2, bias, out2, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=128, epi=L.EPI_GDN_FWD, gmat=gm,
                  beta=beta, out_scale=sc2)
run("g_a.2 conv + GDN", d, (x2, w2, out2, sc2), 192 * n)
d = ops.make_desc(x2, w2, bias, out2, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=128)
run("g_a.2 conv linear", d, (x2, w2, out2), 192 * n)
g2 = torch.randn_like(out2)
gin = torch.empty_like(x2)
d = ops.make_desc(g2, w2, None, gin, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=128, epi=L.EPI_GDN_BWD, gmat=gm,
                  y_prev=out, sc_prev=sc)
run("g_a.2 dgrad phase(0,0) + GDN bwd", d, (g2, w2, gin), 192 * n)
# col2im last layer
x5 = torch.randn(n, H // 2, W // 2, 128, device=dev)
w5 = torch.randn(25, 3, 128, device=dev) / 30
o5 = torch.empty(n, H, W, 3, device=dev)
b5 = torch.zeros(3, device=dev)
d = ops.make_desc(x5, w5, b5, o5, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=3)
run("g_s.6 col2im", d, (x5, w5, o5, b5), 1204 * n)
# g_s.6 dgrad: rgb_in conv + IGDN backward (per-tile kernel, two CTAs per SM)
yp = torch.randn(n, H // 2, W // 2, 128, device=dev)
sp = 0.5 + torch.rand(n, H // 2, W // 2, 128, device=dev)
go = torch.empty(n, H // 2, W // 2, 128, device=dev)
d = ops.make_desc(pad, wp, None, go, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=128, epi=L.EPI_IGDN_BWD, gmat=gm,
                  y_prev=yp, sc_prev=sp, in_pad4=True)
run("g_s.6 dgrad rgb_in + IGDN bwd", d, (pad, wp, go, yp, sp), 768 * n)
# g_s.4 dgrad: stride-2 conv + IGDN backward
yp2 = torch.randn(n, H // 4, W // 4, 128, device=dev)
sp2 = 0.5 + torch.rand(n, H // 4, W // 4, 128, device=dev)
go2 = torch.empty(n, H // 4, W // 4, 128, device=dev)
d = ops.make_desc(x2, w2, None, go2, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=128, epi=L.EPI_IGDN_BWD, gmat=gm,
                  y_prev=yp2, sc_prev=sp2)
run("g_s.4 dgrad conv + IGDN bwd", d, (x2, w2, go2, yp2, sp2), 192 * n)
