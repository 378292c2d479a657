"""Developer tool: run each tensor-path configuration several times on identical inputs and compare bitwise."""
import math
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import ops  # noqa: E402

dev = torch.device("cuda:0")
torch.manual_seed(0)


def check(name, fn, reps=6):
    outs = [fn() for _ in range(reps)]
    torch.cuda.synchronize()
    bad = 0
    for o in outs[1:]:
        for a, b in zip(outs[0], o):
            if not torch.equal(a, b):
                bad += 1
                d = (a - b).abs()
                print(f"   MISMATCH {name}: max diff {float(d.max()):.3e}, n diff {int((d > 0).sum())} of {d.numel()}")
                break
    print(("NONDETERMINISTIC " if bad else "ok ") + name)


for n, H, W in ((2, 32, 48), (8, 128, 192)):
    C = 128
    gm = (0.1 * torch.eye(C, device=dev) + 0.01 * torch.rand(C, C, device=dev)).contiguous()
    beta = 0.5 + torch.rand(C, device=dev)
    bias = torch.randn(C, device=dev)
    x = torch.randn(n, H, W, C, device=dev)
    w = torch.randn(25, C, C, device=dev) / 56
    for epi, nm in ((L.EPI_LINEAR, "lin"), (L.EPI_GDN_FWD, "gdn"), (L.EPI_IGDN_FWD, "igdn")):
        for form, fn_ in ((L.FORM_SCONV, "sconv"), (L.FORM_TCONV, "tconv")):
            kw = dict(gmat=gm, beta=beta) if epi != L.EPI_LINEAR else {}
            def f(epi=epi, form=form, kw=kw):
                r = ops.conv(x, w, bias, form=form, ksize=5, stride=2, n_ch=C, epi=epi, path="tc", **kw)
                return r if isinstance(r, tuple) else (r,)
            check(f"{nm} {fn_} n={n} {H}x{W}", f)
    # backward epilogues need saved y / sc of the matching geometry
    for form, fn_ in ((L.FORM_SCONV, "sconv"), (L.FORM_TCONV, "tconv")):
        oh, ow = ops.out_hw(form, 5, 2, H, W)
        y = torch.randn(n, oh, ow, C, device=dev)
        sc = 0.5 + torch.rand(n, oh, ow, C, device=dev)
        for epi, nm in ((L.EPI_GDN_BWD, "gdn_bwd"), (L.EPI_IGDN_BWD, "igdn_bwd")):
            def f(epi=epi, form=form, y=y, sc=sc):
                return (ops.conv(x, w, None, form=form, ksize=5, stride=2, n_ch=C, epi=epi, gmat=gm, y_prev=y,
                                 sc_prev=sc, path="tc"),)
            check(f"{nm} {fn_} n={n} {H}x{W}", f)
    # rgb_in and col2im
    img = torch.rand(n, H, W, 3, device=dev)
    pad = ops.pad_rgb4(img, ops.alloc_pad4(n, H, W, dev))
    wr = ops.pack_weight_rgb(torch.randn(C, 3, 5, 5, device=dev) / 9)
    check(f"rgb_in gdn n={n}", lambda: ops.conv(pad, wr, bias, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C,
                                               epi=L.EPI_GDN_FWD, gmat=gm, beta=beta, in_pad4=True, path="tc"))
    y = torch.randn(n, H // 2, W // 2, C, device=dev)
    sc = 0.5 + torch.rand(n, H // 2, W // 2, C, device=dev)
    check(f"rgb_in igdn_bwd n={n}", lambda: (ops.conv(pad, wr, None, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C,
                                                      epi=L.EPI_IGDN_BWD, gmat=gm, y_prev=y, sc_prev=sc, in_pad4=True,
                                                      path="tc"),))
    w3 = torch.randn(25, 3, C, device=dev) / 30
    check(f"col2im n={n}", lambda: (ops.conv(x, w3, None, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=3, path="tc"),))
