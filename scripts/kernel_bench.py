"""Micro-benchmark of the contraction kernels at hyperprior-q3 layer shapes (CUDA events, L2-cold
by size).  Usage: python scripts/kernel_bench.py [n_img]"""
import json
import math
import sys

import torch

sys.path.insert(0, ".")
from imagecompression_adversarial_b200 import _lib as L  # noqa: E402
from imagecompression_adversarial_b200 import ops  # noqa: E402


def time_it(fn, iters=5, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 8
    dev = torch.device("cuda:0")
    ops.require_device()
    H, W = 512, 768
    rows = []
    cases = [
        # name, form, kind, K, N, in_h, in_w, epi, path
        ("g_a.0 conv 3->128 (simt)", L.FORM_SCONV, 0, 3, 128, H, W, L.EPI_LINEAR, "simt"),
        ("g_a.1 GDN standalone", L.FORM_SCONV, None, 128, 128, H // 2, W // 2, L.EPI_GDN_FWD, "tc"),
        ("g_a.2 conv 128->128 + GDN", L.FORM_SCONV, 0, 128, 128, H // 2, W // 2, L.EPI_GDN_FWD, "tc"),
        ("g_a.2 conv 128->128 linear", L.FORM_SCONV, 0, 128, 128, H // 2, W // 2, L.EPI_LINEAR, "tc"),
        ("g_a.4 conv 128->128 + GDN", L.FORM_SCONV, 0, 128, 128, H // 4, W // 4, L.EPI_GDN_FWD, "tc"),
        ("g_a.6 conv 128->192", L.FORM_SCONV, 0, 128, 192, H // 8, W // 8, L.EPI_LINEAR, "tc"),
        ("g_s.0 deconv 192->128 + IGDN", L.FORM_TCONV, 2, 192, 128, H // 16, W // 16, L.EPI_IGDN_FWD, "tc"),
        ("g_s.2 deconv 128->128 + IGDN", L.FORM_TCONV, 2, 128, 128, H // 8, W // 8, L.EPI_IGDN_FWD, "tc"),
        ("g_s.4 deconv 128->128 + IGDN", L.FORM_TCONV, 2, 128, 128, H // 4, W // 4, L.EPI_IGDN_FWD, "tc"),
        ("g_s.4 deconv 128->128 linear", L.FORM_TCONV, 2, 128, 128, H // 4, W // 4, L.EPI_LINEAR, "tc"),
        ("g_s.6 deconv 128->3 (simt narrow)", L.FORM_TCONV, 2, 128, 3, H // 2, W // 2, L.EPI_LINEAR, "simt"),
        ("g_s.6 dgrad conv 3->128 (simt)", L.FORM_SCONV, 3, 3, 128, H, W, L.EPI_LINEAR, "simt"),
        ("g_a.0 dgrad deconv 128->3 (simt narrow)", L.FORM_TCONV, 1, 128, 3, H // 2, W // 2, L.EPI_LINEAR, "simt"),
    ]
    for name, form, kind, K, N, ih, iw, epi, path in cases:
        x = torch.randn(n, ih, iw, K, device=dev)
        acc_from_in = kind is None
        wp = None if acc_from_in else torch.randn(25, N, K, device=dev) / math.sqrt(25 * K)
        bias = None if acc_from_in else torch.zeros(N, device=dev)
        gm = (0.1 * torch.eye(N, device=dev)).contiguous() if epi != L.EPI_LINEAR else None
        beta = torch.ones(N, device=dev) if epi != L.EPI_LINEAR else None
        oh, ow = ops.out_hw(form, 1 if acc_from_in else 5, 1 if acc_from_in else 2, ih, iw)
        out = torch.empty(n, oh, ow, N, device=dev)
        sc = torch.empty_like(out) if epi != L.EPI_LINEAR else None
        d = ops.make_desc(x, wp, bias, out, form=form, ksize=1 if acc_from_in else 5, stride=1 if acc_from_in else 2,
                          n_ch=N, epi=epi, gmat=gm, beta=beta, out_scale=sc, acc_from_in=acc_from_in)
        keep = (x, wp, bias, out, sc, gm, beta)
        plan = ops.ConvPlan(d, keep) if path == "tc" else ops.SimtLaunch(d, keep)
        ms = time_it(plan.launch)
        taps = 1 if acc_from_in else 25
        px = n * (oh * ow if form == L.FORM_SCONV else ih * iw)
        macs = px * N * K * taps if not acc_from_in else 0
        if epi != L.EPI_LINEAR:
            macs += n * oh * ow * N * N
        flops = 2.0 * macs
        byts = 4.0 * (x.numel() + out.numel() * (2 if sc is not None else 1))
        rows.append(dict(kernel=name, n_img=n, ms=round(ms, 4), tflops=round(flops / ms / 1e9, 2),
                         gbs=round(byts / ms / 1e6, 1)))
        print(json.dumps(rows[-1]), flush=True)
    json.dump(rows, open("gpurun_out/kernel_bench.json", "w"), indent=1)


if __name__ == "__main__":
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    main()
