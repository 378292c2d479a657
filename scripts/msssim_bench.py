"""MS-SSIM value + gradient at the config-3 shape: row-marching fused kernel vs the two-pass tile kernels.
Usage: python scripts/msssim_bench.py [images]   (CUDA events, best of 5 after warm-up; prints per-kernel split too)"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompression_adversarial_b200 import metrics


def timed(fn, reps=5):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    return best


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
    dev = torch.device("cuda:0")
    g = torch.Generator(device=dev).manual_seed(1)
    a = torch.rand(n, 3, 512, 768, device=dev, generator=g)
    b = (a + 0.05 * torch.randn(n, 3, 512, 768, device=dev, generator=g)).clamp(0, 1)
    up = torch.ones(n, device=dev)
    taps = metrics._taps(11, 1.5)
    px = n * 3 * 512 * 768
    out = {}
    for fused in (True, False):
        metrics.FUSED_VALUE_GRAD = fused
        out[fused] = timed(lambda: metrics.ms_ssim_value_and_grad(a, b, up))
    metrics.FUSED_VALUE_GRAD = True
    t_lvl = timed(lambda: metrics._level_value_grad(a, b, taps, False, 1e-4, 9e-4, False))
    t_fwd = timed(lambda: metrics._level(a, b, taps, False, 1e-4, 9e-4))
    one = torch.ones(n * 3, device=dev); zero = torch.zeros(n * 3, device=dev)
    t_bwd = timed(lambda: metrics._level_bwd(a, b, one, zero, None, (0, 0), taps, 1e-4, 9e-4))
    U = torch.empty_like(a)
    t_cmb = timed(lambda: metrics._combine(U, one, None, (0, 0)))
    v1, g1 = metrics.ms_ssim_value_and_grad(a, b, up)
    metrics.FUSED_VALUE_GRAD = False
    v0, g0 = metrics.ms_ssim_value_and_grad(a, b, up)
    metrics.FUSED_VALUE_GRAD = True
    print(f"images {n}: value+gradient composition fused {out[True]:.3f} ms, two-pass {out[False]:.3f} ms")
    print(f"level 0 ({px/1e6:.1f} Mpx): fused value+grad kernel {t_lvl:.3f} ms = {px/t_lvl/1e6:.1f} Gpx/s "
          f"({20*px/t_lvl/1e6:.0f} GB/s of X,Y read twice + U written); two-pass: forward {t_fwd:.3f} + backward {t_bwd:.3f} ms; "
          f"combine {t_cmb:.3f} ms ({8*px/t_cmb/1e6:.0f} GB/s)")
    print(f"max |value diff| {float((v1-v0).abs().max()):.2e}, gradient rel. max diff {float((g1-g0).abs().max()/g0.abs().max()):.2e}")


if __name__ == "__main__":
    main()
