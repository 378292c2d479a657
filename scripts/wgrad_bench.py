"""Times the weight gradient of every layer of the config-5 codec update (hyper q1, 8 x 256x256) on both kernels:
the tcgen05 kernel (icadv_wgrad_tc.cu) and the fp32 CUDA-core kernel.  Prints one line per layer."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompression_adversarial_b200 import _lib as L, ops

dev = torch.device("cuda:0")
N = int(os.environ.get("WG_N", 8)); H = int(os.environ.get("WG_H", 256)); W = int(os.environ.get("WG_W", 256))
layers = [  # name, form, cin, cout, k, s, in_h divisor
    ("g_a.0", 0, 3, 128, 5, 2, 1), ("g_a.2", 0, 128, 128, 5, 2, 2), ("g_a.4", 0, 128, 128, 5, 2, 4), ("g_a.6", 0, 128, 192, 5, 2, 8),
    ("h_a.0", 0, 192, 128, 3, 1, 16), ("h_a.2", 0, 128, 128, 5, 2, 16), ("h_a.4", 0, 128, 128, 5, 2, 32),
    ("h_s.0", 1, 128, 128, 5, 2, 64), ("h_s.2", 1, 128, 128, 5, 2, 32), ("h_s.4", 0, 128, 192, 3, 1, 16),
    ("g_s.0", 1, 192, 128, 5, 2, 16), ("g_s.2", 1, 128, 128, 5, 2, 8), ("g_s.4", 1, 128, 128, 5, 2, 4), ("g_s.6", 1, 128, 3, 5, 2, 2),
]
tot = {"tc": 0.0, "simt": 0.0}
for name, tr, cin, cout, k, s, div in layers:
    h, w = H // div, W // div
    x = torch.randn(N, h, w, cin, device=dev)
    form = L.FORM_TCONV if tr else L.FORM_SCONV
    oh, ow = ops.out_hw(form, k, s, h, w)
    g = torch.randn(N, oh, ow, cout, device=dev)
    line = f"{name:6s} {cin:3d}->{cout:3d} k{k}s{s} in {h}x{w}: "
    flops = 2.0 * N * (oh * ow if not tr else h * w) * cin * cout * k * k
    for path in ("tc", "simt"):
        if path == "tc" and not ops.conv_wgrad_tc_supported(cin, cout, k, s, form, (h, w)):
            line += "   tc: n/a        "
            continue
        for _ in range(2):
            ops.conv_wgrad(x, g, form=form, ksize=k, stride=s, n_ch=cout, want_bias=False, path=path)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = 10 if path == "tc" else 3
        e0.record()
        for _ in range(reps):
            ops.conv_wgrad(x, g, form=form, ksize=k, stride=s, n_ch=cout, want_bias=False, path=path)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        tot[path] += ms
        line += f"{path:>5s} {ms:8.3f} ms {flops / ms / 1e9:7.1f} TF/s  "
    print(line)
print("total", tot)
