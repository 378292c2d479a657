"""Kernel table (torch.profiler) of one ms-ssim network-branch loss composition as the attack engine runs it:
clamp fused into the layout copy, value + gradient pyramid, layout copy with the clamp's gradient rules.  Usage: python scripts/msssim_profile.py [images]"""
import sys
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from imagecompression_adversarial_b200 import metrics, ops

n = int(sys.argv[1]) if len(sys.argv) > 1 else 32
dev = torch.device("cuda:0")
g = torch.Generator(device=dev).manual_seed(1)
x = torch.rand(n, 512, 768, 3, device=dev, generator=g)
ref = torch.rand(n, 3, 512, 768, device=dev, generator=g)
ones = torch.ones(n, device=dev)


def comp():
    out = ops.clamp01_nhwc_to_nchw(x)
    v, go = metrics.ms_ssim_value_and_grad(out, ref, ones)
    return v, ops.clamp01_backward_nchw_to_nhwc(go, x)


for _ in range(3):
    comp()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    comp()
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=60))
