"""Developer probe for the MN-major operand descriptors of the tcgen05 weight gradient: one-hot operands show which
(pixel, channel) elements the tensor core pairs up."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from imagecompression_adversarial_b200 import _lib as L, ops
dev = torch.device("cuda:0")
C = 32
for variant in (0, 1, 8):
    os.environ["ICADV_WG_VARIANT"] = str(variant)
    print("=== variant", variant)
    for (px0, n0) in ((0, 0), (5, 3), (9, 17), (63, 31)):
        g = torch.zeros(1, 8, 8, C, device=dev)
        g.view(-1, C)[px0, n0] = 1.0
        x = (torch.arange(64, device=dev).view(64, 1) * 100 + torch.arange(C, device=dev).view(1, C)).float().view(1, 8, 8, C).contiguous()
        dw, _ = ops.conv_wgrad(x, g, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C, want_bias=False, path="tc")
        torch.cuda.synchronize()
        dw = dw.view(C, C)
        nz = dw.nonzero()
        print(f"px0={px0} n0={n0}: nonzeros {nz.shape[0]}, rows {sorted(set(nz[:,0].tolist()))[:8]}, "
              f"row n0 -> {dw[n0,:6].tolist()} (want {[px0*100+k for k in range(6)]})")
    # random full check
    g = torch.randn(2, 16, 16, 64, device=dev); x = torch.randn(2, 16, 16, 64, device=dev)
    a, _ = ops.conv_wgrad(x, g, form=L.FORM_SCONV, ksize=3, stride=1, n_ch=64, want_bias=False, path="tc")
    b, _ = ops.conv_wgrad(x, g, form=L.FORM_SCONV, ksize=3, stride=1, n_ch=64, want_bias=False, path="simt")
    print("3x3 random: rel err", float((a - b).abs().max() / b.abs().max()), "nonzero frac", float((a != 0).float().mean()))
