"""Per-kernel parity on a B200: every C-ABI compute entry point against the same op in plain torch
(fp32, TF32 off) / the oracle layers, forward and backward.  Tolerances are written per test:
bit-exact for clamps and layout moves, 1e-5 relative for CUDA-core fp32 contractions, and a TF32
bound (inputs truncated to 10 mantissa bits, fp32 accumulate) for the tcgen05 path.
"""
import math
import os

import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from imagecompression_adversarial_b200 import ops
    ops.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def nhwc(x):
    return x.permute(0, 2, 3, 1).contiguous()


def nchw(x):
    return x.permute(0, 3, 1, 2).contiguous()


def rel_err(a, b):
    return float((a - b).abs().max() / (b.abs().max() + 1e-30))


def rms_err(a, b):
    return float((a - b).pow(2).mean().sqrt() / (b.pow(2).mean().sqrt() + 1e-30))


def ref_contraction(kind, x, w, b, k, s):
    """kind 0 Conv2d fwd, 1 Conv2d dgrad (x = grad_out), 2 ConvT fwd, 3 ConvT dgrad (x = grad_out)."""
    p = k // 2
    if kind == 0:
        return F.conv2d(x, w, b, stride=s, padding=p)
    if kind == 2:
        return F.conv_transpose2d(x, w, b, stride=s, padding=p, output_padding=s - 1)
    if kind == 1:  # gradient of conv2d wrt its input == conv_transpose2d with the same weight
        return F.conv_transpose2d(x, w, None, stride=s, padding=p, output_padding=s - 1)
    return F.conv2d(x, w, None, stride=s, padding=p)  # gradient of conv_transpose2d wrt its input


def run_contraction(kind, x, w, b, k, s, path):
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    form = L.FORM_SCONV if kind in (0, 3) else L.FORM_TCONV
    wp = ops.pack_weight(w, kind)
    out = ops.conv(nhwc(x), wp, b, form=form, ksize=k, stride=s, n_ch=wp.shape[1], path=path)
    return nchw(out)


def make_w(kind, cin_layer, cout_layer, k, dev, g):
    # torch weight of the LAYER (Conv2d [co,ci,k,k]; ConvT [ci,co,k,k]) and the channel count of `x`
    if kind in (0, 1):
        w = torch.randn(cout_layer, cin_layer, k, k, device=dev, generator=g) / math.sqrt(cin_layer * k * k)
    else:
        w = torch.randn(cin_layer, cout_layer, k, k, device=dev, generator=g) / math.sqrt(cin_layer * k * k)
    x_ch = cin_layer if kind in (0, 2) else cout_layer
    return w, x_ch


@pytest.mark.parametrize("shape", [(3, 5, 37, 41), (2, 3, 37, 41), (2, 3, 512, 768), (1, 1, 9, 300), (2, 2, 16, 16),
                                   (3, 4, 33, 65), (1, 192, 8, 12)])
def test_layout_roundtrip(dev, shape):
    """32x32-tile transpose (C > 4) and the few-channel form (RGB images)."""
    from imagecompression_adversarial_b200 import ops
    x = torch.randn(*shape, device=dev)
    y = ops.nchw_to_nhwc(x)
    assert torch.equal(y, x.permute(0, 2, 3, 1).contiguous())
    assert torch.equal(ops.nhwc_to_nchw(y), x)


@pytest.mark.parametrize("shape", [(2, 37, 41, 3), (1, 512, 768, 3), (3, 16, 16, 1)])
def test_clamp01_fused_into_the_layout_copies(dev, shape):
    """clamp01_nhwc_to_nchw / clamp01_backward_nchw_to_nhwc against the separate launches they replace in the ms-ssim
    loop (bound_forward x2 + layout copy; layout copy + bound_backward x2): bit-identical, values on and outside both
    bounds included."""
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(3)
    x = torch.rand(*shape, device=dev, generator=g) * 1.6 - 0.3
    x.view(-1)[::7] = 0.0
    x.view(-1)[3::11] = 1.0
    lo = ops.bound_forward(x.view(-1), 0.0, False)
    ref = ops.nhwc_to_nchw(ops.bound_forward(lo, 1.0, True).view_as(x))
    assert torch.equal(ops.clamp01_nhwc_to_nchw(x), ref)
    gy = torch.randn(ref.shape, device=dev, generator=g)
    gr = ops.nchw_to_nhwc(gy).view(-1)
    gr = ops.bound_backward(lo, gr, 1.0, True)
    gr = ops.bound_backward(x.view(-1), gr, 0.0, False).view_as(x)
    assert torch.equal(ops.clamp01_backward_nchw_to_nhwc(gy, x), gr)
    out = torch.empty_like(x)
    assert ops.clamp01_backward_nchw_to_nhwc(gy, x, out=out) is out and torch.equal(out, gr)


@pytest.mark.parametrize("kind", [0, 1, 2, 3])
def test_pack_unpack_weight(dev, kind):
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(kind)
    w, _ = make_w(kind, 6, 10, 5, dev, g)
    wp = ops.pack_weight(w, kind)
    back = ops.unpack_weight_grad(wp, w, kind)
    assert torch.equal(back, w)


@pytest.mark.parametrize("kind,cin,cout,k,s,hw", [
    (0, 3, 16, 5, 2, (20, 28)), (1, 3, 16, 5, 2, (10, 14)), (2, 16, 3, 5, 2, (10, 14)), (3, 16, 3, 5, 2, (20, 28)),
    (0, 8, 12, 3, 1, (9, 11)), (0, 8, 12, 3, 2, (10, 12)), (2, 128, 3, 5, 2, (6, 10)), (0, 5, 7, 1, 1, (4, 6)),
    (1, 8, 12, 3, 1, (9, 11)), (0, 16, 24, 5, 2, (17, 23)),
])
def test_conv_simt_matches_torch(dev, kind, cin, cout, k, s, hw):
    g = torch.Generator(device=dev).manual_seed(100 + kind)
    w, xc = make_w(kind, cin, cout, k, dev, g)
    x = torch.randn(2, xc, *hw, device=dev, generator=g)
    b = torch.randn(cout, device=dev, generator=g) if kind in (0, 2) else None
    if kind in (1, 3) and (s == 2 and any(d % 2 for d in hw) and kind == 3):
        pytest.skip("odd dgrad geometry")
    ref = ref_contraction(kind, x, w, b, k, s)
    got = run_contraction(kind, x, w, b, k, s, "simt")
    assert got.shape == ref.shape
    assert rel_err(got, ref) < 2e-5


TC_CASES = [
    (0, 128, 128, 5, 2, (32, 48), 2), (1, 128, 128, 5, 2, (16, 24), 2), (2, 192, 128, 5, 2, (8, 12), 1),
    (3, 192, 128, 5, 2, (16, 24), 1), (0, 128, 192, 5, 2, (32, 48), 1), (0, 128, 128, 5, 2, (20, 36), 1),
    (2, 128, 128, 5, 2, (5, 9), 2), (0, 128, 128, 3, 1, (12, 20), 1), (0, 64, 32, 1, 1, (8, 16), 1),
    (0, 192, 128, 3, 1, (9, 17), 1), (2, 128, 128, 5, 2, (16, 16), 3),
    # strided 1x1 (the skip of compressai's ResidualBlockWithStride) and its input gradient, where 3 of the 4 output
    # parity classes have no tap: those pixels are the caller's pre-fill (zero), not an epilogue on an unwritten accumulator
    (0, 128, 128, 1, 2, (16, 24), 2), (1, 128, 128, 1, 2, (8, 12), 2), (1, 128, 128, 1, 2, (40, 56), 3),
    (0, 128, 128, 3, 2, (16, 24), 1), (1, 128, 128, 3, 2, (8, 12), 2),
]


@pytest.mark.parametrize("kind,cin,cout,k,s,hw,n", TC_CASES)
def test_conv_tc_matches_torch(dev, kind, cin, cout, k, s, hw, n):
    g = torch.Generator(device=dev).manual_seed(200 + kind)
    w, xc = make_w(kind, cin, cout, k, dev, g)
    x = torch.randn(n, xc, *hw, device=dev, generator=g)
    b = torch.randn(cout, device=dev, generator=g) if kind in (0, 2) else None
    ref = ref_contraction(kind, x, w, b, k, s)
    got = run_contraction(kind, x, w, b, k, s, "tc")
    assert got.shape == ref.shape
    e = rms_err(got, ref)
    assert e < 2e-3, e           # TF32 operand truncation, fp32 accumulate
    assert rel_err(got, ref) < 1e-2
    # and against the CUDA-core path of this library
    simt = run_contraction(kind, x, w, b, k, s, "simt")
    assert rms_err(got, simt) < 2e-3


@pytest.mark.parametrize("hw,n,cout", [((32, 48), 2, 128), ((20, 36), 1, 128), ((64, 64), 1, 192)])
def test_rgb_in_conv_matches_torch(dev, hw, n, cout):
    """First-layer form: 3 -> N 5x5/2 conv through the padded RGB0 layout + overlapping-window tensor map."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(41)
    w = torch.randn(cout, 3, 5, 5, device=dev, generator=g) / math.sqrt(75)
    b = torch.randn(cout, device=dev, generator=g)
    x = torch.rand(n, 3, *hw, device=dev, generator=g)
    ref = F.conv2d(x, w, b, stride=2, padding=2)
    pad = ops.pad_rgb4(nhwc(x), ops.alloc_pad4(n, *hw, dev))
    out = ops.conv(pad, ops.pack_weight_rgb(w), b, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=cout, in_pad4=True,
                   path="tc")
    assert rms_err(nchw(out), ref) < 1.5e-3
    # same form as the input-gradient of ConvTranspose2d(N, 3): weight [N, 3, 5, 5], no bias
    ref2 = F.conv2d(x, w, None, stride=2, padding=2)
    out2 = ops.conv(pad, ops.pack_weight_rgb(w), None, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=cout, in_pad4=True,
                    path="tc")
    assert rms_err(nchw(out2), ref2) < 1.5e-3


@pytest.mark.parametrize("hw,n,cin", [((16, 24), 2, 128), ((13, 30), 1, 128), ((6, 14), 1, 192), ((7, 15), 3, 64)])
def test_col2im_deconv_matches_torch(dev, hw, n, cin):
    """Last-layer form: N -> 3 5x5/2 transposed conv as a 1x1 GEMM + col2im epilogue."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(43)
    w = torch.randn(cin, 3, 5, 5, device=dev, generator=g) / math.sqrt(cin * 25 / 4)
    b = torch.randn(3, device=dev, generator=g)
    x = torch.randn(n, cin, *hw, device=dev, generator=g)
    ref = F.conv_transpose2d(x, w, b, stride=2, padding=2, output_padding=1)
    out = ops.conv(nhwc(x), ops.pack_weight(w, L.PACK_CONVT_FWD, round_tf32=True), b, form=L.FORM_TCONV, ksize=5,
                   stride=2, n_ch=3, path="tc")
    assert nchw(out).shape == ref.shape
    assert rms_err(nchw(out), ref) < 1.5e-3
    # input-gradient of Conv2d(3, N): weight [N, 3, 5, 5]
    w2 = torch.randn(cin, 3, 5, 5, device=dev, generator=g) / math.sqrt(75)
    ref2 = F.conv_transpose2d(x, w2, None, stride=2, padding=2, output_padding=1)
    out2 = ops.conv(nhwc(x), ops.pack_weight(w2, L.PACK_CONV_DGRAD, round_tf32=True), None, form=L.FORM_TCONV,
                    ksize=5, stride=2, n_ch=3, path="tc")
    assert rms_err(nchw(out2), ref2) < 1.5e-3


def _gdn_params(C, dev, g):
    gamma = 0.1 * torch.eye(C, device=dev) + 0.02 * torch.rand(C, C, device=dev, generator=g)
    beta = 0.5 + torch.rand(C, device=dev, generator=g)
    return gamma.contiguous(), beta.contiguous()


def _gdn_ref(x, gamma, beta, inverse):
    C = x.shape[1]
    n = F.conv2d(x * x, gamma.view(C, C, 1, 1), beta)
    sc = torch.sqrt(n) if inverse else torch.rsqrt(n)
    return x * sc, sc


@pytest.mark.parametrize("inverse", [False, True])
@pytest.mark.parametrize("C,hw", [(128, (16, 32)), (192, (9, 20)), (64, (8, 16))])
def test_gdn_standalone_fwd_bwd(dev, inverse, C, hw):
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(7)
    gamma, beta = _gdn_params(C, dev, g)
    x = torch.randn(2, C, *hw, device=dev, generator=g).requires_grad_(True)
    gy = torch.randn(2, C, *hw, device=dev, generator=g)
    y_ref, sc_ref = _gdn_ref(x, gamma, beta, inverse)
    (gx_ref,) = torch.autograd.grad(y_ref, x, gy)
    xn = nhwc(x.detach())
    y, sc = ops.conv(xn, None, None, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C,
                     epi=L.EPI_IGDN_FWD if inverse else L.EPI_GDN_FWD, gmat=gamma, beta=beta, acc_from_in=True,
                     path="tc")
    assert rms_err(nchw(y), y_ref.detach()) < 1e-3
    assert rms_err(nchw(sc), sc_ref.detach()) < 1e-3
    gx = ops.conv(nhwc(gy), None, None, form=L.FORM_SCONV, ksize=1, stride=1, n_ch=C,
                  epi=L.EPI_IGDN_BWD if inverse else L.EPI_GDN_BWD, gmat=gamma.t().contiguous(), y_prev=y, sc_prev=sc,
                  acc_from_in=True, path="tc")
    assert rms_err(nchw(gx), gx_ref) < 2e-3


@pytest.mark.parametrize("inverse", [False, True])
def test_conv_gdn_fused_fwd_and_bwd(dev, inverse):
    """[conv|deconv] -> (I)GDN fused forward; next-layer dgrad -> (I)GDN backward fused."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(11)
    C = 128
    gamma, beta = _gdn_params(C, dev, g)
    if not inverse:   # g_a: conv -> GDN -> conv
        w1 = torch.randn(C, C, 5, 5, device=dev, generator=g) / math.sqrt(C * 25)
        w2 = torch.randn(C, C, 5, 5, device=dev, generator=g) / math.sqrt(C * 25)
        b1 = torch.randn(C, device=dev, generator=g)
        x = torch.randn(2, C, 32, 48, device=dev, generator=g)
        u = F.conv2d(x, w1, b1, stride=2, padding=2).requires_grad_(True)
        y_ref, _ = _gdn_ref(u, gamma, beta, False)
        z = F.conv2d(y_ref, w2, None, stride=2, padding=2)
        gz = torch.randn_like(z)
        (gu_ref,) = torch.autograd.grad(z, u, gz)
        y, sc = ops.conv(nhwc(x), ops.pack_weight(w1, 0), b1, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C,
                         epi=L.EPI_GDN_FWD, gmat=gamma, beta=beta, path="tc")
        gu = ops.conv(nhwc(gz), ops.pack_weight(w2, 1), None, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=C,
                      epi=L.EPI_GDN_BWD, gmat=gamma.t().contiguous(), y_prev=y, sc_prev=sc, path="tc")
    else:             # g_s: deconv -> IGDN -> deconv
        w1 = torch.randn(C, C, 5, 5, device=dev, generator=g) / math.sqrt(C * 25 / 4)
        w2 = torch.randn(C, C, 5, 5, device=dev, generator=g) / math.sqrt(C * 25 / 4)
        b1 = torch.randn(C, device=dev, generator=g)
        x = torch.randn(2, C, 8, 12, device=dev, generator=g)
        u = F.conv_transpose2d(x, w1, b1, stride=2, padding=2, output_padding=1).requires_grad_(True)
        y_ref, _ = _gdn_ref(u, gamma, beta, True)
        z = F.conv_transpose2d(y_ref, w2, None, stride=2, padding=2, output_padding=1)
        gz = torch.randn_like(z)
        (gu_ref,) = torch.autograd.grad(z, u, gz)
        y, sc = ops.conv(nhwc(x), ops.pack_weight(w1, 2), b1, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=C,
                         epi=L.EPI_IGDN_FWD, gmat=gamma, beta=beta, path="tc")
        gu = ops.conv(nhwc(gz), ops.pack_weight(w2, 3), None, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C,
                      epi=L.EPI_IGDN_BWD, gmat=gamma.t().contiguous(), y_prev=y, sc_prev=sc, path="tc")
    assert rms_err(nchw(y), y_ref.detach()) < 2e-3
    assert rms_err(nchw(gu), gu_ref) < 4e-3


def test_active_list_indirection(dev):
    """Image compaction: only the listed images are computed, in place of the batch."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(5)
    w = torch.randn(64, 64, 3, 3, device=dev, generator=g) / 24
    x = torch.randn(4, 64, 16, 16, device=dev, generator=g)
    ref = F.conv2d(x, w, None, padding=1)
    for path in ("tc", "simt"):
        out = torch.full((4, 16, 16, 64), -7.0, device=dev)
        active = torch.tensor([3, 1, 0, 0], device=dev, dtype=torch.int32)
        n_active = torch.tensor([2], device=dev, dtype=torch.int32)
        ops.conv(nhwc(x), ops.pack_weight(w, 0), None, form=L.FORM_SCONV, ksize=3, stride=1, n_ch=64, active=active,
                 n_active=n_active, out=out, path=path)
        o = nchw(out)
        assert rms_err(o[3], ref[3]) < 2e-3 and rms_err(o[1], ref[1]) < 2e-3
        assert torch.all(o[0] == -7.0) and torch.all(o[2] == -7.0)


@pytest.mark.parametrize("kind,cin,cout,k,s,hw", [(0, 16, 24, 5, 2, (12, 20)), (2, 24, 16, 5, 2, (6, 10)),
                                                  (0, 8, 8, 3, 1, (7, 9))])
def test_conv_wgrad_matches_autograd(dev, kind, cin, cout, k, s, hw):
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(300 + kind)
    w, xc = make_w(kind, cin, cout, k, dev, g)
    w.requires_grad_(True)
    b = torch.randn(cout, device=dev, generator=g).requires_grad_(True)
    x = torch.randn(3, xc, *hw, device=dev, generator=g)
    out = ref_contraction(kind, x, w, b, k, s)
    gout = torch.randn_like(out)
    gw_ref, gb_ref = torch.autograd.grad(out, (w, b), gout)
    form = L.FORM_SCONV if kind == 0 else L.FORM_TCONV
    dwp, db = ops.conv_wgrad(nhwc(x), nhwc(gout), form=form, ksize=k, stride=s, n_ch=cout)
    gw = ops.unpack_weight_grad(dwp, w, kind)
    assert rel_err(gw, gw_ref) < 5e-5
    assert rel_err(db, gb_ref) < 5e-5


# K6: the tcgen05 weight gradient (pixel axis as the reduction axis, MN-major operands).  Operands are pre-rounded to
# TF32 so the comparison against fp32 autograd isolates the kernel (what remains is the tensor core's fp32 accumulate).
@pytest.mark.parametrize("kind,cin,cout,k,s,hw,n", [
    (0, 128, 128, 5, 2, (32, 48), 2),     # g_a.2
    (2, 128, 128, 5, 2, (16, 24), 2),     # g_s.2
    (0, 128, 192, 5, 2, (16, 16), 3),     # g_a.6: second M block is half empty
    (2, 192, 128, 5, 2, (8, 8), 3),       # g_s.0: 192-column accumulators, two taps per CTA
    (0, 192, 128, 3, 1, (16, 16), 2),     # h_a.0
    (0, 128, 192, 3, 1, (13, 21), 1),     # ragged tiles
    (0, 64, 32, 1, 1, (9, 9), 2),
    (0, 64, 64, 3, 2, (17, 12), 2),       # cheng2020 strided 3x3
    (0, 64, 96, 1, 2, (16, 16), 2),       # cheng2020 strided 1x1 skip
    (0, 192, 384, 5, 1, (8, 12), 1),      # context model (5x5 stride 1), three M blocks
    (0, 320, 192, 5, 2, (8, 8), 1),       # two column blocks (256 + 64)
    (2, 32, 320, 5, 2, (4, 4), 1),
    (0, 128, 128, 5, 2, (128, 128), 8),   # config-5 size: many tiles per split
    (0, 3, 128, 5, 2, (64, 96), 2),       # g_a.0: RGB form (overlapping 128-byte windows of the padded RGB0 image)
    (2, 128, 3, 5, 2, (32, 48), 2),       # g_s.6: RGB form
    (0, 3, 192, 5, 2, (20, 12), 1),       # RGB form, two M blocks, ragged tiles
    (2, 192, 3, 5, 2, (5, 7), 3),
])
def test_conv_wgrad_tensor_core_matches_autograd(dev, kind, cin, cout, k, s, hw, n):
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(900 + kind + cin)
    w, xc = make_w(kind, cin, cout, k, dev, g)
    w.requires_grad_(True)
    x = ops.unary(torch.randn(n, xc, *hw, device=dev, generator=g), 5)
    out = ref_contraction(kind, x, w, None, k, s)
    gout = ops.unary(torch.randn_like(out), 5)
    (gw_ref,) = torch.autograd.grad(out, (w,), gout)
    form = L.FORM_SCONV if kind == 0 else L.FORM_TCONV
    assert ops.conv_wgrad_tc_supported(cin, cout, k, s, form, hw)
    res = {}
    for path in ("tc", "simt"):
        dwp, _ = ops.conv_wgrad(nhwc(x), nhwc(gout), form=form, ksize=k, stride=s, n_ch=cout, want_bias=False, path=path)
        res[path] = ops.unpack_weight_grad(dwp, w, kind)
    assert rel_err(res["simt"], gw_ref) < 5e-5
    assert rel_err(res["tc"], gw_ref) < 2e-4, rel_err(res["tc"], gw_ref)
    # deterministic: the split reduction has a fixed order
    dwp2, _ = ops.conv_wgrad(nhwc(x), nhwc(gout), form=form, ksize=k, stride=s, n_ch=cout, want_bias=False, path="tc")
    assert torch.equal(ops.unpack_weight_grad(dwp2, w, kind), res["tc"])


def test_bounds_match_oracle(dev):
    from imagecompression_adversarial_b200 import ops
    from oracle import layers as ol
    g = torch.Generator(device=dev).manual_seed(3)
    x = (torch.rand(10007, device=dev, generator=g) * 3 - 1)
    gy = torch.randn(10007, device=dev, generator=g)
    for bound, upper in ((0.0, False), (1.0, True), (-16 / 255, False), (16 / 255, True)):
        xr = x.clone().requires_grad_(True)
        yr = (ol.up_bound if upper else ol.low_bound)(xr, bound)
        yr.backward(gy)
        assert torch.equal(ops.bound_forward(x, bound, upper), yr.detach())
        assert torch.equal(ops.bound_backward(x, gy, bound, upper), xr.grad)


def test_perturb_step_matches_oracle_and_torch_adam(dev):
    """perturb_forward + perturb_update_adam == Up/Low_bound autograd + torch.optim.Adam + MultiStepLR."""
    from imagecompression_adversarial_b200 import ops
    from oracle import layers as ol
    g = torch.Generator(device=dev).manual_seed(9)
    N, shape = 3, (3, 16, 24)
    per = 3 * 16 * 24
    im_s = torch.rand(N, *shape, device=dev, generator=g)
    im_s[0, :, :4] = 0.0
    im_s[1, :, :4] = 1.0
    eps, budget, steps = 16 / 255, 1e-4, 12
    noise = torch.zeros(N, *shape, device=dev)
    m, v = torch.zeros_like(noise), torch.zeros_like(noise)
    im_in = torch.empty_like(noise)
    st = ops.PerturbState(N, dev)
    refs = []
    for n in range(N):
        z = torch.zeros(1, *shape, device=dev, requires_grad=True)
        opt = torch.optim.Adam([z], lr=0.01)
        sch = torch.optim.lr_scheduler.MultiStepLR(opt, [1, 2, 3], gamma=0.33)
        refs.append((z, opt, sch))
    for i in range(steps):
        gB = torch.randn(N, *shape, device=dev, generator=g) * 1e-3   # stands in for the network gradient
        ops.perturb_forward(im_s, noise, im_in, st, eps=eps, budget=budget, sched_period=steps // 3)
        branch = st.branch.cpu().tolist()
        ops.perturb_update_adam(im_s, noise, gB, m, v, st, eps=eps, gradA_scale=1.0 / per)
        for n, (z, opt, sch) in enumerate(refs):
            nc = ol.up_bound(ol.low_bound(z, -eps), eps)
            x_in = ol.up_bound(ol.low_bound(im_s[n:n + 1] + nc, 0.0), 1.0)
            loss_i = torch.mean((im_s[n:n + 1] - x_in) ** 2)
            assert abs(float(st.loss_i[n]) - float(loss_i.detach())) <= 3e-6 * float(loss_i.detach()) + 1e-12
            want_branch = 0 if float(loss_i.detach()) > budget else 1
            assert branch[n] == want_branch
            torch.testing.assert_close(im_in[n:n + 1], x_in.detach(), rtol=0, atol=1.2e-7)  # 1 ulp: FMA contraction
            loss = loss_i if want_branch == 0 else (x_in * gB[n:n + 1]).sum()
            opt.zero_grad()
            loss.backward()
            opt.step()
            if i % (steps // 3) == 0:
                sch.step()
            torch.testing.assert_close(noise[n:n + 1], z.detach(), rtol=2e-5, atol=1e-7)
    assert int(st.step[0]) == steps


def test_output_loss_matches_oracle(dev):
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    from oracle import layers as ol
    g = torch.Generator(device=dev).manual_seed(13)
    N, shape = 2, (3, 8, 12)
    per = 3 * 8 * 12
    x = (torch.rand(N, *shape, device=dev, generator=g) * 1.6 - 0.3)
    ref = torch.rand(N, *shape, device=dev, generator=g)
    gx = torch.zeros_like(x)
    ws = torch.zeros(N * L.RED_BLOCKS, device=dev)
    s = torch.zeros(N, device=dev)
    ops.output_loss(x, ref, gx, ws, s, do_clamp=True, grad_scale=1.0 / per)
    for n in range(N):
        xr = x[n:n + 1].clone().requires_grad_(True)
        o = ol.up_bound(ol.low_bound(xr, 0.0), 1.0)
        loss = 1.0 - torch.mean((ref[n:n + 1] - o) * (ref[n:n + 1] - o))
        loss.backward()
        torch.testing.assert_close(gx[n:n + 1], xr.grad, rtol=1e-6, atol=1e-9)
        assert abs(float(s[n]) / per - float(1.0 - loss)) < 1e-6


def test_tc_backward_epilogue_is_deterministic(dev):
    """Regression: the saved-y/scale staging is overwritten by TMA while other threads may still be reading it
    unless a proxy fence + barrier separates them; outputs must be bit-identical run to run."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(77)
    C, n, H, W = 128, 4, 64, 96
    gm = (0.1 * torch.eye(C, device=dev) + 0.01 * torch.rand(C, C, device=dev, generator=g)).contiguous()
    x = torch.randn(n, H, W, C, device=dev, generator=g)
    w = torch.randn(25, C, C, device=dev, generator=g) / 56
    for form in (L.FORM_SCONV, L.FORM_TCONV):
        oh, ow = ops.out_hw(form, 5, 2, H, W)
        y = torch.randn(n, oh, ow, C, device=dev, generator=g)
        sc = 0.5 + torch.rand(n, oh, ow, C, device=dev, generator=g)
        for epi in (L.EPI_GDN_BWD, L.EPI_IGDN_BWD):
            outs = [ops.conv(x, w, None, form=form, ksize=5, stride=2, n_ch=C, epi=epi, gmat=gm, y_prev=y, sc_prev=sc,
                             path="tc") for _ in range(5)]
            for o in outs[1:]:
                assert torch.equal(o, outs[0])


@pytest.fixture
def persist_env():
    import os
    old = os.environ.get("ICADV_TC_PERSIST")
    yield lambda v: os.environ.__setitem__("ICADV_TC_PERSIST", str(v))
    if old is None:
        os.environ.pop("ICADV_TC_PERSIST", None)
    else:
        os.environ["ICADV_TC_PERSIST"] = old


@pytest.mark.parametrize("case", ["sconv_gdn_fwd", "tconv_igdn_fwd", "tconv_gdn_bwd", "sconv_igdn_bwd", "linear_relu",
                                  "rgb_in_gdn", "rgb_in_igdn_bwd", "col2im"])
def test_persistent_variant_equals_per_tile_kernel(dev, persist_env, case):
    """The persistent kernel (one CTA per SM, TMEM double buffer, two epilogue warpgroups, all parity classes in one
    launch; ICADV_TC_PERSIST=2 forces it for every eligible shape) and the per-tile kernel (=0) run the same K order and
    the same epilogue arithmetic: results must agree to fp32 round-off, on ragged sizes and with an active-image list."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(77)
    C, n = 128, 5
    gamma, beta = _gdn_params(C, dev, g)
    act_idx = torch.tensor([4, 0, 2, 0, 0], device=dev, dtype=torch.int32)
    n_act = torch.tensor([3], device=dev, dtype=torch.int32)
    outs = []
    for level in (0, 2):
        persist_env(level)
        kw = dict(active=act_idx, n_active=n_act)
        if case == "sconv_gdn_fwd":
            x = torch.randn(n, 37, 50, C, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            w = ops.pack_weight(torch.randn(C, C, 5, 5, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 56, 0)
            out = torch.zeros(n, 19, 25, C, device=dev); sc = torch.zeros_like(out)
            ops.conv(x, w, beta, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_GDN_FWD, gmat=gamma, beta=beta,
                     out=out, out_scale=sc, path="tc", round_out=True, **kw)
            outs.append((out, sc))
        elif case == "tconv_igdn_fwd":
            x = torch.randn(n, 11, 19, 192, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            w = ops.pack_weight(torch.randn(192, C, 5, 5, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 35, 2)
            out = torch.zeros(n, 22, 38, C, device=dev); sc = torch.zeros_like(out)
            ops.conv(x, w, beta, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_IGDN_FWD, gmat=gamma, beta=beta,
                     out=out, out_scale=sc, path="tc", **kw)
            outs.append((out, sc))
        elif case in ("tconv_gdn_bwd", "sconv_igdn_bwd"):
            tconv = case == "tconv_gdn_bwd"
            gsrc = torch.randn(n, 13, 21, C, device=dev, generator=torch.Generator(device=dev).manual_seed(1)) if tconv else \
                torch.randn(n, 26, 42, C, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            oh, ow = (26, 42) if tconv else (13, 21)
            w = ops.pack_weight(torch.randn(C, C, 5, 5, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 56,
                                1 if tconv else 3)
            yp = torch.randn(n, oh, ow, C, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
            sp = 0.5 + torch.rand(n, oh, ow, C, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
            out = torch.zeros(n, oh, ow, C, device=dev)
            ops.conv(gsrc, w, None, form=L.FORM_TCONV if tconv else L.FORM_SCONV, ksize=5, stride=2, n_ch=C,
                     epi=L.EPI_GDN_BWD if tconv else L.EPI_IGDN_BWD, gmat=gamma.t().contiguous(), y_prev=yp, sc_prev=sp,
                     out=out, path="tc", round_out=True, **kw)
            outs.append((out,))
        elif case == "linear_relu":
            x = torch.randn(n, 12, 20, 192, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            w = ops.pack_weight(torch.randn(C, 192, 3, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 40, 0)
            out = torch.zeros(n, 12, 20, C, device=dev)
            ops.conv(x, w, beta, form=L.FORM_SCONV, ksize=3, stride=1, n_ch=C, act=L.ACT_RELU, out=out, path="tc", **kw)
            outs.append((out,))
        elif case == "rgb_in_igdn_bwd":
            # input gradient of the last synthesis layer (deconv N -> 3) with the IGDN backward in its epilogue: the
            # HBM-bound launch the streaming backward kernel exists for
            gx = torch.randn(n, 36, 52, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            pad = ops.pad_rgb4(gx, ops.alloc_pad4(n, 36, 52, dev))
            w = ops.pack_weight_rgb(torch.randn(C, 3, 5, 5, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 9)
            yp = torch.randn(n, 18, 26, C, device=dev, generator=torch.Generator(device=dev).manual_seed(3))
            sp = 0.5 + torch.rand(n, 18, 26, C, device=dev, generator=torch.Generator(device=dev).manual_seed(4))
            out = torch.zeros(n, 18, 26, C, device=dev)
            ops.conv(pad, w, None, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_IGDN_BWD,
                     gmat=gamma.t().contiguous(), y_prev=yp, sc_prev=sp, out=out, path="tc", in_pad4=True, round_out=True,
                     **kw)
            outs.append((out,))
        elif case == "rgb_in_gdn":
            x = torch.rand(n, 36, 52, 3, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            pad = ops.pad_rgb4(x, ops.alloc_pad4(n, 36, 52, dev))
            w = ops.pack_weight_rgb(torch.randn(C, 3, 5, 5, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 9)
            out = torch.zeros(n, 18, 26, C, device=dev); sc = torch.zeros_like(out)
            ops.conv(pad, w, beta, form=L.FORM_SCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_GDN_FWD, gmat=gamma, beta=beta,
                     out=out, out_scale=sc, path="tc", in_pad4=True, **kw)
            outs.append((out, sc))
        else:
            x = torch.randn(n, 15, 23, C, device=dev, generator=torch.Generator(device=dev).manual_seed(1))
            w = ops.pack_weight(torch.randn(C, 3, 5, 5, device=dev, generator=torch.Generator(device=dev).manual_seed(2)) / 30, 2)
            b3 = torch.randn(3, device=dev, generator=torch.Generator(device=dev).manual_seed(5))
            out = torch.zeros(n, 30, 46, 3, device=dev)
            ops.conv(x, w, b3, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=3, out=out, path="tc", **kw)
            outs.append((out,))
    for a, b in zip(*outs):
        assert float(a.abs().max()) > 0
        assert float(a[1].abs().max()) == 0 and float(b[1].abs().max()) == 0     # images not in the active list stay untouched
        # The persistent kernel's backward epilogue accumulates the normalisation chunks of its two warpgroups
        # alternately (0, 2, 1, 3), the per-tile kernel in order: fp32 sums differ in the last bit, and where the output
        # is rounded to TF32 (round_out) that can flip one TF32 ulp (2^-10 relative) on a few elements.
        exact = torch.isclose(b, a, rtol=1e-5, atol=1e-6)
        assert float((~exact).float().mean()) < 1e-3, float((~exact).float().mean())
        torch.testing.assert_close(b, a, rtol=1.2e-3, atol=1e-6)


def test_streaming_backward_kernel_matches_the_in_place_persistent_kernel(dev):
    """ICADV_TC_STREAM_BWD=0 (saved chunks fetched per epilogue group, re-read in pass 2) vs =1 (saved-tensor ring, pass-1
    products stashed in TMEM, plain stores): same chunk order and the same normalisation operand; the last step differs in
    one rounding (g*sc is rounded to fp32 when it is stashed, the other kernel fuses it into the final multiply-add), so
    results agree to fp32 round-off, with a TF32 ulp on the few elements where that crosses the output rounding.  The
    streaming kernel itself is bit-reproducible.  ~17 items per CTA so every ring and the TMEM buffers wrap many times,
    ragged tile edges included."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    C, n, h, w = 128, 6, 75, 101
    gen = lambda seed: torch.Generator(device=dev).manual_seed(seed)
    gamma, _ = _gdn_params(C, dev, gen(3))
    x = torch.randn(n, h, w, C, device=dev, generator=gen(1))
    wp = ops.pack_weight(torch.randn(C, C, 5, 5, device=dev, generator=gen(2)) / 56, 1)
    yp = torch.randn(n, 2 * h, 2 * w, C, device=dev, generator=gen(4))
    sp = 0.5 + torch.rand(n, 2 * h, 2 * w, C, device=dev, generator=gen(5))
    old = os.environ.get("ICADV_TC_STREAM_BWD")
    res = []
    try:
        for level in ("0", "1", "1"):
            os.environ["ICADV_TC_STREAM_BWD"] = level
            out = torch.zeros(n, 2 * h, 2 * w, C, device=dev)
            ops.conv(x, wp, None, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_GDN_BWD,
                     gmat=gamma.t().contiguous(), y_prev=yp, sc_prev=sp, out=out, path="tc", round_out=True)
            res.append(out)
    finally:
        if old is None:
            os.environ.pop("ICADV_TC_STREAM_BWD", None)
        else:
            os.environ["ICADV_TC_STREAM_BWD"] = old
    assert float(res[0].abs().max()) > 0
    assert torch.equal(res[1], res[2])
    exact = torch.isclose(res[1], res[0], rtol=1e-5, atol=1e-6)
    assert float((~exact).float().mean()) < 1e-3, float((~exact).float().mean())
    torch.testing.assert_close(res[1], res[0], rtol=1.2e-3, atol=1e-6)


@pytest.mark.parametrize("epi", ["igdn_fwd", "gdn_bwd", "col2im"])
def test_persistent_kernel_many_items_per_cta(dev, persist_env, epi):
    """Transposed convs at a size where every persistent CTA walks ~10-20 work items (ring phases, TMEM buffer hand-offs,
    in-place saved-chunk staging all wrap many times): persistent vs per-tile kernel, and run-to-run bit-identity."""
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    C, n, h, w = 128, 6, 72, 104
    gen = lambda seed: torch.Generator(device=dev).manual_seed(seed)
    gamma, beta = _gdn_params(C, dev, gen(3))
    x = torch.randn(n, h, w, C, device=dev, generator=gen(1))
    res = {}
    for level in (0, 2, 2):
        persist_env(level)
        if epi == "col2im":
            wp = ops.pack_weight(torch.randn(C, 3, 5, 5, device=dev, generator=gen(2)) / 30, 2)
            out = torch.zeros(n, 2 * h, 2 * w, 3, device=dev)
            ops.conv(x, wp, torch.zeros(3, device=dev), form=L.FORM_TCONV, ksize=5, stride=2, n_ch=3, out=out, path="tc")
        elif epi == "igdn_fwd":
            wp = ops.pack_weight(torch.randn(C, C, 5, 5, device=dev, generator=gen(2)) / 28, 2)
            out = torch.zeros(n, 2 * h, 2 * w, C, device=dev); sc = torch.zeros_like(out)
            ops.conv(x, wp, beta, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_IGDN_FWD, gmat=gamma, beta=beta,
                     out=out, out_scale=sc, path="tc")
            out = torch.cat((out, sc))
        else:
            wp = ops.pack_weight(torch.randn(C, C, 5, 5, device=dev, generator=gen(2)) / 56, 1)
            yp = torch.randn(n, 2 * h, 2 * w, C, device=dev, generator=gen(4))
            sp = 0.5 + torch.rand(n, 2 * h, 2 * w, C, device=dev, generator=gen(5))
            out = torch.zeros(n, 2 * h, 2 * w, C, device=dev)
            ops.conv(x, wp, None, form=L.FORM_TCONV, ksize=5, stride=2, n_ch=C, epi=L.EPI_GDN_BWD,
                     gmat=gamma.t().contiguous(), y_prev=yp, sc_prev=sp, out=out, path="tc")
        res.setdefault(level, []).append(out)
    torch.testing.assert_close(res[2][0], res[0][0], rtol=1e-5, atol=1e-6)
    assert torch.equal(res[2][0], res[2][1])


def test_round_ste_and_universal_quantiser(dev):
    """utils/ops.py:8-25 on the operator surface: Round_STE = torch.round forward / identity backward; UniverseQuant =
    round(x + u) - u with u ~ U(-1/2, 1/2): error in [-1/2, 1/2], unbiased, identity backward, reproducible from the seed."""
    from imagecompression_adversarial_b200 import functional as Fn
    g = torch.Generator(device=dev).manual_seed(5)
    x = (torch.randn(3, 7, 33, 20, device=dev, generator=g) * 4).requires_grad_(True)
    y = Fn.Round_STE.apply(x)
    assert torch.equal(y.detach(), torch.round(x.detach()))
    go = torch.randn_like(y)
    y.backward(go)
    assert torch.equal(x.grad, go)
    x.grad = None
    torch.manual_seed(11)
    q1 = Fn.UniverseQuant.apply(x)
    torch.manual_seed(11)
    q2 = Fn.UniverseQuant.apply(x)
    assert torch.equal(q1, q2)
    err = (q1 - x).detach()
    assert float(err.abs().max()) <= 0.5 + 1e-6 and abs(float(err.mean())) < 0.02
    assert 0.07 < float(err.var()) < 0.10                    # U(-1/2, 1/2): variance 1/12
    q1.backward(go)
    assert torch.equal(x.grad, go)
