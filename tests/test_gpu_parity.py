"""Parity at the north star's LITERAL tolerances, on the BASELINE.json configurations at their stated sizes, in the
3xTF32 parity mode of the tensor path (precision.py, csrc/icadv_split.cu: every product is three tcgen05 kind::tf32
MMAs on hi/lo operand splits, fp32 accumulation in TMEM).

Tolerances asserted here (BASELINE.json north_star): per-step loss AND loss_i within 1e-3 relative, bare-MSE losses
within 1e-3 relative, final PSNR within 0.05 dB, bpp within 1e-3 * max(1, bpp), quantised latent indices equal
outside a 1e-4 guard band.  The oracle (plain torch fp32, cuDNN with TF32 off) runs on the same GPU.
The speed mode ("tf32": one MMA per product, the arithmetic of the reference's own cuDNN-TF32 GPU path) is covered by
tests/test_gpu_e2e.py and tests/test_gpu_fullsize.py with its stated, wider bounds.
"""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu

REL = 1e-3          # per-step loss / loss_i / bare MSE (north star)
PSNR_DB = 0.05
TIE = 1e-3          # a branch flip is legitimate only if loss_i is within this relative distance of the budget


def bpp_tol(b):
    return 1e-3 * max(1.0, abs(b))


@pytest.fixture(scope="module")
def dev():
    from imagecompression_adversarial_b200 import ops, precision
    ops.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    # the ORACLE must be run-to-run reproducible too: with cuDNN free to pick atomics-based backward algorithms its own
    # 100-step trajectories differ between runs at the 1e-3 level this file asserts (seen: loss_i of step 67 of the
    # config-1 forced-branch test off by 1.018e-3 in one run of six, the library's side being bit-reproducible)
    prev_det, prev_bench = torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = True, False
    prev = precision.get()
    precision.set("3xtf32")
    yield torch.device("cuda:0")
    precision.set(prev)
    torch.backends.cudnn.deterministic, torch.backends.cudnn.benchmark = prev_det, prev_bench


def pair(model, quality, dev, seed=0):
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    onet = om.init_model(model, quality, seed=seed).to(dev)
    if model == "cheng2020_attn":
        # kaiming-initialised attention blocks (three un-normalised residual units per branch, four blocks per stack)
        # amplify a random-init network's activations to 1e18-1e22 in BOTH implementations (fp32 overflow in the PSNR);
        # 0.6 x the g_a / g_s conv weights keeps them O(1), like trained weights
        with torch.no_grad():
            for name, m in onet.named_modules():
                if isinstance(m, torch.nn.Conv2d) and name.startswith(("g_a", "g_s")):
                    m.weight.mul_(0.6)
    pnet = pm.init_model(model, quality, "mse", pretrained=False).to(dev)
    pnet.load_state_dict(onet.state_dict(), strict=True)
    return onet, pnet


def images(n, h, w, dev, first=0):
    from oracle.attack import synthetic_image
    return torch.cat([synthetic_image(first + i, h, w) for i in range(n)]).to(dev)


def psnr(a, b):
    return -10.0 * math.log10(float(torch.mean((a - b) ** 2)) + 1e-30)


def relerr(a, b):
    return float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp(min=1e-30))


# ------------------------------------------------------------------------------------------ the mode itself
def test_split_kernels(dev):
    from imagecompression_adversarial_b200 import _lib as L
    from imagecompression_adversarial_b200 import ops
    g = torch.Generator(device=dev).manual_seed(1)
    for c, geom in ((3, ops.split_geom(3, L.FORM_SCONV, 5, 2)), (128, ops.split_geom(128, L.FORM_SCONV, 5, 2)),
                    (128, ops.split_geom(128, L.FORM_TCONV, 5, 2)), (192, ops.split_geom(192))):
        ks, G = geom
        assert ks % 32 == 0 and ks * G >= 3 * c
        x = torch.randn(5, 7, c, device=dev, generator=g) * 3
        for layout, fn in ((0, ops.split3), (1, ops.split3_weight)):
            s = fn(x, ks, G)
            assert s.shape == (G, 5, 7, ks)
            full = s.permute(1, 2, 0, 3).reshape(5, 7, G * ks)          # slices side by side = the whole split form
            a, b, d = full[..., :c], full[..., c:2 * c], full[..., 2 * c:3 * c]
            hi, lo = (a, b) if layout == 0 else (a, d)
            assert torch.equal(a, d if layout == 0 else b)
            assert float(full[..., 3 * c:].abs().sum()) == 0.0
            # hi and lo are TF32 numbers (13 low mantissa bits clear); hi + lo reproduces x to 2^-21
            assert int((hi.contiguous().view(torch.int32) & 0x1FFF).abs().sum()) == 0
            assert int((lo.contiguous().view(torch.int32) & 0x1FFF).abs().sum()) == 0
            assert float(((hi + lo) - x).abs().max() / x.abs().max()) < 2.0 ** -21
        sq = ops.split3(x, ks, G, op=1).permute(1, 2, 0, 3).reshape(5, 7, G * ks)
        assert float(((sq[..., :c] + sq[..., c:2 * c]) - x * x).abs().max() / (x * x).abs().max()) < 2.0 ** -21
    parts = torch.randn(4, 1000, device=dev, generator=g)
    out = ops.sum_slices(parts, torch.empty(1000, device=dev))
    assert torch.equal(out, ((parts[0] + parts[1]) + parts[2]) + parts[3])


def test_philox_noise_statistics_and_reproducibility(dev):
    from imagecompression_adversarial_b200 import ops
    torch.manual_seed(11)
    x = torch.empty(2, 192, 32, 48, device=dev)
    a = ops.uniform_noise_like(x)
    b = ops.uniform_noise_like(x)
    torch.manual_seed(11)
    a2 = ops.uniform_noise_like(x)
    assert torch.equal(a, a2) and not torch.equal(a, b)
    assert float(a.min()) >= -0.5 and float(a.max()) < 0.5
    assert abs(float(a.mean())) < 2e-3 and abs(float(a.var()) - 1.0 / 12) < 2e-3
    # no visible correlation between neighbours / between the two draws
    f = a.flatten()
    assert abs(float((f[1:] * f[:-1]).mean())) < 1e-3 and abs(float((a * b).mean())) < 1e-3


@pytest.mark.parametrize("kind,cin,cout,k,s,hw", [("conv", 128, 128, 5, 2, (64, 96)), ("deconv", 128, 128, 5, 2, (32, 48)),
                                                  ("conv", 3, 128, 5, 2, (64, 96)), ("deconv", 128, 3, 5, 2, (32, 48)),
                                                  ("conv", 192, 320, 3, 1, (16, 24)), ("deconv", 320, 192, 5, 2, (16, 24))])
def test_split_contraction_is_fp32_accurate(dev, kind, cin, cout, k, s, hw):
    """One contraction through the autograd surface in the parity mode against fp64 torch: forward and input gradient
    at fp32 accuracy (the single-pass TF32 mode sits at ~1e-3 here)."""
    from imagecompression_adversarial_b200 import models as pm
    torch.manual_seed(2)
    m = (pm.Conv2d if kind == "conv" else pm.ConvTranspose2d)(cin, cout, k, s).to(dev)
    with torch.no_grad():
        m.bias.uniform_(-0.1, 0.1)
    x = torch.randn(2, cin, *hw, device=dev, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xd = x.detach().double().requires_grad_(True)
    if kind == "conv":
        yd = torch.nn.functional.conv2d(xd, m.weight.double(), m.bias.double(), stride=s, padding=k // 2)
    else:
        yd = torch.nn.functional.conv_transpose2d(xd, m.weight.double(), m.bias.double(), stride=s, padding=k // 2,
                                                  output_padding=s - 1)
    yd.backward(gy.double())
    assert relerr(y.detach().double(), yd.detach()) < 3e-6, relerr(y.detach().double(), yd.detach())
    assert relerr(x.grad.double(), xd.grad) < 3e-6, relerr(x.grad.double(), xd.grad)


@pytest.mark.parametrize("inverse", [False, True])
def test_split_gdn_is_fp32_accurate(dev, inverse):
    from imagecompression_adversarial_b200 import models as pm
    from oracle import layers as ol
    torch.manual_seed(4)
    C = 128
    og, pg = ol.GDN(C, inverse=inverse).to(dev), pm.GDN(C, inverse=inverse).to(dev)
    with torch.no_grad():
        og.gamma.add_(0.02 * torch.rand_like(og.gamma))
    pg.load_state_dict(og.state_dict())
    x = torch.randn(2, C, 24, 40, device=dev)
    outs = []
    for mod, dt in ((og.double(), torch.float64), (pg, torch.float32)):
        xi = x.to(dt).requires_grad_(True)
        y = mod(xi)
        y.backward(torch.ones_like(y) * 0.5 + y.detach() * 0.1)
        outs.append((y.detach().double(), xi.grad.double()))
    assert relerr(outs[1][0], outs[0][0]) < 3e-6 and relerr(outs[1][1], outs[0][1]) < 5e-6


@pytest.mark.parametrize("model,quality", [("hyper", 3), ("factorized", 1)])
def test_split_stack_program_matches_oracle(dev, model, quality):
    """g_s(g_a(x)) and its input gradient through the fused-program surface (attack_rd.py:344-349,547)."""
    onet, pnet = pair(model, quality, dev)
    onet.train(); pnet.train()
    x = images(1, 128, 192, dev)
    ref = torch.rand_like(x)
    outs = []
    for net in (onet, pnet):
        xi = x.clone().requires_grad_(True)
        out = net.g_s(net.g_a(xi))
        loss = torch.mean((ref - out) * (ref - out))
        loss.backward()
        outs.append((out.detach(), xi.grad.detach(), float(loss)))
    (oo, og, ol_), (po, pg, pl) = outs
    assert relerr(po, oo) < 2e-5, relerr(po, oo)
    assert relerr(pg, og) < 1e-4, relerr(pg, og)
    assert abs(pl - ol_) < 1e-4 * abs(ol_)      # raw MSE against a random target (the speed mode sits at 3e-3 here)


@pytest.mark.parametrize("model,quality,hw", [("factorized", 1, (128, 192)), ("hyper", 3, (128, 192)),
                                              ("context", 4, (128, 192)), ("cheng2020", 6, (128, 128)),
                                              ("cheng2020_attn", 1, (128, 128)), ("debug", 1, (128, 128))])
def test_eval_forward_parity(dev, model, quality, hw):
    """net(x) in eval mode: latent indices, bpp, PSNR at the literal tolerances for all four families (+ the attention
    variant of cheng2020)."""
    onet, pnet = pair(model, quality, dev)
    x = images(2, *hw, dev)
    onet.eval(); pnet.eval()
    with torch.no_grad():
        o, p = onet(x), pnet(x)
        yo, yp = onet.g_a(x), pnet.g_a(x)
    assert relerr(yp, yo) < 5e-5, relerr(yp, yo)
    frac = (yo - torch.floor(yo) - 0.5).abs()
    safe = frac > 1e-4 * yo.abs().clamp(min=1.0)
    assert bool((torch.round(yo)[safe] == torch.round(yp)[safe]).all())
    num_px = x.shape[0] * hw[0] * hw[1]
    bpp_o = sum(float(torch.log(l).sum()) for l in o["likelihoods"].values()) / (-math.log(2) * num_px)
    bpp_p = sum(float(torch.log(l).sum()) for l in p["likelihoods"].values()) / (-math.log(2) * num_px)
    assert abs(bpp_o - bpp_p) < bpp_tol(bpp_o), (bpp_o, bpp_p)
    assert abs(psnr(p["x_hat"], x) - psnr(o["x_hat"], x)) < PSNR_DB


# ------------------------------------------------------------------------------------------ trajectories
def compare_trajectory(rec, orec, i, budget, roi=False):
    """Per-step (branch, loss_i, loss) of image i against the oracle's record.  Returns the index of the first branch
    flip (which must be a tie at the budget boundary) or None; every step before it is held to REL."""
    for t, (br, loss, loss_i) in enumerate(orec):
        pb, pli, pl = int(rec[t][0][i]), float(rec[t][1][i]), float(rec[t][2][i])
        if (pb == 1) != (br == "B"):
            assert abs(loss_i - budget) < TIE * budget, ("branch flip away from the budget boundary", t, loss_i, pli)
            return t
        assert abs(pli - loss_i) <= REL * max(loss_i, 1e-12), ("loss_i", t, pli, loss_i)
        assert abs(pl - loss) <= REL * abs(loss), ("loss", t, pl, loss, br)
    return None


def final_metrics_agree(p, o, x, pnet=None, args=None):
    """End of an attack whose branch sequence matched: the PSNR of the adversarial image agrees within 0.05 dB, the
    clean-pass rate agrees, and the final evaluation (self_ensemble.py:173-252) agrees: reconstruction PSNR within
    0.05 dB, bpp within 1e-3 * max(1, bpp).
    (The perturbations themselves are NOT compared element-wise: Adam moves every pixel by ~lr * sign(g), and wherever
    the input gradient is at fp32 noise level -- most pixels under a saturated random-init reconstruction -- the sign
    is arbitrary in the oracle as well; measured: the two perturbations differ by 10 % rms after 100 forced steps while
    every per-step loss agrees to 1e-3.)"""
    im_adv, out_adv, out_s, bpp_ori, bpp = p[0], p[1], p[2], p[3], p[4]
    assert abs(psnr(im_adv, x) - psnr(o[0], x)) < PSNR_DB
    assert abs(float(bpp_ori) - float(o[3])) < bpp_tol(float(o[3])), (float(bpp_ori), float(o[3]))
    if pnet is not None:
        # evaluation of the SAME adversarial image by both implementations: isolates eval parity from the (chaotic)
        # sensitivity of a random-init codec's saturated reconstruction to 1e-6-level differences of its input
        from imagecompression_adversarial_b200 import attack as patk
        _, out_adv, bpp, _, _ = patk.eval(o[0].contiguous(), x, o[2], pnet, args)
    assert abs(psnr(out_adv, o[2]) - psnr(o[1], o[2])) < PSNR_DB, (psnr(out_adv, o[2]), psnr(o[1], o[2]))
    assert abs(float(bpp) - float(o[4])) < bpp_tol(float(o[4])), (float(bpp), float(o[4]))


@pytest.mark.parametrize("force", [-1, 1])
def test_config1_factorized_q1_256_100steps(dev, force):
    """BASELINE configs[0] exactly: Balle2016 factorized q=1, MSE distortion attack, one synthetic 256x256 image,
    100 steps, random-init weights -- with the natural branch mix and with the network branch forced on every step."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = pair("factorized", 1, dev)
    x = images(1, 256, 256, dev)
    args = oatk.default_args(model="factorized", quality=1, metric="mse", steps=100, force_branch=force)
    rec, orec = [], []
    p = patk.attack_(x, pnet, args, record=rec)
    o = oatk.attack_(x, onet, args, record=orec)
    assert len(rec) == len(orec) == 100
    flip = compare_trajectory(rec, orec, 0, args.noise)
    if force == 1:
        assert flip is None and all(br == "B" for br, _, _ in orec)
    if flip is None:
        final_metrics_agree(p, o, x, pnet, args)
    else:
        assert flip >= 3, flip      # the trajectory up to the tie was compared step by step above


def test_config2_hyper_q3_fullsize_60_forced_steps(dev):
    """BASELINE configs[1] (the benchmarked step) on one 768x512 image: 60 consecutive network-branch iterations,
    every step's loss and loss_i within 1e-3 of the fp32 oracle; then the public call with the natural mix."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = pair("hyper", 3, dev)
    x = images(1, 512, 768, dev)
    args = oatk.default_args(model="hyper", quality=3, metric="mse", steps=60, force_branch=1)
    rec, orec = [], []
    p = patk.attack_(x, pnet, args, record=rec)
    o = oatk.attack_(x, onet, args, record=orec)
    assert compare_trajectory(rec, orec, 0, args.noise) is None
    final_metrics_agree(p, o, x, pnet, args)
    args = oatk.default_args(model="hyper", quality=3, metric="mse", steps=30)
    rec, orec = [], []
    p = patk.attack_(x, pnet, args, record=rec)
    o = oatk.attack_(x, onet, args, record=orec)
    flip = compare_trajectory(rec, orec, 0, args.noise)
    if flip is None:
        final_metrics_agree(p, o, x, pnet, args)


def test_config2_batch_of_fullsize_images_each_matches_its_own_oracle_run(dev):
    """Per-image semantics of the batched loop: image i of a 3-image 768x512 batch follows the oracle run on image i."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = pair("hyper", 3, dev)
    x = images(3, 512, 768, dev, first=4)
    args = oatk.default_args(model="hyper", quality=3, metric="mse", steps=12)
    rec = []
    patk.attack_(x, pnet, args, record=rec)
    for i in range(3):
        orec = []
        oatk.attack_(x[i:i + 1], onet, args, record=orec)
        compare_trajectory(rec, orec, i, args.noise)


def test_config3_context_q4_msssim_fullsize(dev):
    """BASELINE configs[2]: Minnen2018 context model q=4, -att_metric ms-ssim, 768x512."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = pair("context", 4, dev)
    x = images(1, 512, 768, dev)
    args = oatk.default_args(model="context", quality=4, metric="ms-ssim", steps=9, att_metric="ms-ssim", noise=2e-5)
    rec, orec = [], []
    p = patk.attack_(x, pnet, args, record=rec)
    o = oatk.attack_(x, onet, args, record=orec)
    assert {br for br, _, _ in orec} == {"A", "B"}
    flip = compare_trajectory(rec, orec, 0, args.noise)
    if flip is None:
        final_metrics_agree(p, o, x, pnet, args)


def test_config4_cheng2020_q6_targeted_roi_fullsize(dev):
    """BASELINE configs[3]: Cheng2020 (anchor) q=6, targeted attack (-t) with an ROI mask (--mask_loc), 768x512.
    Semantics of the targeted / ROI loss: oracle.attack.attack_our_roi (the reference's own path is dead code)."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = pair("cheng2020", 6, dev)
    x = images(1, 512, 768, dev)
    t = images(1, 512, 768, dev, first=7)
    args = oatk.default_args(model="cheng2020", quality=6, metric="mse", steps=6, noise=3e-5,
                             mask_loc=[192, 576, 128, 384], lamb_bkg_in=0.5, lamb_bkg_out=2.0, lamb_tar=1.5)
    rec, orec = [], []
    p = patk.attack_(x, pnet, args, record=rec, im_t=t)
    o = oatk.attack_(x, onet, args, record=orec, im_t=t)
    assert "B" in {br for br, _, _ in orec}
    flip = compare_trajectory(rec, orec, 0, args.noise, roi=True)
    if flip is None:
        final_metrics_agree(p, o, x, pnet, args)


def test_ifgsm_parity(dev):
    """Sign update (attack_ifgsm.py:364-438): the bare-MSE loss of every step within 1e-3."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = pair("hyper", 3, dev)
    x = images(1, 192, 256, dev)
    args = oatk.default_args(model="hyper", quality=3, metric="mse", steps=8)
    rec, orec = [], []
    p = patk.attack_ifgsm(x, pnet, args, record=rec)
    o = oatk.attack_ifgsm(x, onet, args, record=orec)
    for t in range(8):
        assert abs(float(rec[t][0]) - orec[t][1]) <= REL * abs(orec[t][1]), (t, float(rec[t][0]), orec[t][1])
    assert abs(psnr(p[0], x) - psnr(o[0], x)) < PSNR_DB


@pytest.mark.parametrize("n_img", [1, 2])
def test_config5_adv_train_step_300_attack_steps(dev, n_img):
    """BASELINE configs[4] at world size 1: one train.py --adv iteration (train.py:335-366) = attack_ with -steps 300
    on the batch, then the codec update; hyperprior q1, 256x256 crops.  The reference tests its budget on the BATCH mean
    (attack_rd.py:333-334) and so does adv_train_step (budget_scope="batch"); this 300-step comparison runs one image and a
    batch of two per step, the batch semantics themselves (shared branch, 1/B gradient scale) are also pinned step by step by
    test_gpu_e2e.py::test_batch_budget_scope_matches_the_reference_batch_semantics and
    test_gpu_train.py::test_adv_train_step_runs_and_tracks_oracle."""
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    onet, pnet = pair("hyper", 1, dev)
    x = images(n_img, 256, 256, dev)     # n_img = 2: the batch-mean budget test and the 1/B gradient scale over 300 steps
    # same quantisation noise in both implementations
    g = torch.Generator(device=dev).manual_seed(5)
    with torch.no_grad():
        onet.eval()
        y = onet.g_a(x)
        z = onet.h_a(torch.abs(y))
    ny = torch.rand(y.shape, device=dev, generator=g) - 0.5
    nz = torch.rand(z.shape, device=dev, generator=g) - 0.5
    for net in (onet, pnet):
        net.gaussian_conditional.noise_override = ny
        net.entropy_bottleneck.noise_override = nz
    args = oatk.default_args(model="hyper", quality=1, metric="mse", steps=300, lr_train=1e-5, adv=True, noise=1e-4)
    lm = ptr.LAMBDA_MSE[1]
    oopt, oaux = oatk.configure_optimizers(onet, args.lr_train)
    popt, paux = ptr.configure_optimizers(pnet, args)
    before = {n: q.detach().clone() for n, q in onet.named_parameters()}
    oout, oa = oatk.adv_train_step(x, onet, args, oatk.RateDistortionLoss("mse", lm).to(dev), oopt, oaux)
    pout, pa = ptr.adv_train_step(x, pnet, args, ptr.RateDistortionLoss("mse", lm), popt, paux)
    # every parameter gradient of the update (the oracle's are clipped in place by clip_grad_norm_, the product clips
    # inside its Adam kernel: compare directions, each side normalised by its own global norm)
    og = {n: q.grad.detach() for n, q in onet.named_parameters() if q.grad is not None and not n.endswith(".quantiles")}
    pgrads = {}
    for q, (off, cnt) in zip(popt.params, popt._spans):
        pgrads[id(q)] = popt.flat_grad[off:off + cnt].view(q.shape)
    pg = {n: pgrads[id(q)] for n, q in pnet.named_parameters() if id(q) in pgrads}
    assert set(og) == set(pg)
    on = math.sqrt(sum(float(v.pow(2).sum()) for v in og.values()))
    pn = math.sqrt(sum(float(v.pow(2).sum()) for v in pg.values()))
    errs = {n: relerr(pg[n] / pn, og[n] / on) for n in og if float(og[n].abs().max()) > 0}
    # codec stacks and the factorised prior: the attacked batch the two updates see differs at the 1e-4 level (two
    # 300-step trajectories), nothing else.  h_a / h_s get the rate gradient through GaussianConditional's
    # LowerBound(0.11) gate on the scales, which flips for the scales that sit on the bound (Precision): they
    # are pinned on IDENTICAL inputs by tests/test_gpu_train.py and only bounded loosely here.
    stack = max((e, n) for n, e in errs.items() if not n.startswith(("h_a", "h_s")))
    hyper = max((e, n) for n, e in errs.items() if n.startswith(("h_a", "h_s")))
    assert stack[0] < 1e-2, stack
    assert hyper[0] < 0.3, hyper
    for k in ("loss", "bpp_loss", "distortion_loss"):
        assert abs(float(pout[k]) - float(oout[k])) <= REL * abs(float(oout[k])), (k, float(pout[k]), float(oout[k]))
    assert abs(float(pa) - float(oa)) <= 1e-5 * abs(float(oa))
    # the update itself: first Adam step moves every weight by <= lr; the two implementations agree except for the
    # sign of gradients that are zero to fp32 accuracy
    lr, tot, cnt = args.lr_train, 0.0, 0
    op = dict(onet.named_parameters())
    for n, q in pnet.named_parameters():
        if n.endswith(".quantiles"):
            continue
        assert float((op[n].detach() - before[n]).abs().max()) <= 1.01 * lr
        tot += float((q.detach() - op[n].detach()).abs().sum()); cnt += q.numel()
    # first Adam step = lr * g / (|g| + 1e-8): parameters whose gradient is ~1e-8 (dead units of a random-init codec)
    # move by an amount that depends on the last digits of g; measured ~3 % of lr on average
    assert tot / cnt < 0.05 * lr, tot / cnt / lr


def test_attention_block_matches_oracle(dev):
    """compressai.layers.AttentionBlock (cheng2020_attn): forward, input gradient and every parameter gradient."""
    from imagecompression_adversarial_b200 import models as pm
    from oracle import layers as ol
    torch.manual_seed(11)
    o = ol.AttentionBlock(64).to(dev)
    p = pm.AttentionBlock(64).to(dev)
    p.load_state_dict(o.state_dict())
    x = torch.randn(2, 64, 24, 40, device=dev)
    gout = torch.randn(2, 64, 24, 40, device=dev)
    res = []
    for net in (o, p):
        xi = x.clone().requires_grad_(True)
        with pm._param_grads_on(True):
            out = net(xi)
        out.backward(gout)
        res.append((out.detach(), xi.grad.detach(), {k: v.grad.detach() for k, v in net.named_parameters()}))
    (oo, og, ow), (po, pg, pw) = res
    assert relerr(po, oo) < 2e-5 and relerr(pg, og) < 1e-4, (relerr(po, oo), relerr(pg, og))
    for k in ow:
        assert relerr(pw[k], ow[k]) < 2e-4, (k, relerr(pw[k], ow[k]))
