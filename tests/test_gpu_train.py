"""Codec update of adversarial training (train.py:335-366) on a B200: the product's train-mode
forward / RD loss / backward with parameter gradients / clip + Adam against the oracle (plain torch fp32 + torch.optim)
on shared weights, inputs and quantisation noise."""
import math

import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def dev():
    from imagecompression_adversarial_b200 import ops
    ops.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


def pair(model, quality, dev, seed=0):
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    onet = om.init_model(model, quality, seed=seed).to(dev)
    pnet = pm.init_model(model, quality, "mse", pretrained=False).to(dev)
    pnet.load_state_dict(onet.state_dict(), strict=True)
    return onet, pnet


def share_noise(onet, pnet, x, dev, seed=5):
    """Same U(-.5,.5) quantisation noise in both implementations (train-mode forward)."""
    g = torch.Generator(device=dev).manual_seed(seed)
    with torch.no_grad():
        onet.eval()
        y = onet.g_a(x)
        z = None
        if hasattr(onet, "h_a"):                       # hyper: h_a(|y|); context / cheng2020: h_a(y) (anchors/model.py:92,98)
            z = onet.h_a(y if hasattr(onet, "context_prediction") else torch.abs(y))
    ny = torch.rand(y.shape, device=dev, generator=g) - 0.5
    for net in (onet, pnet):
        if z is not None:
            net.gaussian_conditional.noise_override = ny
            net.entropy_bottleneck.noise_override = torch.rand(z.shape, device=dev, generator=torch.Generator(device=dev).manual_seed(seed + 1)) - 0.5
        else:
            net.entropy_bottleneck.noise_override = ny


def rel(a, b):
    return float((a - b).pow(2).sum().sqrt() / b.pow(2).sum().sqrt().clamp(min=1e-30))


@pytest.mark.parametrize("mode", ["3xtf32", "tf32"])
@pytest.mark.parametrize("model,quality", [("context", 1), ("cheng2020", 1)])
def test_autoregressive_families_train_g_a_through_the_distortion_term(dev, model, quality, mode):
    """The train-mode quantiser ``y + noise`` is differentiable (compressai quantize(.,"noise"), anchors/model.py:102):
    the distortion gradient through g_s(y_hat) and context_prediction(y_hat) must reach g_a.  Every g_a parameter
    gradient of one RD-loss backward against the oracle, plus a direct check that the distortion term alone
    (lambda-weighted MSE, no rate) produces a non-zero g_a gradient."""
    from imagecompression_adversarial_b200 import precision
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    from oracle.attack import synthetic_image
    onet, pnet = pair(model, quality, dev)
    prev_mode = precision.get()
    precision.set(mode)
    try:
        _autoregressive_train_check(dev, model, quality, mode, onet, pnet, ptr, oatk, synthetic_image)
    finally:
        precision.set(prev_mode)


def _autoregressive_train_check(dev, model, quality, mode, onet, pnet, ptr, oatk, synthetic_image):
    if model == "cheng2020":
        # a random-init cheng2020 (20 contractions with residual adds and IGDN) blows its reconstruction up to ~1e10 and
        # the RD loss to ~1e21, where fp32 comparisons mean nothing: damp the synthesis weights in BOTH implementations
        with torch.no_grad():
            for n, q in onet.named_parameters():
                if n.startswith("g_s.") and n.endswith(".weight"):
                    q.mul_(0.5)
        pnet.load_state_dict(onet.state_dict(), strict=True)
    x = torch.cat([synthetic_image(i, 128, 128) for i in range(2)]).to(dev)
    share_noise(onet, pnet, x, dev)
    lm = ptr.LAMBDA_MSE[quality]
    ocrit, pcrit = oatk.RateDistortionLoss("mse", lm).to(dev), ptr.RateDistortionLoss("mse", lm)
    onet.train(); pnet.train()
    oout = ocrit(onet(x), x)
    assert float(oout["loss"]) < 1e8, float(oout["loss"])
    onet.zero_grad(); oout["loss"].backward()
    pout = pcrit(pnet(x), x)
    pnet.zero_grad(); pout["loss"].backward()
    og = {n: q.grad.detach().clone() for n, q in onet.named_parameters() if q.grad is not None}
    pg = {n: q.grad.detach().clone() for n, q in pnet.named_parameters() if q.grad is not None}
    ga = [n for n in og if n.startswith("g_a.") and float(og[n].abs().max()) > 0]
    assert ga and set(ga) <= set(pg)
    # parity mode: fp32-accurate contractions -> every g_a gradient tensor to 2e-3.  Speed mode: TF32 contractions; the
    # rate gradient reaching y depends on the scales through GaussianConditional's LowerBound(0.11) gate, which flips for
    # scales on the bound (Precision) -- context (8 layers) stays within 1e-2, cheng2020 (20 layers, damped
    # synthesis so the rate term dominates) is only bounded loosely there and pinned by the parity-mode run.
    bound = 2e-3 if mode == "3xtf32" else (0.5 if model == "cheng2020" else 1e-2)
    worst = max((rel(pg[n], og[n]), n) for n in ga)
    assert worst[0] < bound, worst
    assert abs(float(pout["loss"]) - float(oout["loss"])) <= (5e-4 if mode == "3xtf32" else 3e-3) * abs(float(oout["loss"]))
    # distortion term alone
    pnet.zero_grad()
    (pcrit(pnet(x), x)["distortion_loss"]).backward()
    assert float(pnet.g_a[0].conv1.weight.grad.abs().max() if model == "cheng2020" else
                 pnet.g_a[0].weight.grad.abs().max()) > 0


def test_optimizer_checkpoint_resume_continues_identically(dev):
    """train.py:443-454 / coder.py:104-116: save {state_dict, optimizer, aux_optimizer} after two steps, resume in a
    fresh model + optimisers, take a third step in both -- bit-identical weights."""
    import io
    from imagecompression_adversarial_b200 import models as pm
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    from oracle.attack import synthetic_image
    torch.manual_seed(0)
    net = pm.init_model("hyper", 1, "mse", pretrained=False).to(dev)
    x = torch.cat([synthetic_image(i, 64, 64) for i in range(2)]).to(dev)
    args = oatk.default_args(model="hyper", quality=1, metric="mse", lr_train=1e-4)
    crit = ptr.RateDistortionLoss("mse", ptr.LAMBDA_MSE[1])
    opt, aux = ptr.configure_optimizers(net, args)

    def step(net, opt, aux, seed):
        g = torch.Generator(device=dev).manual_seed(seed)
        net.train()
        net.gaussian_conditional.noise_override = torch.rand(2, 192, 4, 4, device=dev, generator=g) - 0.5
        net.entropy_bottleneck.noise_override = torch.rand(2, 128, 1, 1, device=dev, generator=g) - 0.5
        out = crit(net(x), x)
        opt.zero_grad(); aux.zero_grad()
        out["loss"].backward()
        opt.step()
        a = net.aux_loss(); a.backward(); aux.step(exchange=False)

    step(net, opt, aux, 1); step(net, opt, aux, 2)
    buf = io.BytesIO()
    torch.save({"epoch": 0, "step": 2, "state_dict": net.state_dict(), "optimizer": opt.state_dict(),
                "aux_optimizer": aux.state_dict()}, buf)
    buf.seek(0)
    ck = torch.load(buf, map_location=dev, weights_only=False)
    net2 = pm.init_model("hyper", 1, "mse", pretrained=False).to(dev)
    net2.load_state_dict(ck["state_dict"], strict=True)
    opt2, aux2 = ptr.configure_optimizers(net2, args)
    opt2.load_state_dict(ck["optimizer"]); aux2.load_state_dict(ck["aux_optimizer"])
    assert opt2.steps == 2 and opt2.param_groups[0]["lr"] == args.lr_train
    step(net, opt, aux, 3); step(net2, opt2, aux2, 3)
    for (n, a), (_, b) in zip(net.named_parameters(), net2.named_parameters()):
        assert torch.equal(a.detach(), b.detach()), n


@pytest.mark.parametrize("model,quality,metric", [("hyper", 1, "mse"), ("factorized", 1, "mse"), ("hyper", 3, "ms-ssim")])
def test_update_step_matches_oracle(dev, model, quality, metric):
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    from oracle.attack import synthetic_image
    onet, pnet = pair(model, quality, dev)
    hw = (192, 192) if metric == "ms-ssim" else (128, 128)
    x = torch.cat([synthetic_image(i, *hw) for i in range(2)]).to(dev)
    share_noise(onet, pnet, x, dev)
    lm = (ptr.LAMBDA_MSE if metric == "mse" else ptr.LAMBDA_MSSSIM)[quality]
    args = oatk.default_args(model=model, quality=quality, metric=metric, lr_train=1e-4)
    ocrit, pcrit = oatk.RateDistortionLoss(metric, lm).to(dev), ptr.RateDistortionLoss(metric, lm)
    oopt, oaux = oatk.configure_optimizers(onet, args.lr_train)
    popt, paux = ptr.configure_optimizers(pnet, args)
    before = {n: p.detach().clone() for n, p in onet.named_parameters()}
    # ---- oracle: train.py:351-365
    onet.train()
    oout = ocrit(onet(x), x)
    oopt.zero_grad(); oaux.zero_grad()
    oout["loss"].backward()
    ograds = {n: p.grad.detach().clone() for n, p in onet.named_parameters() if p.grad is not None}
    onorm = torch.nn.utils.clip_grad_norm_(onet.parameters(), 1.0)
    oopt.step()
    oa = onet.aux_loss(); oa.backward(); oaux.step()
    # ---- product
    pnet.train()
    pout = pcrit(pnet(x), x)
    popt.zero_grad(); paux.zero_grad()
    pout["loss"].backward()
    pgrads = {n: p.grad.detach().clone() for n, p in pnet.named_parameters() if p.grad is not None}
    popt.step()
    pa = pnet.aux_loss(); pa.backward(); paux.step()
    # ---- losses
    assert abs(float(pout["bpp_loss"]) - float(oout["bpp_loss"])) <= 2e-3 * abs(float(oout["bpp_loss"]))
    # ms-ssim of a random-init codec's output is ~1e-3 (the loss uses 1 - ms_ssim): absolute bound there
    dtol = 1e-4 if metric == "ms-ssim" else 3e-3 * abs(float(oout["distortion_loss"]))
    assert abs(float(pout["distortion_loss"]) - float(oout["distortion_loss"])) <= dtol
    assert abs(float(pout["loss"]) - float(oout["loss"])) <= 2e-3 * abs(float(oout["loss"]))
    assert abs(float(pa) - float(oa)) <= 1e-5 * abs(float(oa))
    # ---- every parameter gradient (TF32 contractions: relative L2 error per tensor)
    main = [n for n in ograds if not n.endswith(".quantiles")]
    assert set(main) <= set(pgrads), set(main) - set(pgrads)
    # Codec stacks and the factorised prior: TF32 contraction error only.  The hyper path (h_a, h_s) sees the rate term
    # through GaussianConditional's scale LowerBound(0.11): with random-init weights ~0.3 % of the scales sit within 1e-3
    # of the bound, so TF32-level differences in h_s's output flip that gate for a few elements and the rate gradient
    # (curvature ~1/s^2 at s ~ 0.11) moves by percents.  The kernels themselves are pinned to autograd at 1e-4 on
    # identical inputs in the tests below.
    errs = {n: rel(pgrads[n], ograds[n]) for n in main if float(ograds[n].abs().max()) > 0}
    stack = {n: e for n, e in errs.items() if not n.startswith(("h_a", "h_s"))}
    hyper = {n: e for n, e in errs.items() if n.startswith(("h_a", "h_s"))}
    # (ms-ssim of a random-init reconstruction is ~1e-3, where its relu / power terms are ill-conditioned: the distortion
    # gradient is pinned separately on well-conditioned inputs in test_msssim_distortion_term_matches_oracle)
    assert max(stack.values()) < (0.15 if metric == "ms-ssim" else 5e-3), max((e, n) for n, e in stack.items())
    if hyper:
        assert max(hyper.values()) < 0.15, max((e, n) for n, e in hyper.items())
    if metric == "ms-ssim":
        return
    # ---- global norm of clip_grad_norm_ and the Adam step
    pnorm = math.sqrt(float(popt.last_sumsq))
    assert abs(pnorm - float(onorm)) <= 5e-3 * float(onorm), (pnorm, float(onorm))
    lr = args.lr_train
    tot = cnt = 0.0
    for n, p in pnet.named_parameters():
        q = dict(onet.named_parameters())[n]
        if n.endswith(".quantiles"):
            torch.testing.assert_close(p.detach(), q.detach(), rtol=0, atol=2e-6)
            continue
        moved = (q.detach() - before[n]).abs()
        assert float(moved.max()) <= 1.01 * lr       # first Adam step: |delta| <= lr
        tot += float((p.detach() - q.detach()).abs().sum()); cnt += p.numel()
    assert tot / cnt < 0.08 * lr, tot / cnt / lr     # sign flips of near-zero gradients only


def test_adv_train_step_runs_and_tracks_oracle(dev):
    """train.py:335-366 end to end (attack_ + update) at test size; losses against the oracle's adv_train_step."""
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    from oracle.attack import synthetic_image
    onet, pnet = pair("hyper", 1, dev)
    x = torch.cat([synthetic_image(i, 192, 192) for i in range(2)]).to(dev)
    share_noise(onet, pnet, x, dev)
    args = oatk.default_args(model="hyper", quality=1, metric="mse", steps=6, lr_train=1e-5, adv=True, noise=1e-4)
    lm = ptr.LAMBDA_MSE[1]
    oopt, oaux = oatk.configure_optimizers(onet, args.lr_train)
    popt, paux = ptr.configure_optimizers(pnet, args)
    oout, oa = oatk.adv_train_step(x, onet, args, oatk.RateDistortionLoss("mse", lm).to(dev), oopt, oaux)
    pout, pa = ptr.adv_train_step(x, pnet, args, ptr.RateDistortionLoss("mse", lm), popt, paux)
    assert abs(float(pout["loss"]) - float(oout["loss"])) <= 5e-3 * abs(float(oout["loss"]))
    assert abs(float(pout["bpp_loss"]) - float(oout["bpp_loss"])) <= 5e-3 * abs(float(oout["bpp_loss"]))
    # the update moved the weights the engines read: a second attack must see the new parameters
    w0 = pnet.g_a[0].weight.detach().clone()
    ptr.adv_train_step(x, pnet, args, ptr.RateDistortionLoss("mse", lm), popt, paux)
    assert float((pnet.g_a[0].weight.detach() - w0).abs().max()) > 0


def test_gc_backward_matches_autograd(dev):
    """GaussianConditional likelihood backward kernel vs autograd of the oracle on identical inputs."""
    from imagecompression_adversarial_b200 import models as pm
    from oracle import layers as ol
    g = torch.Generator(device=dev).manual_seed(11)
    for with_means in (False, True):
        y = 3.0 * torch.randn(2, 64, 8, 12, device=dev, generator=g)
        sc = (0.6 * torch.rand(2, 64, 8, 12, device=dev, generator=g)).clamp(min=0.0)   # ~18 % below the 0.11 bound
        mu = torch.randn(2, 64, 8, 12, device=dev, generator=g) if with_means else None
        nz = torch.rand(2, 64, 8, 12, device=dev, generator=g) - 0.5
        up = torch.randn(2, 64, 8, 12, device=dev, generator=g)
        outs = []
        for mod in (ol.GaussianConditional().to(dev), pm.GaussianConditional().to(dev)):
            mod.train()
            mod.noise_override = nz
            yy, ss = y.clone().requires_grad_(True), sc.clone().requires_grad_(True)
            mm = mu.clone().requires_grad_(True) if with_means else None
            y_hat, lik = mod(yy, ss, means=mm)
            ((lik * up).sum() + (y_hat * up).sum() * 0.25).backward()
            outs.append((lik.detach(), yy.grad, ss.grad, mm.grad if with_means else None))
        o, p = outs
        torch.testing.assert_close(p[0], o[0], rtol=1e-4, atol=1e-7)
        for a, b in zip(p[1:], o[1:]):
            if b is not None:
                assert rel(a, b) < 1e-4, rel(a, b)


def test_eb_backward_matches_autograd(dev):
    """EntropyBottleneck backward kernel (x and all 14 parameter tensors) vs autograd of the oracle."""
    from imagecompression_adversarial_b200 import models as pm
    from oracle import layers as ol
    torch.manual_seed(4)
    C = 96
    oe, pe = ol.EntropyBottleneck(C).to(dev), pm.EntropyBottleneck(C).to(dev)
    with torch.no_grad():
        for i in range(4):
            getattr(oe, f"_factor{i}").uniform_(-0.5, 0.5)
        for i in range(5):
            getattr(oe, f"_matrix{i}").add_(0.3 * torch.randn_like(getattr(oe, f"_matrix{i}")))
    pe.load_state_dict(oe.state_dict())
    g = torch.Generator(device=dev).manual_seed(12)
    x = 4.0 * torch.randn(3, C, 6, 10, device=dev, generator=g)
    nz = torch.rand(3, C, 6, 10, device=dev, generator=g) - 0.5
    up = torch.randn(3, C, 6, 10, device=dev, generator=g)
    outs = []
    for mod in (oe, pe):
        mod.train()
        mod.noise_override = nz
        xx = x.clone().requires_grad_(True)
        x_hat, lik = mod(xx)
        ((torch.log(lik) * up).sum() + (x_hat * up).sum() * 0.1).backward()
        outs.append((lik.detach(), xx.grad, {n: q.grad for n, q in mod.named_parameters() if q.grad is not None}))
    o, p = outs
    torch.testing.assert_close(p[0], o[0], rtol=2e-4, atol=1e-7)
    assert rel(p[1], o[1]) < 2e-4, rel(p[1], o[1])
    assert set(o[2]) == set(p[2]) and len(o[2]) == 14, (sorted(o[2]), sorted(p[2]))
    for n in o[2]:
        assert p[2][n].shape == o[2][n].shape
        assert rel(p[2][n], o[2][n]) < 5e-4, (n, rel(p[2][n], o[2][n]))


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_parameter_gradients_match_autograd(dev, inverse):
    from imagecompression_adversarial_b200 import models as pm
    from oracle import layers as ol
    torch.manual_seed(6)
    C = 64
    og, pg = ol.GDN(C, inverse=inverse).to(dev), pm.GDN(C, inverse=inverse).to(dev)
    with torch.no_grad():
        og.gamma.add_(0.05 * torch.rand_like(og.gamma))
        og.beta.add_(0.2 * torch.rand_like(og.beta))
        og.gamma[0, 1] = 0.0      # below the reparametrisation bound: LowerBound backward rule
    pg.load_state_dict(og.state_dict())
    g = torch.Generator(device=dev).manual_seed(13)
    x = torch.randn(2, C, 20, 28, device=dev, generator=g)
    up = torch.randn(2, C, 20, 28, device=dev, generator=g)
    outs = []
    from imagecompression_adversarial_b200 import models as pmod
    for mod in (og, pg):
        xx = x.clone().requires_grad_(True)
        with pmod._param_grads_on(True):
            y = mod(xx)
        (y * up).sum().backward()
        outs.append((y.detach(), xx.grad, mod.beta.grad, mod.gamma.grad))
    o, p = outs
    assert rel(p[0], o[0]) < 2e-3
    for a, b, nm in zip(p[1:], o[1:], ("x", "beta", "gamma")):
        assert a is not None, nm
        assert rel(a, b) < 3e-3, (nm, rel(a, b))


@pytest.mark.parametrize("inverse", [False, True])
def test_gdn_parameter_gradients_tensor_path_matches_cuda_core_kernel(dev, inverse):
    """d gamma / d beta through the tcgen05 weight-gradient kernel (1x1 case: operands T and x^2 written by
    icadv_gdn_param_operands, TF32 products, fp32 accumulation) against the fp32 CUDA-core kernel on the same saved
    tensors, ragged 8x8 tile edges included; the LowerBound rule of the reparametrisation is exercised on both."""
    from imagecompression_adversarial_b200 import ops
    C, n, h, w = 128, 3, 37, 50
    g = torch.Generator(device=dev).manual_seed(17)
    gy = torch.randn(n, h, w, C, device=dev, generator=g)
    y = torch.randn(n, h, w, C, device=dev, generator=g)
    sc = 0.5 + torch.rand(n, h, w, C, device=dev, generator=g)
    beta_raw = 1.0 + torch.rand(C, device=dev, generator=g)
    gamma_raw = 0.1 * torch.rand(C, C, device=dev, generator=g)
    gamma_raw[0, :8] = 0.0            # below the bound
    kw = dict(inverse=inverse, beta_bound=1e-3, gamma_bound=3.8e-6 ** 0.5)
    res = {}
    for tc in (True, False):
        ops.GDN_PARAM_GRAD_TC = tc
        try:
            res[tc] = ops.gdn_param_grad(gy, y, sc, beta_raw, gamma_raw, **kw)
        finally:
            ops.GDN_PARAM_GRAD_TC = True
    for a, b, nm in zip(res[True], res[False], ("beta", "gamma")):
        assert float(b.abs().max()) > 0
        assert rel(a, b) < 1e-3, (nm, rel(a, b))


def test_fused_clip_adam_matches_torch(dev):
    from imagecompression_adversarial_b200 import training as ptr
    torch.manual_seed(8)
    shapes = [(64, 32, 5, 5), (64,), (32, 32), (7, 3, 1)]
    ref = [torch.nn.Parameter(torch.randn(*s, device=dev)) for s in shapes]
    mine = [torch.nn.Parameter(q.detach().clone()) for q in ref]
    topt = torch.optim.Adam(ref, lr=3e-4)
    fopt = ptr.FusedAdamClip(mine, lr=3e-4, max_norm=1.0)
    for step in range(4):
        topt.zero_grad(); fopt.zero_grad()
        scale = 10.0 if step % 2 == 0 else 1e-3        # with and without clipping
        for a, b in zip(ref, mine):
            gr = scale * torch.randn_like(a)
            a.grad = gr.clone()
            b.grad.copy_(gr)
        tn = torch.nn.utils.clip_grad_norm_(ref, 1.0)
        topt.step()
        fopt.step()
        assert abs(math.sqrt(float(fopt.last_sumsq)) - float(tn)) <= 1e-5 * float(tn)
        for a, b in zip(ref, mine):
            torch.testing.assert_close(b.detach(), a.detach(), rtol=0, atol=2e-7)


def test_msssim_distortion_term_matches_oracle(dev):
    """RateDistortionLoss(metric="ms-ssim") distortion term and its gradient (train.py:44,88) on two similar images."""
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    g = torch.Generator(device=dev).manual_seed(31)
    t = torch.rand(2, 3, 192, 224, device=dev, generator=g)
    x = (t + 0.05 * torch.randn(2, 3, 192, 224, device=dev, generator=g)).clamp(0, 1)
    lik = {"y": 0.2 + 0.7 * torch.rand(2, 8, 12, 14, device=dev, generator=g)}
    outs = []
    for crit in (oatk.RateDistortionLoss("ms-ssim", 8.73).to(dev), ptr.RateDistortionLoss("ms-ssim", 8.73)):
        xx = x.clone().requires_grad_(True)
        ll = lik["y"].clone().requires_grad_(True)
        out = crit({"x_hat": xx, "likelihoods": {"y": ll}}, t)
        out["loss"].backward()
        outs.append((float(out["loss"]), float(out["distortion_loss"]), float(out["bpp_loss"]), xx.grad, ll.grad))
    o, p = outs
    assert abs(p[0] - o[0]) <= 1e-4 * abs(o[0]) and abs(p[1] - o[1]) <= 2e-5 and abs(p[2] - o[2]) <= 1e-5 * abs(o[2])
    assert rel(p[3], o[3]) < 3e-3, rel(p[3], o[3])
    assert rel(p[4], o[4]) < 1e-5, rel(p[4], o[4])


@pytest.mark.parametrize("adv", [False, True])
def test_test_epoch_matches_oracle(dev, tmp_path, adv):
    """train.py:196-242: RD-loss averages without --adv, mean VI of the attacked batches with it."""
    from imagecompression_adversarial_b200 import models as pm
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    from oracle import models as om
    from oracle.attack import synthetic_image
    onet = om.init_model("hyper", 1, seed=0).to(dev)
    pnet = pm.init_model("hyper", 1, "mse", pretrained=False).to(dev)
    pnet.load_state_dict(onet.state_dict())
    loader = [torch.cat([synthetic_image(2 * b + i, 192, 192) for i in range(2)]) for b in range(2)]
    args = oatk.default_args(model="hyper", quality=1, metric="mse", steps=6, adv=adv, noise=1e-3)
    log = str(tmp_path / "log.txt")
    p = ptr.test_epoch(0, loader, pnet, ptr.RateDistortionLoss(metric="mse", lmbda=0.0018), log, args)
    o, avg = oatk.test_epoch(loader, onet, oatk.RateDistortionLoss(metric="mse", lmbda=0.0018), args)
    assert args.noise == 1e-3                                             # the forced budget is restored (:216)
    assert "Test epoch 0" in open(log).read()
    if adv:
        assert abs(p - o) < 0.1, (p, o)                                   # VI in dB over a 6-step attack
    else:
        assert abs(p - o) <= 2e-3 * abs(o), (p, o)


def test_batch_test_matches_oracle(dev):
    """test.py:28-60: the RD evaluation driver -- per-image eval forward, bpp / PSNR / MS-SSIM,
    the four averages of its AVG: line."""
    from imagecompression_adversarial_b200 import models as pm
    from imagecompression_adversarial_b200 import training as ptr
    from oracle import attack as oatk
    from oracle import models as om
    from oracle.attack import synthetic_image
    onet = om.init_model("hyper", 3, seed=0).to(dev)
    pnet = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev)
    pnet.load_state_dict(onet.state_dict())
    imgs = [synthetic_image(i, 192, 256).to(dev) for i in range(3)]
    p = ptr.batch_test(imgs, pnet)
    o = oatk.batch_test(imgs, onet)
    assert abs(p[0] - o[0]) <= max(1e-3, 2e-3 * o[0]), (p[0], o[0])      # bpp
    assert abs(p[1] - o[1]) < 0.05, (p[1], o[1])                          # PSNR, dB
    assert abs(p[2] - o[2]) < 1e-3, (p[2], o[2])                          # MS-SSIM
    assert abs(p[3] - o[3]) < 0.05, (p[3], o[3])                          # MS-SSIM, dB
