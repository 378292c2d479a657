"""Parity at BASELINE.json's full size (hyperprior q3, 768x512 images).

The oracle (plain torch fp32, TF32 off) is run on the same GPU for ONE full-size image -- a few seconds -- and the
fused loop is compared with it step by step; the 8-image shard of the headline batch is then checked through
size-independent properties: an image's trajectory does not depend on which batch it rides in (bit-exact), a CUDA
graph replay equals the eager launches (bit-exact), and a second run of the same call reproduces the first bit for bit.
The time the oracle takes on the GPU (stock PyTorch eager, the like-for-like GPU bar of SURVEY.md section 8d) is printed.
"""
import math
import time

import pytest
import torch

pytestmark = pytest.mark.gpu

H, W = 512, 768


@pytest.fixture(scope="module")
def dev():
    from imagecompression_adversarial_b200 import ops
    ops.require_device()
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    return torch.device("cuda:0")


@pytest.fixture(scope="module")
def nets(dev):
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    onet = om.init_model("hyper", 3, seed=0).to(dev)
    pnet = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev)
    pnet.load_state_dict(onet.state_dict(), strict=True)
    return onet, pnet


def images(n, dev):
    from oracle.attack import synthetic_image
    return torch.cat([synthetic_image(i, H, W) for i in range(n)]).to(dev)


def psnr(a, b):
    return -10.0 * math.log10(float(torch.mean((a - b) ** 2)) + 1e-30)


def test_fullsize_forced_network_branch_matches_oracle(dev, nets):
    """Every iteration through the network (the benchmarked step) on one 768x512 image: per-step loss within 1e-3
    relative of the fp32 oracle, loss_i within 2.5e-3 (TF32 sign flips), final noise PSNR within 0.05 dB."""
    from imagecompression_adversarial_b200.engine import AttackEngine
    from oracle import attack as oatk
    from oracle.layers import low_bound, up_bound
    onet, pnet = nets
    x = images(1, dev)
    steps = 6
    a = oatk.default_args(model="hyper", quality=3, metric="mse", steps=steps)
    output_s, _, _ = oatk.clean_pass(x, onet, a)
    pnet.train(); onet.train()
    eng = AttackEngine(pnet, 1, H, W, steps=steps, force_branch=1, use_graph=False)
    eng.load(x, output_s)
    rec = []
    eng.run(steps, record=rec)
    noise = torch.zeros_like(x, requires_grad=True)
    opt = torch.optim.Adam([noise], lr=a.lr_attack)
    sch = torch.optim.lr_scheduler.MultiStepLR(opt, [1, 2, 3], gamma=0.33)
    eps = a.epsilon / 255.0
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for i in range(steps):
        nc = up_bound(low_bound(noise, -eps), eps)
        im_in = up_bound(low_bound(x + nc, 0.0), 1.0)
        loss_i = torch.mean((x - im_in) ** 2)
        out = up_bound(low_bound(onet.g_s(onet.g_a(im_in)), 0.0), 1.0)
        loss = 1.0 - torch.mean((output_s - out) ** 2)
        opt.zero_grad(); loss.backward(); opt.step()
        if i % (steps // 3) == 0:
            sch.step()
        pli, pl = float(rec[i][1][0]), float(rec[i][2][0])
        assert int(rec[i][0][0]) == 1
        assert abs(pl - float(loss)) <= 1e-3 * abs(float(loss)), (i, pl, float(loss))
        assert abs(pli - float(loss_i)) <= 2.5e-3 * float(loss_i) + 1e-9, (i, pli, float(loss_i))
    torch.cuda.synchronize()
    print(f"\n[stock PyTorch eager fp32 on this GPU, oracle] {steps / (time.perf_counter() - t0):.1f} image-iterations/s "
          f"(1 image 768x512, forced network branch, wgrad on)")
    nc = torch.clamp(noise.detach(), -eps, eps)
    assert abs(psnr(torch.clamp(x + nc, 0, 1), x) - psnr(torch.clamp(x + torch.clamp(eng.noise.permute(0, 3, 1, 2), -eps, eps), 0, 1), x)) < 0.05


def test_fullsize_attack_call_matches_oracle(dev, nets):
    """The public attack_() call (natural branch mix) on one full-size image: trajectory, final PSNR and bpp."""
    from imagecompression_adversarial_b200 import attack as patk
    from oracle import attack as oatk
    onet, pnet = nets
    x = images(1, dev)
    a = oatk.default_args(model="hyper", quality=3, metric="mse", steps=9)
    rec, orec = [], []
    im_adv, out_adv, out_s, bpp_ori, bpp, mse, vi = patk.attack_(x, pnet, a, record=rec)
    o = oatk.attack_(x, onet, a, record=orec)
    for t, (br, loss, loss_i) in enumerate(orec):
        pb, pli, pl = int(rec[t][0][0]), float(rec[t][1][0]), float(rec[t][2][0])
        if (pb == 1) != (br == "B"):
            assert abs(loss_i - a.noise) < 2e-3 * a.noise, (t, loss_i, pli)
            return
        assert abs(pli - loss_i) <= 2.5e-3 * max(loss_i, 1e-7) + 1e-9, (t, pli, loss_i)
        assert abs(pl - loss) <= (1e-3 if pb == 1 else 2.5e-3) * abs(loss) + 1e-9, (t, pl, loss)
    assert abs(psnr(im_adv, x) - psnr(o[0], x)) < 0.05
    assert abs(psnr(out_adv, out_s) - psnr(o[1], o[2])) < 0.05
    assert abs(float(bpp_ori) - float(o[3])) < max(1e-3, 2e-3 * float(o[3]))
    assert abs(float(bpp) - float(o[4])) < max(1e-3, 5e-3 * float(o[4]))


def test_fullsize_latent_indices_match_oracle(dev, nets):
    """Eval-mode round(y) of a full-size image: equal to the fp32 oracle's outside a 5e-3 guard band around .5."""
    onet, pnet = nets
    x = images(1, dev)
    onet.eval(); pnet.eval()
    with torch.no_grad():
        yo, yp = onet.g_a(x), pnet.g_a(x)
    frac = (yo - torch.floor(yo) - 0.5).abs()
    safe = frac > 5e-3
    assert bool((torch.round(yo)[safe] == torch.round(yp)[safe]).all())
    assert float((torch.round(yo) != torch.round(yp)).float().mean()) < 5e-3


def _run(pnet, x, ref, steps, use_graph):
    from imagecompression_adversarial_b200.engine import AttackEngine
    eng = AttackEngine(pnet, x.shape[0], H, W, steps=steps, force_branch=1, use_graph=use_graph)
    eng.load(x, ref)
    eng.run(steps)
    torch.cuda.synchronize()
    return eng.noise.clone(), eng.st.loss_i.clone()


def test_fullsize_batch_independence_graph_and_rerun(dev, nets):
    """8 full-size images (one GPU's shard of the headline batch at 8 GPUs): image 3 alone == image 3 inside the
    batch, graph replay == eager launches, second run == first run -- all bit-exact."""
    _, pnet = nets
    pnet.train()
    x = images(8, dev)
    ref = torch.rand_like(x)
    n8, l8 = _run(pnet, x, ref, 4, use_graph=True)
    n8b, l8b = _run(pnet, x, ref, 4, use_graph=True)
    assert torch.equal(n8, n8b) and torch.equal(l8, l8b)
    n8e, l8e = _run(pnet, x, ref, 4, use_graph=False)
    assert torch.equal(n8, n8e) and torch.equal(l8, l8e)
    n1, l1 = _run(pnet, x[3:4].contiguous(), ref[3:4].contiguous(), 4, use_graph=False)
    assert torch.equal(n1[0], n8[3]) and torch.equal(l1[0], l8[3])
