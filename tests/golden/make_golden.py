"""Generate golden fixtures by running the REFERENCE's own code (read-only, /root/reference).

Run in the build container only:  ``python tests/golden/make_golden.py``.
The reference cannot travel to the GPU box, so its outputs are committed as small ``.npz``/``.json``
fixtures next to this script and replayed by ``tests/test_oracle_golden.py``.

What runs from the reference, unmodified:
  * ``utils/ops.py``          Low_bound / Up_bound (fwd + custom bwd), GDN (fwd + autograd bwd)
  * ``utils/torch_msssim.py`` MS_SSIM  (variant 2; its hard-coded ``.cuda()`` is made a no-op)
  * ``anchors/utils.py``      conv / deconv hyper-parameters
  * ``attack_rd.py``          ``attack_`` + ``attack_our`` and ``self_ensemble.eval`` -- the loop itself.
    Its missing third-party imports are satisfied by stand-ins: ``compressai.*`` -> the oracle's
    model classes (so the fixture pins the LOOP, not the CompressAI arithmetic), ``pytorch_msssim``
    -> oracle variant 1, ``lpips``/``thop``/``matplotlib`` -> inert stubs.
"""
import json
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF = "/root/reference"
sys.path.insert(0, ROOT)

from oracle import attack as oatk  # noqa: E402
from oracle import layers as olayers  # noqa: E402
from oracle import models as omodels  # noqa: E402
from oracle import msssim as omsssim  # noqa: E402


def install_stand_ins():
    def mod(name, **attrs):
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m
        return m

    zoo = dict(
        bmshj2018_factorized=lambda quality, metric="mse", pretrained=False: omodels.init_model("factorized", quality),
        bmshj2018_hyperprior=lambda quality, metric="mse", pretrained=False: omodels.init_model("hyper", quality),
        mbt2018=lambda quality, metric="mse", pretrained=False: omodels.init_model("context", quality),
        cheng2020_anchor=lambda quality, metric="mse", pretrained=False: omodels.init_model("cheng2020", quality),
    )
    mod("compressai")
    mod("compressai.zoo", **zoo)
    mod("compressai.models", CompressionModel=omodels.CompressionModel, FactorizedPrior=omodels.FactorizedPrior,
        MeanScaleHyperprior=omodels.ScaleHyperprior, ScaleHyperprior=omodels.ScaleHyperprior)
    mod("compressai.layers", GDN=olayers.GDN)
    mod("compressai.datasets", ImageFolder=object)
    mod("pytorch_msssim", ms_ssim=omsssim.ms_ssim, MS_SSIM=omsssim.MS_SSIM)

    class _LPIPS(torch.nn.Module):
        def __init__(self, *a, **k):
            super().__init__()
    mod("lpips", LPIPS=_LPIPS)
    mod("thop", profile=lambda *a, **k: (0, 0))
    plt = mod("matplotlib.pyplot")
    mod("matplotlib", pyplot=plt)
    # the reference hard-codes .cuda(); on this CPU-only container make it the identity
    torch.Tensor.cuda = lambda self, *a, **k: self
    torch.nn.Module.cuda = lambda self, *a, **k: self


def golden_ops(out):
    from utils import ops as rops
    g = torch.Generator().manual_seed(7)
    x = (torch.rand(4, 3, 8, 8, generator=g) * 3 - 1).requires_grad_(True)
    gy = torch.randn(4, 3, 8, 8, generator=g)
    y = rops.Up_bound.apply(rops.Low_bound.apply(x, 0.0), 1.0)
    y.backward(gy)
    out["clamp_x"], out["clamp_gy"] = x.detach().numpy(), gy.numpy()
    out["clamp_y"], out["clamp_gx"] = y.detach().numpy(), x.grad.numpy()
    # eps-clamp as used at attack_rd.py:507
    n = (torch.randn(2, 3, 8, 8, generator=g) * 0.08).requires_grad_(True)
    gn = torch.randn(2, 3, 8, 8, generator=g)
    e = 16 / 255.0
    nc = rops.Up_bound.apply(rops.Low_bound.apply(n, -e), e)
    nc.backward(gn)
    out["eps_n"], out["eps_gn"] = n.detach().numpy(), gn.numpy()
    out["eps_nc"], out["eps_gnin"] = nc.detach().numpy(), n.grad.numpy()
    # the reference's in-repo GDN (utils/ops.py:58-97), forward + autograd backward, both modes
    for inv in (False, True):
        C = 8
        gdn = rops.GDN(C, inverse=inv)
        with torch.no_grad():
            gdn.gama.add_(torch.rand(C, C, 1, 1, generator=g) * 0.05)
            gdn.beta.add_(torch.rand(C, generator=g) * 0.5)
        xx = torch.randn(2, C, 6, 5, generator=g).requires_grad_(True)
        gg = torch.randn(2, C, 6, 5, generator=g)
        yy = gdn(xx)
        yy.backward(gg)
        k = "igdn" if inv else "gdn"
        out[k + "_gamma_raw"], out[k + "_beta_raw"] = gdn.gama.detach().numpy(), gdn.beta.detach().numpy()
        out[k + "_x"], out[k + "_gy"] = xx.detach().numpy(), gg.numpy()
        out[k + "_y"], out[k + "_gx"] = yy.detach().numpy(), xx.grad.numpy()


def golden_msssim_v2(out):
    from utils import torch_msssim as rms
    g = torch.Generator().manual_seed(11)
    a = torch.rand(2, 3, 96, 80, generator=g)
    b = (a + 0.05 * torch.randn(2, 3, 96, 80, generator=g)).clamp(0, 1)
    m = rms.MS_SSIM(max_val=1.0)
    out["ms2_a"], out["ms2_b"] = a.numpy(), b.numpy()
    out["ms2_val"] = np.float32(m(a, b).item())


def golden_conv_helpers(meta):
    from anchors import utils as rutils
    c, d = rutils.conv(3, 8), rutils.deconv(8, 3)
    meta["conv"] = dict(k=c.kernel_size, s=c.stride, p=c.padding)
    meta["deconv"] = dict(k=d.kernel_size, s=d.stride, p=d.padding, op=d.output_padding)


def golden_loop(out, meta):
    """Run the reference's attack_ (attack_rd.py:381-575) with the oracle's codecs as `net`."""
    import attack_rd
    import coder
    cases = [
        dict(name="hyper_q3_L2", model="hyper", quality=3, size=(192, 192), steps=12, att_metric="L2", noise=1e-4),
        dict(name="fact_q1_L2", model="factorized", quality=1, size=(176, 176), steps=9, att_metric="L2", noise=1e-4),
        dict(name="hyper_q3_msssim", model="hyper", quality=3, size=(192, 192), steps=6, att_metric="ms-ssim", noise=2e-5),
    ]
    for c in cases:
        argv = ["-m", c["model"], "-q", str(c["quality"]), "-metric", "mse", "--new", "-steps", str(c["steps"]),
                "-att_metric", c["att_metric"], "-noise", str(c["noise"]), "-device", "cpu"]
        args = coder.config().parse_args(argv)
        torch.manual_seed(0)
        net = omodels.init_model(c["model"], c["quality"], seed=0)
        im_s = oatk.synthetic_image(3, *c["size"])
        rec = []
        orig = attack_rd.attack_our

        def spy(*a, **k):
            r = orig(*a, **k)
            rec.append((float(r[0].detach()), float(r[1].detach())))
            return r
        attack_rd.attack_our = spy
        try:
            im_adv, output_adv, output_s, bpp_ori, bpp, mse_results, vi_results = attack_rd.attack_(im_s, net, args)
        finally:
            attack_rd.attack_our = orig
        n = c["name"]
        out[n + "_trace"] = np.array(rec, dtype=np.float64)
        out[n + "_im_adv_sub"] = im_adv[0, :, ::4, ::4].numpy()
        out[n + "_output_s_sub"] = output_s[0, :, ::4, ::4].numpy()
        meta[n] = dict(case=c, bpp_ori=float(bpp_ori), bpp=float(bpp), mse=mse_results, vi=vi_results,
                       im_adv_sum=float(im_adv.double().sum()), output_adv_sum=float(output_adv.double().sum()))
        print(n, meta[n])


def main():
    install_stand_ins()
    sys.path.insert(0, REF)
    cwd = os.getcwd()
    os.chdir("/tmp")  # the reference writes ./ckpts, ./logs relative paths on import paths
    out, meta = {}, {"torch": torch.__version__}
    golden_ops(out)
    golden_msssim_v2(out)
    golden_conv_helpers(meta)
    golden_loop(out, meta)
    os.chdir(cwd)
    np.savez_compressed(os.path.join(HERE, "reference_golden.npz"), **out)
    with open(os.path.join(HERE, "reference_golden.json"), "w") as f:
        json.dump(meta, f, indent=1, sort_keys=True)
    print("wrote", len(out), "arrays")


if __name__ == "__main__":
    main()
