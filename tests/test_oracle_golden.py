"""Oracle vs fixtures produced by the reference's own code (tests/golden/make_golden.py)."""
import json
import os

import numpy as np
import pytest
import torch

from oracle import attack as oatk
from oracle import layers as ol
from oracle import models as om
from oracle import msssim as oms

G = os.path.join(os.path.dirname(__file__), "golden")


@pytest.fixture(scope="module")
def gold():
    return np.load(os.path.join(G, "reference_golden.npz")), json.load(open(os.path.join(G, "reference_golden.json")))


def T(a):
    return torch.from_numpy(np.asarray(a))


def test_clamp_ops_match_reference(gold):
    g, _ = gold
    x = T(g["clamp_x"]).requires_grad_(True)
    y = ol.up_bound(ol.low_bound(x, 0.0), 1.0)
    y.backward(T(g["clamp_gy"]))
    assert torch.equal(y.detach(), T(g["clamp_y"]))
    assert torch.equal(x.grad, T(g["clamp_gx"]))
    n = T(g["eps_n"]).requires_grad_(True)
    e = 16 / 255.0
    nc = ol.up_bound(ol.low_bound(n, -e), e)
    nc.backward(T(g["eps_gn"]))
    assert torch.equal(nc.detach(), T(g["eps_nc"]))
    assert torch.equal(n.grad, T(g["eps_gnin"]))


@pytest.mark.parametrize("inv", [False, True])
def test_gdn_matches_reference_inrepo_gdn(gold, inv):
    g, _ = gold
    k = "igdn" if inv else "gdn"
    C = g[k + "_beta_raw"].shape[0]
    m = ol.GDN(C, inverse=inv)
    with torch.no_grad():
        m.beta.copy_(T(g[k + "_beta_raw"]))
        m.gamma.copy_(T(g[k + "_gamma_raw"]).reshape(C, C))
    x = T(g[k + "_x"]).requires_grad_(True)
    y = m(x)
    y.backward(T(g[k + "_gy"]))
    # the reference's gamma bound is reparam_offset (2^-18), CompressAI's is sqrt(0+pedestal) = same
    torch.testing.assert_close(y.detach(), T(g[k + "_y"]), rtol=1e-6, atol=1e-6)
    torch.testing.assert_close(x.grad, T(g[k + "_gx"]), rtol=1e-5, atol=1e-6)


def test_msssim_v2_matches_reference(gold):
    g, _ = gold
    v = oms.ms_ssim_v2(T(g["ms2_a"]), T(g["ms2_b"]), max_val=1.0)
    assert abs(float(v) - float(g["ms2_val"])) < 2e-6


def test_conv_helpers_match_reference(gold):
    _, meta = gold
    c, d = ol.conv(3, 8), ol.deconv(8, 3)
    assert list(c.kernel_size) == meta["conv"]["k"] and list(c.stride) == meta["conv"]["s"] and list(c.padding) == meta["conv"]["p"]
    assert list(d.kernel_size) == meta["deconv"]["k"] and list(d.output_padding) == meta["deconv"]["op"]


@pytest.mark.parametrize("name", ["hyper_q3_L2", "fact_q1_L2", "hyper_q3_msssim"])
def test_loop_matches_reference_attack(gold, name):
    """oracle.attack.attack_ reproduces the reference's attack_rd.attack_ step by step."""
    g, meta = gold
    c = meta[name]["case"]
    args = oatk.default_args(model=c["model"], quality=c["quality"], metric="mse", steps=c["steps"],
                             att_metric=c["att_metric"], noise=c["noise"])
    torch.manual_seed(0)
    net = om.init_model(c["model"], c["quality"], seed=0)
    im_s = oatk.synthetic_image(3, *c["size"])
    rec = []
    im_adv, output_adv, output_s, bpp_ori, bpp, mse, vi = oatk.attack_(im_s, net, args, record=rec)
    trace = g[name + "_trace"]
    mine = np.array([(r[1], r[2]) for r in rec])
    np.testing.assert_allclose(mine, trace, rtol=2e-4, atol=1e-7)
    np.testing.assert_allclose(output_s[0, :, ::4, ::4].numpy(), g[name + "_output_s_sub"], atol=2e-5)
    np.testing.assert_allclose(im_adv[0, :, ::4, ::4].numpy(), g[name + "_im_adv_sub"], atol=2e-3)
    assert abs(float(bpp_ori) - meta[name]["bpp_ori"]) < 1e-3
    assert abs(float(bpp) - meta[name]["bpp"]) < 5e-3
    assert abs(mse["mse_in"] - meta[name]["mse"]["mse_in"]) < 1e-6
    assert abs(vi["vi"] - meta[name]["vi"]["vi"]) < 0.05


def test_param_counts_match_compressai_published():
    for key, want in om.PARAM_COUNTS.items():
        fam = key[0]
        q = next(q for q, cfg in om.ZOO[fam].items() if tuple(cfg) == tuple(key[1:]))
        assert om.count_parameters(om.init_model(fam, q)) == want, key


def test_lr_schedule_matches_torch_multisteplr():
    lrs = oatk.lr_schedule(1001)
    assert lrs[0] == 0.01 and abs(lrs[1] - 0.0033) < 1e-12 and abs(lrs[334] - 0.001089) < 1e-12
    assert abs(lrs[667] - 0.00035937) < 1e-12 and abs(lrs[1000] - 0.00035937) < 1e-12
