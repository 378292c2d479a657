"""Drop-in check: the reference's own `attack_rd.py` CLI, UNMODIFIED, running on this package's operator surface
(stand-ins for the missing compressai / pytorch_msssim / lpips / thop / matplotlib imports).

Needs a copy of the reference's host scripts under baseline/_ref/reference (git-ignored; it travels to the GPU
box with the snapshot).  Skipped when absent -- nothing here reads /root/reference."""
import os
import re
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref", "reference")


def _png(tmp_path, i, h, w):
    import numpy as np
    from PIL import Image
    from oracle.attack import synthetic_image
    x = synthetic_image(i, h, w)[0].permute(1, 2, 0).numpy()
    p = os.path.join(tmp_path, f"img{i:02d}.png")
    Image.fromarray(np.round(x * 255).astype("uint8")).save(p)
    return p


@pytest.mark.parametrize("fused", [False, True])
def test_reference_cli_runs_unmodified(tmp_path, fused):
    if not os.path.exists(os.path.join(REF, "attack_rd.py")):
        pytest.skip("no reference copy under baseline/_ref/reference")
    for i in range(2):
        _png(str(tmp_path), i, 192, 256)
    cmd = [sys.executable, "-m", "imagecompression_adversarial_b200.launch", "--ref", REF]
    if fused:
        cmd.append("--fused")
    cmd += ["attack_rd.py", "-m", "hyper", "-q", "3", "-metric", "mse", "--new", "-steps", "9", "-noise", "1e-4",
            "-s", os.path.join(str(tmp_path), "*.png")]
    env = dict(os.environ, PYTHONPATH=ROOT)
    r = subprocess.run(cmd, cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    lines = [l for l in r.stdout.splitlines() if "Time:" in l]
    assert len(lines) == 2, r.stdout
    avg = [l for l in r.stdout.splitlines() if l.startswith("AVG:")]
    assert len(avg) == 1 and "hyper-mse-3" in avg[0]
    # per-image line: <file> bpp_ori bpp vi vi_msim Time: t   (attack_rd.py:670)
    nums = re.findall(r"[-+]?\d+\.\d+(?:e[-+]?\d+)?", lines[0].split("png", 1)[1])
    assert len(nums) >= 4
    bpp_ori, bpp, vi = float(nums[0]), float(nums[1]), float(nums[2])
    assert 0.0 < bpp_ori < 64.0 and 0.0 < bpp < 64.0 and vi > 0.0   # the attack amplifies the distortion


def _reference_module(name):
    """Import a module of the reference copy with the launcher's stand-ins for its missing third-party imports."""
    import importlib
    from imagecompression_adversarial_b200 import launch
    launch.install_shims()
    if REF not in sys.path:
        sys.path.insert(0, REF)
    return importlib.import_module(name)


def test_reference_self_ensemble_defence_on_the_operator_surface():
    """the reference's geometric self-ensemble (self_ensemble.py:34-131: 8 flips / rot90,
    two batched net(x) calls, best-MSE pick), UNMODIFIED, on this package's codec vs on the oracle codec."""
    if not os.path.exists(os.path.join(REF, "self_ensemble.py")):
        pytest.skip("no reference copy under baseline/_ref/reference")
    import torch
    from imagecompression_adversarial_b200 import models as pm
    from oracle import models as om
    from oracle.attack import synthetic_image
    torch.backends.cudnn.allow_tf32 = False
    torch.backends.cuda.matmul.allow_tf32 = False
    det = torch.are_deterministic_algorithms_enabled()
    se = _reference_module("self_ensemble")     # the module switches torch.use_deterministic_algorithms on at import (:31)
    try:
        _self_ensemble_checks(se, torch, pm, om, synthetic_image)
    finally:
        torch.use_deterministic_algorithms(det)  # do not leak the global into the rest of the test session


def _self_ensemble_checks(se, torch, pm, om, synthetic_image):
    dev = torch.device("cuda:0")
    onet = om.init_model("hyper", 3, seed=0).to(dev).eval()
    pnet = pm.init_model("hyper", 3, "mse", pretrained=False).to(dev).eval()
    pnet.load_state_dict(onet.state_dict())
    x = synthetic_image(3, 128, 192).to(dev)            # non-square: the rot90 members have the transposed shape
    with torch.no_grad():
        o_mse, o_x, o_xhat, o_lik = se.self_ensemble(onet, x)
        p_mse, p_x, p_xhat, p_lik = se.self_ensemble(pnet, x)
    assert abs(float(p_mse) - float(o_mse)) <= 2e-3 * float(o_mse)
    assert torch.equal(p_x, o_x)                          # same ensemble member picked
    assert p_xhat.shape == o_xhat.shape == x.shape
    # eval mode rounds the latent: a TF32-vs-fp32 near-tie flips isolated symbols, so compare reconstruction quality
    import math
    psnr = lambda a: -10.0 * math.log10(float(torch.mean((a - x) ** 2)))
    assert abs(psnr(p_xhat) - psnr(o_xhat)) < 0.05
    assert set(p_lik) == set(o_lik) == {"y", "z"}
    for k in o_lik:
        assert p_lik[k].shape == o_lik[k].shape
