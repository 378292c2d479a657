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
