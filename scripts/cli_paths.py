"""Developer measurement: the reference's UNMODIFIED attack_rd.py CLI on this package's operator surface, both ways --
as it is (attack_our on the operator surface: autograd walk, the reference's per-step host syncs) and with --fused (the
launcher swaps attack_rd.attack_ for the device-resident loop) -- iterations per second from the CLI's own "Time:" column.
Natural branch mix (the reference's schedule), one 768x512 image at a time as the CLI does."""
import os
import re
import subprocess
import sys
import tempfile

import numpy as np
from PIL import Image

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.attack import synthetic_image  # noqa: E402  (image generator only)

REF = os.path.join(ROOT, "baseline", "_ref", "reference")
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 100
with tempfile.TemporaryDirectory() as tmp:
    for i in range(3):
        x = synthetic_image(i, 512, 768)[0].permute(1, 2, 0).numpy()
        Image.fromarray(np.round(x * 255).astype("uint8")).save(os.path.join(tmp, f"img{i:02d}.png"))
    for metric in ("L2", "ms-ssim"):
        for fused in (False, True):
            cmd = [sys.executable, "-m", "imagecompression_adversarial_b200.launch", "--ref", REF]
            if fused:
                cmd.append("--fused")
            cmd += ["attack_rd.py", "-m", "hyper", "-q", "3", "-metric", "mse", "--new", "-steps", str(steps), "-noise", "1e-4",
                    "-att_metric", metric, "-s", os.path.join(tmp, "*.png")]
            r = subprocess.run(cmd, cwd=tmp, env=dict(os.environ, PYTHONPATH=ROOT), capture_output=True, text=True, timeout=1200)
            if r.returncode != 0:
                print(metric, "fused" if fused else "unfused", "FAILED", r.stderr[-800:])
                continue
            times = [float(re.findall(r"Time:\s*([0-9.eE+-]+)", l)[0]) for l in r.stdout.splitlines() if "Time:" in l]
            # first image includes start-up (library load, plan creation, graph capture): report the later ones
            rest = times[1:] if len(times) > 1 else times
            print(f"attack_rd.py -att_metric {metric:7s} {'--fused' if fused else 'as is  '}: {steps} steps per image, "
                  f"per-image times {['%.2f' % t for t in times]} s -> {steps / (sum(rest) / len(rest)):8.1f} iterations/s "
                  f"(images after the first)")
