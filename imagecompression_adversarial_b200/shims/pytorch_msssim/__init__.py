"""`pytorch_msssim` surface (attack_rd.py:19, self_ensemble.py:20, train.py:18) on the fused level kernel."""
from imagecompression_adversarial_b200.metrics import MS_SSIM, ms_ssim  # noqa: F401
