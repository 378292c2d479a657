"""Oracle MS-SSIM, both variants the reference uses.  Test infrastructure only.

* variant 1 -- ``pytorch_msssim.ms_ssim`` (un-vendored, unpinned; call sites
  ``/root/reference/attack_rd.py:336,362``, ``self_ensemble.py:225,228``, ``train.py:44,88``):
  separable 11-tap sigma=1.5 *valid* Gaussian, relu on cs/ssim, avg_pool2d with (H%2, W%2) padding,
  per-(batch, channel) product, then mean.
* variant 2 -- ``/root/reference/utils/torch_msssim.py:18-76``: 2-D window with zero "same"
  padding, global means, no relu, plain avg_pool2d(2, 2).
"""
import torch
import torch.nn.functional as F

WEIGHTS = (0.0448, 0.2856, 0.3001, 0.2363, 0.1333)


def gauss_window(size=11, sigma=1.5, dtype=torch.float32):
    c = torch.arange(size, dtype=dtype) - size // 2
    g = torch.exp(-(c ** 2) / (2 * sigma ** 2))
    return g / g.sum()


def _filter_valid(x, win):
    C = x.shape[1]
    w = win.to(x.dtype).to(x.device)
    x = F.conv2d(x, w.view(1, 1, -1, 1).expand(C, 1, -1, 1), groups=C)
    x = F.conv2d(x, w.view(1, 1, 1, -1).expand(C, 1, 1, -1), groups=C)
    return x


def _ssim_valid(X, Y, win, data_range, K=(0.01, 0.03)):
    C1, C2 = (K[0] * data_range) ** 2, (K[1] * data_range) ** 2
    mu1, mu2 = _filter_valid(X, win), _filter_valid(Y, win)
    mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
    s1 = _filter_valid(X * X, win) - mu1_sq
    s2 = _filter_valid(Y * Y, win) - mu2_sq
    s12 = _filter_valid(X * Y, win) - mu12
    cs_map = (2 * s12 + C2) / (s1 + s2 + C2)
    ssim_map = ((2 * mu12 + C1) / (mu1_sq + mu2_sq + C1)) * cs_map
    return ssim_map.flatten(2).mean(-1), cs_map.flatten(2).mean(-1)


def ms_ssim(X, Y, data_range=1.0, size_average=True, win_size=11, win_sigma=1.5,
            weights=WEIGHTS, K=(0.01, 0.03)):
    """Variant 1 (pytorch_msssim.ms_ssim)."""
    assert X.shape == Y.shape and X.dim() == 4
    assert min(X.shape[-2:]) > (win_size - 1) * 2 ** 4, "image too small for 5-level MS-SSIM"
    win = gauss_window(win_size, win_sigma)
    w = torch.tensor(weights, dtype=X.dtype, device=X.device)
    mcs = []
    for lvl in range(len(weights)):
        ssim_pc, cs = _ssim_valid(X, Y, win, data_range, K)
        if lvl < len(weights) - 1:
            mcs.append(torch.relu(cs))
            pad = [s % 2 for s in X.shape[2:]]
            X = F.avg_pool2d(X, kernel_size=2, padding=pad)
            Y = F.avg_pool2d(Y, kernel_size=2, padding=pad)
    ssim_pc = torch.relu(ssim_pc)
    stack = torch.stack(mcs + [ssim_pc], dim=0)  # [levels, B, C]
    val = torch.prod(stack ** w.view(-1, 1, 1), dim=0)
    return val.mean() if size_average else val.mean(1)


class MS_SSIM(torch.nn.Module):
    """pytorch_msssim.MS_SSIM module form (train.py:44)."""

    def __init__(self, data_range=1.0, size_average=True, channel=3, win_size=11, win_sigma=1.5):
        super().__init__()
        self.kw = dict(data_range=data_range, size_average=size_average, win_size=win_size,
                       win_sigma=win_sigma)

    def forward(self, X, Y):
        return ms_ssim(X, Y, **self.kw)


def ms_ssim_v2(img1, img2, max_val=1.0, levels=5):
    """Variant 2 (utils/torch_msssim.py:26-71): 2-D window, zero 'same' padding, global means."""
    w = torch.tensor(WEIGHTS, dtype=img1.dtype, device=img1.device)
    C = img1.shape[1]
    msssim, mcs = [], []
    for _ in range(levels):
        _, _, a, b = img1.shape
        ws = min(a, b, 11)
        g = gauss_window(ws, 1.5 * ws / 11, img1.dtype).to(img1.device)
        win = (g[:, None] * g[None, :]).expand(C, 1, ws, ws).contiguous()
        f = lambda t: F.conv2d(t, win, padding=ws // 2, groups=C)
        mu1, mu2 = f(img1), f(img2)
        mu1_sq, mu2_sq, mu12 = mu1 * mu1, mu2 * mu2, mu1 * mu2
        s1, s2, s12 = f(img1 * img1) - mu1_sq, f(img2 * img2) - mu2_sq, f(img1 * img2) - mu12
        C1, C2 = (0.01 * max_val) ** 2, (0.03 * max_val) ** 2
        V1, V2 = 2.0 * s12 + C2, s1 + s2 + C2
        msssim.append((((2 * mu12 + C1) * V1) / ((mu1_sq + mu2_sq + C1) * V2)).mean())
        mcs.append((V1 / V2).mean())
        img1, img2 = F.avg_pool2d(img1, 2, 2), F.avg_pool2d(img2, 2, 2)
    mcs_t, ms_t = torch.stack(mcs), torch.stack(msssim)
    return torch.prod(mcs_t[:levels - 1] ** w[:levels - 1]) * (ms_t[levels - 1] ** w[levels - 1])
