"""Synthetic inputs of the benchmark and the CLI smoke runs: seeded, Kodak-shaped images on
the k/255 lattice that ``coder.read_image`` (coder.py:20-33) produces from an 8-bit PNG.

Generator (per image i): ``torch.Generator().manual_seed(1234 + i)``; U[0,1) field [3,H,W] -> separable Gaussian blur
sigma = 3 px (reflect padding) -> affine rescale to [0.05, 0.95] -> round(. * 255) / 255.  Host-side set-up, never
inside a timed region."""
import torch


def synthetic_image(i, H=512, W=768):
    """[1, 3, H, W] fp32 on the CPU."""
    g = torch.Generator().manual_seed(1234 + i)
    x = torch.rand(1, 3, H, W, generator=g)
    r = 9
    k = torch.exp(-(torch.arange(-r, r + 1, dtype=torch.float32) ** 2) / (2 * 3.0 ** 2))
    k = k / k.sum()
    xp = torch.nn.functional.pad(x, (r, r, r, r), mode="reflect")
    xp = torch.nn.functional.conv2d(xp, k.view(1, 1, 1, -1).expand(3, 1, 1, -1), groups=3)
    xp = torch.nn.functional.conv2d(xp, k.view(1, 1, -1, 1).expand(3, 1, -1, 1), groups=3)
    lo, hi = xp.amin(), xp.amax()
    xp = 0.05 + 0.9 * (xp - lo) / (hi - lo)
    return torch.round(xp * 255.0) / 255.0


def synthetic_batch(indices, H=512, W=768):
    """[len(indices), 3, H, W] fp32 on the CPU: images ``indices`` of the seeded set."""
    return torch.cat([synthetic_image(int(i), H, W) for i in indices])


def write_png_set(directory, n, H=512, W=768):
    """The same images as 8-bit PNG files for the ``-s '<glob>'`` CLI path (needs Pillow, like the reference's
    ``coder.write_image``, coder.py:36-48).  Returns the file names."""
    import os
    from PIL import Image
    os.makedirs(directory, exist_ok=True)
    names = []
    for i in range(n):
        arr = (synthetic_image(i, H, W)[0].permute(1, 2, 0) * 255.0).round().to(torch.uint8).numpy()
        name = os.path.join(directory, f"synthetic_{i:03d}.png")
        Image.fromarray(arr).save(name)
        names.append(name)
    return names
