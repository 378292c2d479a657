"""compressai.datasets.ImageFolder as train.py:114-123 uses it: ``root/<split>/`` holds the image files, every item is
the RGB image passed through ``transform``.  Host-side data loading (PIL), not part of the attack hot path."""
import os

from PIL import Image

_EXT = (".png", ".jpg", ".jpeg", ".bmp", ".ppm", ".tif", ".tiff", ".webp")


class ImageFolder:
    def __init__(self, root, transform=None, split="train"):
        splitdir = os.path.join(str(root), split)
        if not os.path.isdir(splitdir):
            raise RuntimeError(f'Invalid directory "{root}"')
        self.samples = sorted(os.path.join(splitdir, f) for f in os.listdir(splitdir)
                              if os.path.isfile(os.path.join(splitdir, f)) and f.lower().endswith(_EXT))
        self.transform = transform

    def __getitem__(self, index):
        img = Image.open(self.samples[index]).convert("RGB")
        return self.transform(img) if self.transform else img

    def __len__(self):
        return len(self.samples)
