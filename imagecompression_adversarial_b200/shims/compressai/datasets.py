class ImageFolder:  # train.py:114 only; data loading is out of scope of the hot path
    def __init__(self, *a, **k):
        raise NotImplementedError("compressai.datasets.ImageFolder is outside the attack hot path")
