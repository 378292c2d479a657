from imagecompression_adversarial_b200.models import (CompressionModel, FactorizedPrior,  # noqa: F401
                                                      ScaleHyperprior)
from imagecompression_adversarial_b200._lib import IcadvError


class MeanScaleHyperprior(CompressionModel):
    """Only referenced as the base class of the reference's `debug` model (anchors/model.py:9)."""

    def __init__(self, N, M, **kwargs):
        raise IcadvError("MeanScaleHyperprior (the reference's `debug` model) is not built on the sm_100a path")
