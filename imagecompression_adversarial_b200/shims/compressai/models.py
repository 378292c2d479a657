from imagecompression_adversarial_b200.models import (CompressionModel, FactorizedPrior,  # noqa: F401
                                                      MeanScaleHyperprior, ScaleHyperprior)
