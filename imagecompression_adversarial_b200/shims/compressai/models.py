from imagecompression_adversarial_b200.models import (CompressionModel, FactorizedPrior,  # noqa: F401
                                                      MeanScaleHyperprior, ScaleHyperprior)
from imagecompression_adversarial_b200.models import (Cheng2020Anchor, Cheng2020Attention,  # noqa: F401
                                                      JointAutoregressiveHierarchicalPriors)
