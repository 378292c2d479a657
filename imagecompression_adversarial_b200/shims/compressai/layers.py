from imagecompression_adversarial_b200.models import GDN  # noqa: F401
from imagecompression_adversarial_b200.models import (AttentionBlock, MaskedConv2d, ResidualBlock,  # noqa: F401
                                                      ResidualBlockUpsample, ResidualBlockWithStride, conv3x3,
                                                      subpel_conv3x3)
