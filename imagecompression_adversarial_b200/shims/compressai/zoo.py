from imagecompression_adversarial_b200.models import (bmshj2018_factorized, bmshj2018_hyperprior,  # noqa: F401
                                                      cheng2020_anchor, cheng2020_attn, mbt2018)
