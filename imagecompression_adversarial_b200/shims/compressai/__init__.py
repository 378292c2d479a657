"""`compressai`-compatible import surface backed by the sm_100a kernels (only the names the reference's hot
path imports: anchors/model.py:3-5, anchors/balle.py:7-8, train.py:15).  Used only when the real package is
not installed (see imagecompression_adversarial_b200/launch.py)."""
__version__ = "0.0+icadv_b200"
