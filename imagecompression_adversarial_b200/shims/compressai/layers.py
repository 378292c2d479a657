from imagecompression_adversarial_b200.models import GDN  # noqa: F401
