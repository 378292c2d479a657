"""Arithmetic mode of the tensor-path contractions.

``"tf32"`` (default, the speed mode): operands rounded to TF32 once where they are produced, one tcgen05 kind::tf32
MMA per product, fp32 accumulation in TMEM -- the arithmetic of the reference's own GPU path (cuDNN with TF32 allowed).
``"3xtf32"`` (the parity mode): every operand is split into two TF32 terms and every product is three MMAs
(hi*hi + lo*hi + hi*lo) on the same kernels (csrc/icadv_split.cu), GDN / IGDN unfused so the normalisation GEMM is
split too; results track the fp32 reference to ~1e-6 per contraction, which is what the literal 1e-3 per-step loss
tolerance of the north star needs (the loop integrates the sign of every input-gradient element).

Select with the environment variable ``ICADV_PRECISION`` or ``precision.set(mode)`` / ``with precision.use(mode)``
before engines / programs are built (they read the mode at construction).
"""
import os

MODES = ("tf32", "3xtf32")
_mode = [os.environ.get("ICADV_PRECISION", "tf32")]
if _mode[0] not in MODES:
    raise ValueError(f"ICADV_PRECISION must be one of {MODES}, got {_mode[0]!r}")


def get():
    return _mode[0]


def split():
    """True in the 3xTF32 parity mode."""
    return _mode[0] == "3xtf32"


def set(mode):  # noqa: A001
    if mode not in MODES:
        raise ValueError(f"precision mode must be one of {MODES}, got {mode!r}")
    _mode[0] = mode


class use:
    def __init__(self, mode):
        self.mode = mode

    def __enter__(self):
        self.prev = get()
        set(self.mode)
        return self

    def __exit__(self, *exc):
        set(self.prev)
