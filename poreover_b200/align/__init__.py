"""Drop-in for poreover.align (align.pyx): same function names, arguments and return values; the
dynamic programme runs on the GPU (poreover_b200/csrc/align.cu)."""
from .align import global_pair, global_pair_banded  # noqa: F401
