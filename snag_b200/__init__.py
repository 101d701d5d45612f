"""snag_b200 — B200-native (sm_100a) implementation of the SNAG_MMEA data-parallel hot path:
Gauss modality noise masking, ICL/IAL in-batch contrastive losses, alignment evaluation.

The compute lives in libsnag_b200.so (hand-written CUDA behind a C ABI, see include/snag_b200.h);
this package is the host-side mirror of the reference's Python interface for that path.
"""
from ._lib import KT, LIB_PATH, SnagError, load  # noqa: F401

__version__ = "0.1.0"
