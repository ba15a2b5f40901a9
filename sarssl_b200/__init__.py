"""sarssl_b200 - B200-native (sm_100a) implementation of the SAR-SSL pre-training hot path.

Host side: Python mirroring the reference's module interface (code/model.py, code/learner.py,
code/common/utils_module.py).  Device side: hand-written CUDA in libsarssl_b200.so behind a C ABI
(include/sarssl_b200.h).  No CPU fallback, no Triton, no dispatch layer."""
from ._lib import SarsslError, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
