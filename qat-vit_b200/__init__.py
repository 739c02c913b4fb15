"""qatvit_b200 -- B200-native (sm_100a) kernels for the QAT-distillation hot path of bdina9/qat-vit.

Host code is Python/PyTorch (as in the reference); all arithmetic of the hot path runs in
hand-written CUDA behind the C-ABI declared in include/qatvit_b200.h (libqatvit_b200.so).
"""
from . import _lib  # noqa: F401  (raises if the shared library is missing -- no fallback)

__version__ = "0.1.0"


def build(verbose: bool = False) -> str:
    return _lib.build(verbose)
