"""
mimeo_b200 -- B200-native (sm_100a) implementation of mimeo's alignment-to-annotation hot path.

Only what that path needs lives here:
  csrc/        hand-written CUDA kernels + the C ABI (include/mimeo_b200.h) -> libmimeo_b200.so
  _lib.py      ctypes binding of the C ABI (fails loudly when the library or a GPU is missing)
  coverage.py  host mirror of the reference's coverage/threshold/merge script stage
There is no CPU fallback anywhere in this package; the CPU oracle lives in /oracle and is test-only.
"""
__version__ = '0.1.0'
