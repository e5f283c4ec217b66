"""zerfoo_b200 -- B200-native (sm_100a) decode hot path of zerfoo/zerfoo.

The product is ``zerfoo_b200/lib/libkernels.so``: a C-ABI shared library that
exports (a) every launcher of the reference's ``libkernels.so`` under the same
names (include/zerfoo_kernels.h) and (b) the ``zb_*`` engine and B200 entry
points (include/zb200.h).  The Python modules here are only the host-side
mirror of the reference's Go wrappers, used by tests and bench.py:

  * ``zerfoo_b200.lib``      ctypes loader (fails loudly when the CUDA library is missing)
  * ``zerfoo_b200.kernels``  mirrors internal/cuda/kernels/*_purego.go (GemmQ4F32, GemvQ4KF32, ...)
  * ``zerfoo_b200.engine``   mirrors generate.Generator / InferenceSession (Generate, decode step)
  * ``zerfoo_b200.gguf``     GGUF writer/reader + block quantizers for synthetic models
"""
__all__ = ["lib", "kernels", "engine", "gguf"]
