"""softrender_b200 -- B200-native implementation of the draw hot path of novacrazy/rust-softrender.

Layout:
  csrc/       hand-written CUDA kernels (sm_100a) + the C ABI (include/softrender_b200.h)
  _abi.py     ctypes binding of libsoftrender_b200.so (fails loudly when the CUDA library is missing)
  pipeline.py host-side mirror of the reference's builder API (Pipeline -> VertexShader -> GeometryShader
              -> FragmentShader), same verbs and argument meaning as src/pipeline/ in the reference
  scenes.py   scene/mesh/uniform construction for the benchmark configs and tests
"""
from .constants import *  # noqa: F401,F403
from . import scenes  # noqa: F401
