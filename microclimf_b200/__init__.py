"""microclimf_b200 — B200-native grid solver behind the reference's `.Call` boundary.

Only what the hot path needs lives here:
  csrc/       hand-written sm_100a CUDA kernels + the C ABI (include/microclimf_b200.h)
  _abi.py     ctypes mirror of the C ABI structs
  _lib.py     loader for the built shared library (fails loudly: there is no CPU fallback)
  problem.py  host-side packing of the reference drivers' argument lists
  api.py      Python mirror of the reference's Rcpp-exported operators (runmicro1Cpp ... runbioclim4Cpp)
  bands.py    column-band sharding across ranks (one process per GPU)
  synth.py    seeded synthetic rasters / forcing for tests and benchmarks
"""
from ._abi import BIO_NAMES, OUT_NAMES  # noqa: F401
from .problem import GridProblem  # noqa: F401

__version__ = "0.1.0"
