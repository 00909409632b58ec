"""soundkit_b200 — Blackwell-native (sm_100a) implementation of soundkit's canonical audio normalization path.

Layout:
  csrc/       hand-written CUDA kernels + the C ABI (include/soundkit_cuda.h) -> libsoundkit_cuda.so
  capi.py     ctypes binding of that ABI (what tests and bench.py call)
  csrc/host/  C++ host-side mirror of the reference's Rust interface for this path (same names, argument
              meaning and error strings) on top of the C ABI; `_host` is its pybind11 view for pytest
  shard.py    multi-GPU partition of a batch of independent streams (no data-path collective)

There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
