"""soundkit_b200 — Blackwell-native (sm_100a) implementation of soundkit's canonical audio normalization path.

Layout:
  csrc/       hand-written CUDA kernels + the C ABI (include/soundkit_cuda.h) -> libsoundkit_cuda.so
  capi.py     ctypes binding of that ABI (what tests and bench.py call)
  host.py     host-side mirror of the reference's Rust interface for this path (same names, argument
              meaning and error strings), on top of the C ABI

There is no CPU fallback anywhere in this package.
"""
from . import capi  # noqa: F401

__all__ = ["capi"]
