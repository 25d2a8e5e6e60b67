"""rdis_b200 — B200-native (sm_100a CUDA, fp64) subspace solves and factor sweeps for RDIS.

The product is the C-ABI shared library `librdis_b200.so` (include/rdis_gpu.h) plus the C++
adapter in rdis_b200/host/ that keeps the reference's SubspaceOptimizer plugin surface.  The
Python modules here are thin ctypes plumbing for tests and bench.py; importing them without the
built library raises (no CPU fallback).
"""
from .capi import Context, Batch, ProblemSet, RdisGpuError, lib, LIB_PATH  # noqa: F401

lib()  # fail loudly at import time if the CUDA library is missing
