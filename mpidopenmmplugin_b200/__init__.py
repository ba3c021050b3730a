"""MPIDB200: B200-native (sm_100a) implementation of the MPIDForce hot path of andysim/MPIDOpenMMPlugin.

The product is the CUDA library `libmpidb200.so` (C ABI: include/mpidb200.h).  This package is the thin
Python host above it: `MPIDForce` mirrors the plugin's parameter object and `MPIDB200Kernel` mirrors the
`CalcMPIDForceKernel` contract, so tests read like the reference's own.  There is no CPU fallback: loading
fails loudly when the library is missing, and creating a kernel fails when no CUDA device is present."""
from .api import (MPIDForce, MPIDB200Kernel, MPIDB200Error, load_library, library_path)  # noqa: F401

__all__ = ["MPIDForce", "MPIDB200Kernel", "MPIDB200Error", "load_library", "library_path"]
