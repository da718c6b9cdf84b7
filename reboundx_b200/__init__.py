"""reboundx_b200 -- B200-native implementation of REBOUNDx's `rebx_additional_forces` hot path.

Layout:
  csrc/          CUDA kernels (sm_100a) + the C-ABI drop-in library sources
  lib/           built artefacts (libreboundx_b200.so), git-ignored
  rebound_shim/  stand-in for the absent REBOUND host library
  host.py        ctypes mirror of the reference's Python interface (Extras/Force/Params)
  device.py      resident-state API over torch tensors (device memory, streams, NCCL plumbing)
  workloads.py   seeded synthetic ensembles for the BASELINE.json configs
"""
from ._lib import load, load_shim, NativeLibraryMissing, LIB_PATH  # noqa: F401

__version__ = "0.1.0"
