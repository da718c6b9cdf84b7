"""Loader for the native libraries of this package.

``load()`` returns the drop-in C-ABI library ``lib/libreboundx_b200.so`` (REBOUNDx force
path API + CUDA kernels).  It raises loudly when the library has not been built: there is
no Python or CPU fallback for the force path.
"""
import ctypes
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("REBX_B200_LIB") or os.path.join(_HERE, "lib", "libreboundx_b200.so")
SHIM_PATH = os.path.join(_HERE, "rebound_shim", "librebound_shim.so")

_cache = {}


class NativeLibraryMissing(RuntimeError):
    pass


def load_shim():
    """The stand-in REBOUND host library (simulation/particle storage, messages)."""
    if "shim" not in _cache:
        if not os.path.exists(SHIM_PATH):
            raise NativeLibraryMissing(
                f"{SHIM_PATH} not built; run `python -c 'import __graft_entry__ as g; g.build()'`")
        _cache["shim"] = ctypes.CDLL(SHIM_PATH, mode=ctypes.RTLD_GLOBAL)
    return _cache["shim"]


def load(path=None):
    """The REBOUNDx-compatible library.  `path` selects another build with the same ABI
    (tests pass the reference-compiled oracle library to drive both through one harness)."""
    path = os.path.abspath(path or LIB_PATH)
    if path not in _cache:
        load_shim()
        if not os.path.exists(path):
            raise NativeLibraryMissing(
                f"{path} not built; the force path has no fallback. "
                "Run `python -c 'import __graft_entry__ as g; g.build()'`")
        _cache[path] = ctypes.CDLL(path, mode=ctypes.RTLD_LOCAL)
    return _cache[path]
