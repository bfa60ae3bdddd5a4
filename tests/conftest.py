import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box via gpurun)")


def _gpu_unavailable_reason():
    """None when libaobake.so can create a context; otherwise why not.  A missing library is NOT a reason
    to skip — the gpu tests must then fail loudly (there is no CPU fallback to hide behind)."""
    lib_path = os.environ.get("AOBAKE_LIB") or os.path.join(ROOT, "optix_prime_baking_b200", "libaobake.so")
    if not os.path.exists(lib_path):
        return None
    try:
        lib = ctypes.CDLL(lib_path)
        lib.aobake_create.argtypes = [ctypes.c_void_p, ctypes.POINTER(ctypes.c_void_p)]
        lib.aobake_destroy.argtypes = [ctypes.c_void_p]
        lib.aobake_destroy.restype = None
        h = ctypes.c_void_p()
        rc = lib.aobake_create(None, ctypes.byref(h))
        if rc == 0:
            lib.aobake_destroy(h)
            return None
        if rc == 5:   # AOBAKE_ERR_NO_DEVICE
            return "no CUDA device on this host (aobake_create -> AOBAKE_ERR_NO_DEVICE)"
    except OSError:
        return None
    return None


def pytest_collection_modifyitems(config, items):
    if not any("gpu" in it.keywords for it in items):
        return
    reason = _gpu_unavailable_reason()
    if reason is None:
        return
    skip = pytest.mark.skip(reason=reason)
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)
