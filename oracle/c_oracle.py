"""ctypes binding of oracle/direct_convolve.c (TEST INFRASTRUCTURE ONLY; see that file's header)."""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_build", "libfc_oracle.so")
_lib = None


def build(force=False):
    """Compile the C restatement (and the reference into oracle/_ref when its sources are present)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(os.path.join(_HERE, "direct_convolve.c")):
        subprocess.check_call(["make", "-C", _HERE, os.path.join(_HERE, "_build", "libfc_oracle.so")],
                              stdout=subprocess.DEVNULL)


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = ctypes.CDLL(_SO)
        fp = ctypes.POINTER(ctypes.c_float)
        ip = ctypes.POINTER(ctypes.c_int)
        _lib.fc_oracle_convolve.argtypes = [fp, ip, fp, ip, fp, ip]
        _lib.fc_oracle_convolve.restype = None
        _lib.fc_oracle_convolve_omp.argtypes = [fp, ip, fp, ip, fp, ip, ctypes.c_int, ctypes.c_int]
        _lib.fc_oracle_convolve_omp.restype = ctypes.c_int
        _lib.fc_oracle_num_threads.restype = ctypes.c_int
    return _lib


def _fp(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def _ip(seq):
    return (ctypes.c_int * 3)(*[int(v) for v in seq])


def direct_convolve(image, kernel, offset, threads="single", z_range=None):
    """image/kernel are [z][y][x] float32 arrays.  Returns (result, threads_used)."""
    image = np.ascontiguousarray(image, dtype=np.float32)
    kernel = np.ascontiguousarray(kernel, dtype=np.float32)
    result = image.copy()          # the reference pre-fills with the padded image (test_fixtures.hpp:212-215)
    L = lib()
    if threads == "single":
        L.fc_oracle_convolve(_fp(image), _ip(image.shape), _fp(kernel), _ip(kernel.shape), _fp(result), _ip(offset))
        return result, 1
    z0, z1 = z_range if z_range is not None else (0, image.shape[0])
    n = L.fc_oracle_convolve_omp(_fp(image), _ip(image.shape), _fp(kernel), _ip(kernel.shape), _fp(result),
                                 _ip(offset), int(z0), int(z1))
    return result, n
