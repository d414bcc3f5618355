"""ctypes binding of the reference's own cuFFT build (oracle/_ref, built by oracle/Makefile from the
sources under /root/reference).  Checker only: used by the -m gpu parity tests, smoke() and
bench.py --impl reference.  Needs a GPU (links libcuda.so.1)."""
import ctypes
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF_SO = os.path.join(ROOT, "oracle", "_ref", "libFourierConvolutionCUDALib.so")
_ref = None


def available():
    return os.path.exists(REF_SO)


def load():
    global _ref
    if _ref is None:
        if not available():
            raise RuntimeError("oracle/_ref/libFourierConvolutionCUDALib.so missing: run `make -C oracle ref` where "
                               "/root/reference exists")
        for dep in ("/usr/local/cuda/lib64/libcufft.so.11",):
            if os.path.exists(dep):
                ctypes.CDLL(dep, mode=ctypes.RTLD_GLOBAL)
        lib = ctypes.CDLL(REF_SO)      # RTLD_LOCAL: same symbol names as the product library
        ip = ctypes.POINTER(ctypes.c_int)
        lib.convolution3DfftCUDAInPlace.argtypes = [ctypes.c_void_p, ip, ctypes.c_void_p, ip, ctypes.c_int]
        lib.convolution3DfftCUDAInPlace.restype = None
        lib.selectDeviceWithHighestComputeCapability.restype = ctypes.c_int
        lib.getNumDevicesCUDA.restype = ctypes.c_int
        lib.getMemDeviceCUDA.restype = ctypes.c_longlong
        lib.getMemDeviceCUDA.argtypes = [ctypes.c_int]
        lib.getNameDeviceCUDA.argtypes = [ctypes.c_int, ctypes.c_char_p]
        lib.cuda_version.restype = ctypes.c_int
        _ref = lib
    return _ref


def convolve_inplace(im, imDim, kernel, kernelDim, dev=0):
    """runs the REFERENCE convolution3DfftCUDAInPlace on copies; returns the flat result"""
    lib = load()
    out = np.ascontiguousarray(im, np.float32).reshape(-1).copy()
    k = np.ascontiguousarray(kernel, np.float32).reshape(-1).copy()
    idim = (ctypes.c_int * 3)(*[int(v) for v in imDim])
    kdim = (ctypes.c_int * 3)(*[int(v) for v in kernelDim])
    lib.convolution3DfftCUDAInPlace(ctypes.c_void_p(out.ctypes.data), idim, ctypes.c_void_p(k.ctypes.data), kdim,
                                    int(dev))
    return out
