"""Loader of the in-tree CUDA shared library (the product).  No CPU fallback: if the library is
missing or a symbol cannot be resolved, importing/using the package fails loudly."""
import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libFourierConvolutionCUDALib.so")

# every symbol declared in include/convolution3Dfft.h and include/fcb200_ext.h
ABI_SYMBOLS = (
    "convolution3DfftCUDAInPlace", "convolution3DfftCUDAInPlaceSaveMemory",
    "convolution3DfftCUDA", "convolution3DfftCUDA_test",
    "selectDeviceWithHighestComputeCapability", "getCUDAcomputeCapabilityMinorVersion",
    "getCUDAcomputeCapabilityMajorVersion", "getNumDevicesCUDA", "getNameDeviceCUDA",
    "getMemDeviceCUDA", "cuda_version", "gpu_mem_needed_mb",
    "fcb200_last_error", "fcb200_set_error_mode", "fcb200_free_result",
)
EXT_SYMBOLS = (
    "fcb200_convolve_device_async", "fcb200_convolve_device_async_savememory", "fcb200_plan_radices", "fcb200_plan_tables",
    "fcb200_plan_radices_style", "fcb200_plan_tables_style",
    "fcb200_spectrum_pitch", "fcb200_workspace_bytes", "fcb200_psf_active_rows",
    "fcb200_debug_rfft3", "fcb200_debug_irfft3", "fcb200_debug_psf_spectrum",
    "fcb200_slab_xy_forward", "fcb200_slab_z_fused", "fcb200_slab_yx_inverse", "fcb200_slab_psf_scratch_elems",
    "fcb200_slab_psf", "fcb200_slab_xy_forward_peer", "fcb200_slab_z_fused_peer", "fcb200_device_malloc",
    "fcb200_device_free", "fcb200_ipc_get_handle", "fcb200_ipc_open_handle", "fcb200_ipc_close_handle", "fcb200_convolve_batch",
    "fcb200_release", "fcb200_launch_count", "fcb200_profile_enable", "fcb200_profile_read",
    "fcb200_padded_extents", "fcb200_convolve_padded", "fcb200_convolve_padded_device_async",
    "fcb200_convolve_batch_padded",
    "fcb200_convolve_slab", "fcb200_convolve_slab_device", "fcb200_slab_last_timing", "fcb200_slab_devices",
    "fcb200_convolve_batch_multi", "fcb200_psf_window_planes", "fcb200_plan_rader",
)


def build(verbose=False):
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    cmd = ["make", "-C", os.path.join(_HERE, "csrc"), "-j4"]
    subprocess.check_call(cmd, stdout=None if verbose else subprocess.DEVNULL)
    return LIB_PATH


_lib = None


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(this package has no CPU fallback)")
    lib = ctypes.CDLL(LIB_PATH)   # RTLD_LOCAL: may coexist with the reference build of the same name
    fp = ctypes.POINTER(ctypes.c_float)
    ip = ctypes.POINTER(ctypes.c_int)
    vp = ctypes.c_void_p
    i = ctypes.c_int
    sig = {
        "convolution3DfftCUDAInPlace": (None, [vp, ip, vp, ip, i]),
        "convolution3DfftCUDAInPlaceSaveMemory": (None, [vp, ip, vp, ip, i]),
        "convolution3DfftCUDA": (vp, [vp, ip, vp, ip, i]),
        "convolution3DfftCUDA_test": (vp, [vp, ip, vp, i]),
        "selectDeviceWithHighestComputeCapability": (i, []),
        "getCUDAcomputeCapabilityMinorVersion": (i, [i]),
        "getCUDAcomputeCapabilityMajorVersion": (i, [i]),
        "getNumDevicesCUDA": (i, []),
        "getNameDeviceCUDA": (None, [i, ctypes.c_char_p]),
        "getMemDeviceCUDA": (ctypes.c_longlong, [i]),
        "cuda_version": (i, []),
        "gpu_mem_needed_mb": (i, [ip, i]),
        "fcb200_last_error": (ctypes.c_char_p, []),
        "fcb200_free_result": (None, [vp]),
        "fcb200_set_error_mode": (None, [i]),
        "fcb200_convolve_device_async": (None, [vp, ip, vp, ip, i, vp]),
        "fcb200_convolve_device_async_savememory": (None, [vp, ip, vp, ip, i, vp]),
        "fcb200_plan_radices": (i, [i, ip, ip]),
        "fcb200_plan_tables": (None, [i, ip, ip, fp]),
        "fcb200_plan_radices_style": (i, [i, i, ip, ip]),
        "fcb200_plan_tables_style": (None, [i, i, ip, ip, fp]),
        "fcb200_spectrum_pitch": (i, [i]),
        "fcb200_workspace_bytes": (ctypes.c_longlong, [ip]),
        "fcb200_psf_active_rows": (ctypes.c_longlong, [ip, ip, ip, ctypes.c_longlong]),
        "fcb200_debug_rfft3": (None, [fp, ip, fp, i, i]),
        "fcb200_debug_irfft3": (None, [fp, ip, fp, i]),
        "fcb200_debug_psf_spectrum": (None, [fp, ip, ip, fp, i]),
        "fcb200_slab_xy_forward": (None, [vp, vp, vp, ip, i, i, i, i, vp]),
        "fcb200_slab_z_fused": (None, [vp, vp, ip, i, i, vp]),
        "fcb200_slab_yx_inverse": (None, [vp, vp, vp, ip, i, i, i, i, vp]),
        "fcb200_slab_psf_scratch_elems": (ctypes.c_longlong, [ip, ip, i]),
        "fcb200_slab_psf": (None, [vp, ip, ip, i, i, vp, vp, i, vp]),
        "fcb200_slab_xy_forward_peer": (None, [vp, vp, vp, ip, i, i, i, i, i, vp]),
        "fcb200_slab_z_fused_peer": (None, [vp, vp, vp, ip, i, i, i, i, vp]),
        "fcb200_device_malloc": (vp, [ctypes.c_longlong, i]),
        "fcb200_device_free": (None, [vp, i]),
        "fcb200_ipc_get_handle": (None, [vp, ctypes.c_char_p]),
        "fcb200_ipc_open_handle": (vp, [ctypes.c_char_p, i]),
        "fcb200_ipc_close_handle": (None, [vp, i]),
        "fcb200_convolve_batch": (None, [vp, i, ip, vp, ip, i]),
        "fcb200_padded_extents": (None, [ip, ip, i, ip]),
        "fcb200_convolve_padded": (None, [vp, ip, vp, ip, i, i, i]),
        "fcb200_convolve_batch_padded": (None, [vp, i, ip, vp, ip, i, i, i]),
        "fcb200_convolve_padded_device_async": (None, [vp, ip, vp, ip, i, i, i, vp]),
        "fcb200_convolve_slab": (None, [vp, ip, vp, ip, ip, i]),
        "fcb200_convolve_slab_device": (None, [vp, ip, vp, ip, ip, i]),
        "fcb200_slab_last_timing": (i, [ip, ip, i, fp, i]),
        "fcb200_slab_devices": (i, [ip, i, ip, i]),
        "fcb200_convolve_batch_multi": (None, [vp, i, ip, vp, ip, ip, i, ip]),
        "fcb200_psf_window_planes": (i, [ip, ip, i]),
        "fcb200_plan_rader": (i, [i, ip, ip, ip, fp, fp]),
        "fcb200_release": (None, []),
        "fcb200_launch_count": (ctypes.c_longlong, []),
        "fcb200_profile_enable": (None, [i]),
        "fcb200_profile_read": (i, [fp, ctypes.POINTER(ctypes.c_longlong), i]),
    }
    for name, (res, args) in sig.items():
        fn = getattr(lib, name)          # AttributeError if the symbol is not exported
        fn.restype = res
        fn.argtypes = args
    lib.fcb200_set_error_mode(1)     # ctypes cannot catch C++ exceptions; api.py raises from last_error
    _lib = lib
    return lib
