"""Host-side mirror of the reference's C interface (names, argument meaning and error behaviour of
/root/reference/src/convolution3Dfft.h), implemented by calling the in-tree CUDA library through its
C ABI.  numpy arrays stand in for the host `float*` / `int*` arguments; torch CUDA tensors (or raw
device addresses) may be passed where the library accepts device pointers.
"""
import ctypes

import numpy as np

from . import _lib


def _ints(seq):
    return (ctypes.c_int * len(seq))(*[int(v) for v in seq])


def _ptr(a):
    """address of a numpy array, a torch tensor, or a raw integer address"""
    if isinstance(a, np.ndarray):
        if a.dtype != np.float32 or not a.flags["C_CONTIGUOUS"]:
            raise TypeError("arrays handed to the C ABI must be C-contiguous float32")
        return ctypes.c_void_p(a.ctypes.data)
    if hasattr(a, "data_ptr"):
        if str(a.dtype) != "torch.float32" or not a.is_contiguous():
            raise TypeError("tensors handed to the C ABI must be contiguous float32")
        return ctypes.c_void_p(a.data_ptr())
    return ctypes.c_void_p(int(a))


class FourierConvolutionError(RuntimeError):
    pass


class _Checked:
    """Proxy of the ctypes library: every call is followed by a look at fcb200_last_error(), so a
    failure inside the CUDA library surfaces as FourierConvolutionError (the Python stand-in for the
    std::runtime_error the C ABI throws at C++ callers)."""

    # entry points that never fail and therefore do not reset the per-thread error string
    _UNGUARDED = {"fcb200_profile_enable", "fcb200_profile_read", "cuda_version", "fcb200_free_result", "fcb200_spectrum_pitch", "fcb200_launch_count",
                  "fcb200_release", "fcb200_workspace_bytes", "getCUDAcomputeCapabilityMajorVersion",
                  "getCUDAcomputeCapabilityMinorVersion"}

    def __init__(self, lib):
        self._lib = lib

    def __getattr__(self, name):
        fn = getattr(self._lib, name)
        lib = self._lib
        if name in self._UNGUARDED:
            return fn

        def call(*args):
            res = fn(*args)
            err = lib.fcb200_last_error()
            if err:
                raise FourierConvolutionError(err.decode())
            return res

        return call


def _load():
    return _Checked(_lib.load())


def convolution3DfftCUDAInPlace(im, imDim, kernel, kernelDim, devCUDA):
    """reference: src/convolution3Dfft.h:56.  `im` is overwritten; imDim[0] is its fastest axis."""
    _load().convolution3DfftCUDAInPlace(_ptr(im), _ints(imDim), _ptr(kernel), _ints(kernelDim), int(devCUDA))


def convolution3DfftCUDAInPlaceSaveMemory(im, imDim, kernel, kernelDim, devCUDA):
    """reference: src/convolution3Dfft.h:58-64."""
    _load().convolution3DfftCUDAInPlaceSaveMemory(_ptr(im), _ints(imDim), _ptr(kernel), _ints(kernelDim),
                                                      int(devCUDA))


def convolve_batch(ims, imDim, kernel, kernelDim, devCUDA):
    """extension (include/fcb200_ext.h: fcb200_convolve_batch): every array of `ims` (numpy / torch, host or
    device, one kind) is convolved in place with the same PSF; transfers and convolutions are pipelined."""
    ptrs = (ctypes.c_void_p * len(ims))(*[_ptr(im) for im in ims])
    _load().fcb200_convolve_batch(ptrs, len(ims), _ints(imDim), _ptr(kernel), _ints(kernelDim), int(devCUDA))


def convolve_batch_multi(ims, imDim, kernel, kernelDim, devs):
    """extension (fcb200_convolve_batch_multi): host blocks dealt over several devices from one shared queue;
    returns how many blocks each device took"""
    ptrs = (ctypes.c_void_p * len(ims))(*[_ptr(im) for im in ims])
    taken = (ctypes.c_int * len(devs))()
    _load().fcb200_convolve_batch_multi(ptrs, len(ims), _ints(imDim), _ptr(kernel), _ints(kernelDim), _ints(devs),
                                            len(devs), taken)
    return list(taken)


def convolve_slab(im, imDim, kernel, kernelDim, devs):
    """extension (fcb200_convolve_slab): ONE host volume convolved in place, cut in z slabs over `devs`"""
    _load().fcb200_convolve_slab(_ptr(im), _ints(imDim), _ptr(kernel), _ints(kernelDim), _ints(devs), len(devs))


def convolve_slab_device(slabs, imDim, kernel, kernelDim, devs):
    """extension (fcb200_convolve_slab_device): slabs[r] = rank r's z slab, resident on devs[r]"""
    ptrs = (ctypes.c_void_p * len(slabs))(*[_ptr(s) for s in slabs])
    _load().fcb200_convolve_slab_device(ptrs, _ints(imDim), _ptr(kernel), _ints(kernelDim), _ints(devs), len(devs))


def slab_last_timing(imDim, devs):
    """-> per rank (forward, fused z, inverse, total) device milliseconds of the most recent slab call"""
    ms = (ctypes.c_float * (4 * len(devs)))()
    n = _load().fcb200_slab_last_timing(_ints(imDim), _ints(devs), len(devs), ms, 4 * len(devs))
    return [tuple(ms[4 * r:4 * r + 4]) for r in range(n)]


def slab_devices(imDim, devCUDA=0):
    out = (ctypes.c_int * 64)()
    n = _load().fcb200_slab_devices(_ints(imDim), int(devCUDA), out, 64)
    return list(out[:n])


def psf_window_planes(imDim, kernelDim, devCUDA=0):
    """extension (fcb200_psf_window_planes): 16 / 32 / 64 when the fused z pass derives the PSF spectrum on the fly
    from that many PSF planes, 0 when the image-sized PSF spectrum is materialised"""
    return _load().fcb200_psf_window_planes(_ints(imDim), _ints(kernelDim), int(devCUDA))


def slab_partition(imDim, world):
    """(nzp, nyl, [planes of rank r]) of the slab decomposition used by convolve_slab*"""
    d0, d1, d2 = (int(v) for v in imDim)
    nzp, nyl = -(-d2 // world), -(-d1 // world)
    return nzp, nyl, [max(0, min(nzp, d2 - r * nzp)) for r in range(world)]


def convolution3DfftCUDA(im, imDim, kernel, kernelDim, devCUDA):
    """reference: src/convolution3Dfft.h:41-45 (legacy: imDim[2] fastest).  Returns a new array."""
    lib = _load()
    n = int(np.prod(imDim))
    p = lib.convolution3DfftCUDA(_ptr(im), _ints(imDim), _ptr(kernel), _ints(kernelDim), int(devCUDA))
    out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n,)).copy()
    lib.fcb200_free_result(p)
    return out


def convolution3DfftCUDA_test(im, imDim, kernel, devCUDA):
    """reference: src/convolution3Dfft.h:29-32 (kernel already image-sized, no shift)."""
    lib = _load()
    n = int(np.prod(imDim))
    p = lib.convolution3DfftCUDA_test(_ptr(im), _ints(imDim), _ptr(kernel), int(devCUDA))
    out = np.ctypeslib.as_array(ctypes.cast(p, ctypes.POINTER(ctypes.c_float)), shape=(n,)).copy()
    lib.fcb200_free_result(p)
    return out


def selectDeviceWithHighestComputeCapability():
    return _load().selectDeviceWithHighestComputeCapability()


def getCUDAcomputeCapabilityMajorVersion(devCUDA):
    return _load().getCUDAcomputeCapabilityMajorVersion(int(devCUDA))


def getCUDAcomputeCapabilityMinorVersion(devCUDA):
    return _load().getCUDAcomputeCapabilityMinorVersion(int(devCUDA))


def getNumDevicesCUDA():
    return _load().getNumDevicesCUDA()


def getNameDeviceCUDA(devCUDA):
    buf = ctypes.create_string_buffer(256)       # the ABI writes exactly 256 bytes
    _load().getNameDeviceCUDA(int(devCUDA), buf)
    return buf.value.decode()


def getMemDeviceCUDA(devCUDA):
    return _load().getMemDeviceCUDA(int(devCUDA))


def cuda_version():
    return _load().cuda_version()


def gpu_mem_needed_mb(shape):
    return _load().gpu_mem_needed_mb(_ints(shape), len(shape))


# ---- extensions (include/fcb200_ext.h) ------------------------------------------------------
def convolve_device_async(im_dev, imDim, kernel_dev, kernelDim, devCUDA, stream=0):
    _load().fcb200_convolve_device_async(_ptr(im_dev), _ints(imDim), _ptr(kernel_dev), _ints(kernelDim),
                                             int(devCUDA), ctypes.c_void_p(int(stream)))


def convolve_device_async_savememory(im_dev, imDim, kernel_dev, kernelDim, devCUDA, stream=0):
    _load().fcb200_convolve_device_async_savememory(_ptr(im_dev), _ints(imDim), _ptr(kernel_dev), _ints(kernelDim),
                                                        int(devCUDA), ctypes.c_void_p(int(stream)))


PAD_ZERO, PAD_MIRROR = 0, 1            # mode
PAD_EXACT, PAD_SMOOTH = 0, 1           # policy


def padded_extents(imDim, kernelDim, policy=PAD_SMOOTH):
    """extension (fcb200_padded_extents): grid the padded entry points convolve on.  policy 0 is the
    reference's zero_padd (tests/padd_utils.h:12-24,99-108), 1 rounds up to 7-smooth sizes."""
    out = (ctypes.c_int * 3)()
    _load().fcb200_padded_extents(_ints(imDim), _ints(kernelDim), int(policy), out)
    return tuple(out)


def convolve_padded(im, imDim, kernel, kernelDim, devCUDA, mode=PAD_ZERO, policy=PAD_SMOOTH):
    """extension (fcb200_convolve_padded): `im` is the UNPADDED volume (host or device), overwritten with the
    cropped result of convolution3DfftCUDAInPlace on the padded grid (reference callers pad on the host:
    tests/padd_utils.h:157-171)."""
    _load().fcb200_convolve_padded(_ptr(im), _ints(imDim), _ptr(kernel), _ints(kernelDim), int(mode), int(policy),
                                       int(devCUDA))


def convolve_batch_padded(ims, imDim, kernel, kernelDim, devCUDA, mode=PAD_ZERO, policy=PAD_SMOOTH):
    """extension (fcb200_convolve_batch_padded): convolve_batch with the padding done in the library"""
    ptrs = (ctypes.c_void_p * len(ims))(*[_ptr(im) for im in ims])
    _load().fcb200_convolve_batch_padded(ptrs, len(ims), _ints(imDim), _ptr(kernel), _ints(kernelDim), int(mode),
                                             int(policy), int(devCUDA))


def convolve_padded_device_async(im_dev, imDim, kernel_dev, kernelDim, devCUDA, mode=PAD_ZERO, policy=PAD_SMOOTH,
                                 stream=0):
    _load().fcb200_convolve_padded_device_async(_ptr(im_dev), _ints(imDim), _ptr(kernel_dev), _ints(kernelDim),
                                                    int(mode), int(policy), int(devCUDA), ctypes.c_void_p(int(stream)))


def plan_radices(L, style=0):
    r = (ctypes.c_int * 16)()
    g = ctypes.c_int(0)
    n = _load().fcb200_plan_radices_style(int(L), int(style), r, ctypes.byref(g))
    return list(r[:n]), bool(g.value)


def plan_tables(L, style=0):
    rev = np.zeros(L, np.int32)
    pos = np.zeros(L, np.int32)
    tw = np.zeros(2 * L, np.float32)
    _load().fcb200_plan_tables_style(int(L), int(style), rev.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                         pos.ctypes.data_as(ctypes.POINTER(ctypes.c_int)),
                                         tw.ctypes.data_as(ctypes.POINTER(ctypes.c_float)))
    return rev, pos, tw[0::2] + 1j * tw[1::2]


def plan_rader(p):
    """extension (fcb200_plan_rader): Rader tables of the prime p, or None when p - 1 is not smooth enough"""
    n = int(p) - 1
    rad = (ctypes.c_int * 8)()
    perm, iperm = np.zeros(n, np.int32), np.zeros(n, np.int32)
    bf, bi = np.zeros(2 * n, np.float32), np.zeros(2 * n, np.float32)
    ip = ctypes.POINTER(ctypes.c_int)
    ns = _load().fcb200_plan_rader(int(p), rad, perm.ctypes.data_as(ip), iperm.ctypes.data_as(ip), _f(bf), _f(bi))
    if ns == 0:
        return None
    return {"radices": list(rad[:ns]), "perm": perm, "iperm": iperm, "bf": bf[0::2] + 1j * bf[1::2], "bi": bi[0::2] + 1j * bi[1::2]}


def spectrum_pitch(nx):
    return _load().fcb200_spectrum_pitch(int(nx))


def workspace_bytes(imDim):
    return _load().fcb200_workspace_bytes(_ints(imDim))


def psf_active_rows(imDim, kernelDim):
    lib = _load()
    n = lib.fcb200_psf_active_rows(_ints(imDim), _ints(kernelDim), None, 0)
    rows = np.zeros(n, np.int32)
    lib.fcb200_psf_active_rows(_ints(imDim), _ints(kernelDim), rows.ctypes.data_as(ctypes.POINTER(ctypes.c_int)), n)
    return rows


def _f(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_float))


def debug_rfft3(im, imDim, passes=3, devCUDA=0):
    d0, d1, d2 = (int(v) for v in imDim)
    im = np.ascontiguousarray(im, np.float32).reshape(-1)
    out = np.zeros((d2, d1, d0 // 2 + 1, 2), np.float32)
    _load().fcb200_debug_rfft3(_f(im), _ints(imDim), _f(out), int(passes), int(devCUDA))
    return out[..., 0] + 1j * out[..., 1]


def debug_irfft3(spec, imDim, devCUDA=0):
    d0, d1, d2 = (int(v) for v in imDim)
    s = np.zeros((d2, d1, d0 // 2 + 1, 2), np.float32)
    s[..., 0] = spec.real
    s[..., 1] = spec.imag
    out = np.zeros(d0 * d1 * d2, np.float32)
    _load().fcb200_debug_irfft3(_f(s), _ints(imDim), _f(out), int(devCUDA))
    return out.reshape(d2, d1, d0)


def debug_psf_spectrum(kernel, kernelDim, imDim, devCUDA=0):
    d0, d1, d2 = (int(v) for v in imDim)
    kernel = np.ascontiguousarray(kernel, np.float32).reshape(-1)
    out = np.zeros((d2, d1, d0 // 2 + 1, 2), np.float32)
    _load().fcb200_debug_psf_spectrum(_f(kernel), _ints(kernelDim), _ints(imDim), _f(out), int(devCUDA))
    return out[..., 0] + 1j * out[..., 1]


def release():
    _load().fcb200_release()


def launch_count():
    return _load().fcb200_launch_count()


PASS_NAMES = ("psf_clear", "psf_x", "psf_y", "psf_z", "x_fwd", "y_fwd", "z_fused", "y_inv", "x_inv")


def profile_enable(on=True):
    _load().fcb200_profile_enable(1 if on else 0)


def profile_read():
    """-> {pass name: (total ms, launches)} since the last read (waits for the recorded events)"""
    n = len(PASS_NAMES)
    ms = (ctypes.c_float * n)()
    cnt = (ctypes.c_longlong * n)()
    _load().fcb200_profile_read(ms, cnt, n)
    return {PASS_NAMES[i]: (float(ms[i]), int(cnt[i])) for i in range(n)}
