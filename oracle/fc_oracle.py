"""CPU oracle for the 3D FFT-convolution hot path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may
import this module.  The product path (fourierconvolutioncudalib_b200) never does; it fails
loudly when its CUDA library is missing.

What is restated here (reference = /root/reference, StephanPreibisch/FourierConvolutionCUDALib):

* place_psf()            -- fftShiftKernel as *called* by convolution3DfftCUDAInPlace
                            (src/convolution3Dfft.cu:128-166, call site :454-461)
* convolve_inplace_ref() -- convolution3DfftCUDAInPlace end to end (src/convolution3Dfft.cu:399-578):
                            R2C(image) * R2C(placed PSF) * 1/N -> C2R, on the grid the caller passed.
                            The FFT itself lives in NVIDIA cuFFT (closed source, not in /root/reference;
                            the reference pins no version; this image ships cuFFT 11.4.1 / CUDA 12.9).
                            Its published contract -- unnormalised forward/inverse DFT -- is restated with
                            numpy's pocketfft in float64.
* direct_convolve()      -- tests/test_algorithms.hpp:10-58 (the reference test-suite's CPU oracle),
                            numpy version for small cases; the C version is oracle/direct_convolve.c.
* zero_padd helpers      -- tests/padd_utils.h:12-38,99-115,157-171
* l2norm()               -- tests/test_utils.hpp:75-88 (note: divides by N, not sqrt(N))

Pinning status: pinned.  tests/test_oracle.py checks this module against the closed-form /
direct-convolution expectations the reference's own tests use (tests/test_gpu_convolve.cpp,
tests/test_gpu_numerical_stability.cpp) and tests/test_parity_gpu.py checks it -- and the CUDA
product -- against the reference's own cuFFT build (oracle/_ref, built by oracle/Makefile from the
sources under /root/reference) on the GPU box.
"""
import numpy as np


def psf_tap_positions(kernel_dim, im_dim):
    """Flat position (int64) in the padded volume of every PSF tap t = 0..K-1, exactly as the reference's
    fftShiftKernel computes it (src/convolution3Dfft.cu:139-165 with the arguments of :454-461:
    (k0,k1,k2, d0,d1,d2) -- legacy convention, index 2 fastest)."""
    k0, k1, k2 = (int(v) for v in kernel_dim)
    d0, d1, d2 = (int(v) for v in im_dim)
    t = np.arange(k0 * k1 * k2, dtype=np.int64)
    c = t % k2
    aux = (t - c) // k2
    b = aux % k1
    a = (aux - b) // k1
    a = a - k0 // 2
    b = b - k1 // 2
    c = c - k2 // 2
    a = np.where(a < 0, a + d0, a)
    b = np.where(b < 0, b + d1, b)
    c = np.where(c < 0, c + d2, c)
    return c + d2 * (b + d1 * a)


def place_psf(kernel, kernel_dim, im_dim):
    """Zero-padded, circularly shifted PSF exactly as the reference builds it.

    kernel: flat float array of k0*k1*k2 values; kernel_dim=(k0,k1,k2); im_dim=(d0,d1,d2).
    Returns the flat float64 array S of d0*d1*d2 values; the result is later consumed with d0 fastest
    (src/convolution3Dfft.cu:474-486) although the positions were computed with index 2 fastest.
    """
    d0, d1, d2 = (int(v) for v in im_dim)
    kernel = np.asarray(kernel).reshape(-1)
    pos = psf_tap_positions(kernel_dim, im_dim)
    assert kernel.size == pos.size
    S = np.zeros(d0 * d1 * d2, dtype=np.float64)
    # the reference scatters with one thread per tap; positions are distinct when k_i <= d_i
    S[pos] = kernel.astype(np.float64)
    return S


def convolve_inplace_ref(im, im_dim, kernel, kernel_dim):
    """float64 model of convolution3DfftCUDAInPlace; returns the flat result (same layout as im)."""
    d0, d1, d2 = (int(v) for v in im_dim)
    I3 = np.asarray(im, dtype=np.float64).reshape(d2, d1, d0)          # d0 fastest (:417-421,:500-510)
    S3 = place_psf(kernel, kernel_dim, im_dim).reshape(d2, d1, d0)     # same strides (:474-486)
    F = np.fft.rfftn(I3) * np.fft.rfftn(S3)                            # :523-536 (scale folded below)
    out = np.fft.irfftn(F, s=I3.shape, axes=(0, 1, 2))                               # numpy divides by N == scale at :531
    return out.reshape(-1)


def direct_convolve(image, kernel, offset):
    """tests/test_algorithms.hpp:10-58 on [z][y][x] arrays; float32 accumulate like the reference."""
    image = np.asarray(image, dtype=np.float32)
    kernel = np.asarray(kernel, dtype=np.float32)
    res = np.array(image, dtype=np.float32, copy=True)
    hk = [s // 2 for s in kernel.shape]
    flipped = kernel[::-1, ::-1, ::-1]
    Z, Y, X = image.shape
    for z in range(offset[0], Z - offset[0]):
        for y in range(offset[1], Y - offset[1]):
            for x in range(offset[2], X - offset[2]):
                acc = np.float32(0)
                win = image[z - hk[0]:z - hk[0] + kernel.shape[0],
                            y - hk[1]:y - hk[1] + kernel.shape[1],
                            x - hk[2]:x - hk[2] + kernel.shape[2]]
                # same accumulation order as the reference loop nest (kz, ky, kx)
                for v in (flipped * win).reshape(-1):
                    acc = np.float32(acc + v)
                res[z, y, x] = acc
    return res


def zero_padd_extents(image_shape, kernel_shape, factor=1):
    """tests/padd_utils.h:12-24,99-108: extent_i = image_i + 2*factor*(kernel_i/2)."""
    return tuple(int(i) + 2 * factor * (int(k) // 2) for i, k in zip(image_shape, kernel_shape))


def zero_padd_offsets(kernel_shape, factor=1):
    """tests/padd_utils.h:26-38,110-114: offset_i = (kernel_i/2)*factor."""
    return tuple((int(k) // 2) * factor for k in kernel_shape)


def zero_padd(image, kernel_shape, factor=1):
    """tests/padd_utils.h:157-171: embed image in a zero volume at the offsets."""
    ext = zero_padd_extents(image.shape, kernel_shape, factor)
    off = zero_padd_offsets(kernel_shape, factor)
    out = np.zeros(ext, dtype=image.dtype)
    out[off[0]:off[0] + image.shape[0], off[1]:off[1] + image.shape[1], off[2]:off[2] + image.shape[2]] = image
    return out, off


def crop(padded, off, shape):
    return padded[off[0]:off[0] + shape[0], off[1]:off[1] + shape[1], off[2]:off[2] + shape[2]]


def smooth_extent(e, even=False):
    """smallest 7-smooth integer >= e (even if asked): the padding policy 1 of the padded entry points"""
    def smooth(v):
        for f in (2, 3, 5, 7):
            while v % f == 0:
                v //= f
        return v == 1
    e = int(e)
    while not smooth(e) or (even and e % 2):
        e += 1
    return e


def padded_extents(im_dim, kernel_dim, policy):
    """policy 0: zero_padd_extents (tests/padd_utils.h:12-24,99-108); 1: rounded up to 7-smooth, im_dim[0] even."""
    ext = zero_padd_extents(im_dim, kernel_dim)
    if policy == 1:
        ext = tuple(smooth_extent(e, even=(i == 0)) for i, e in enumerate(ext))
    return ext


def convolve_padded_ref(im, im_dim, kernel, kernel_dim, mode=0, policy=0):
    """What a reference caller does around convolution3DfftCUDAInPlace (tests/test_fixtures.hpp:254-268 with
    tests/padd_utils.h:157-171): embed the volume at offsets kernel/2 in the padded grid (zeros, or
    numpy "reflect" mirroring for mode 1), convolve on that grid, read the interior back.  Flat float64 result."""
    d0, d1, d2 = (int(v) for v in im_dim)
    p0, p1, p2 = padded_extents(im_dim, kernel_dim, policy)
    o0, o1, o2 = zero_padd_offsets(kernel_dim)
    I3 = np.asarray(im, dtype=np.float64).reshape(d2, d1, d0)
    widths = ((o2, p2 - d2 - o2), (o1, p1 - d1 - o1), (o0, p0 - d0 - o0))
    if mode == 0:
        P3 = np.pad(I3, widths, mode="constant")
    else:
        # numpy's reflect folds repeatedly when the halo is wider than the axis, like the library's index fold
        P3 = np.pad(I3, widths, mode="reflect") if min(I3.shape) > 1 else _reflect_pad_degenerate(I3, widths)
    out = convolve_inplace_ref(P3.reshape(-1), (p0, p1, p2), kernel, kernel_dim).reshape(p2, p1, p0)
    return crop(out, (o2, o1, o0), (d2, d1, d0)).reshape(-1)


def _reflect_pad_degenerate(a, widths):
    """reflect padding where an axis has a single sample (numpy refuses): that axis is replicated"""
    for ax, (lo, hi) in enumerate(widths):
        if a.shape[ax] == 1:
            a = np.repeat(a, lo + hi + 1, axis=ax)
        else:
            w = [(0, 0)] * a.ndim
            w[ax] = (lo, hi)
            a = np.pad(a, w, mode="reflect")
    return a


def modulate_subsampled_ref(im, im_dim, kernel, kernel_dim):
    """numpy restatement of modulateAndNormalizeSubsampled_kernel (src/convolution3Dfft.cu:65-125), the only piece of the
    legacy "SaveMemory" idea that survives in the reference snapshot (no host function calls it; the declaration at
    src/convolution3Dfft.h:58-64 is commented out).  Host flow as the header comment describes it: FFT of the image at
    image size, FFT of the RAW (unpadded, unshifted) PSF at PSF size, this kernel, inverse FFT.  Legacy convention:
    index 2 is the fastest axis (:80-84).  Returns the flat float64 result.

    Kept in the oracle ONLY to document why the product does not offer these semantics: the result is a
    nearest-neighbour resampling of the PSF spectrum with a 0/pi phase stand-in (:106-117) and is several per cent away
    from the convolution (tests/test_oracle.py::test_legacy_subsampled_multiply_is_not_a_convolution)."""
    d0, d1, d2 = (int(v) for v in im_dim)
    k0, k1, k2 = (int(v) for v in kernel_dim)
    F = np.fft.rfftn(np.asarray(im, np.float64).reshape(d0, d1, d2))                 # [d0][d1][d2/2+1]
    K = np.fft.rfftn(np.asarray(kernel, np.float64).reshape(k0, k1, k2)).reshape(-1)   # [k0][k1][k2/2+1], flat
    r0, r1, r2 = (np.float32(k0) / np.float32(d0), np.float32(k1) / np.float32(d1), np.float32(k2) / np.float32(d2))
    i0, i1, i2 = np.meshgrid(np.arange(d0), np.arange(d1), np.arange(d2 // 2 + 1), indexing="ij")     # :80-84
    q2 = (r2 * i2.astype(np.float32) + np.float32(0.5)).astype(np.int64) % k2                         # :106
    q1 = (r1 * i1.astype(np.float32) + np.float32(0.5)).astype(np.int64) % k1                         # :107
    q0 = (r0 * i0.astype(np.float32) + np.float32(0.5)).astype(np.int64) % k0                         # :108
    flat = q2 + (1 + k2 // 2) * (q1 + k1 * q0)                                                        # :110-111
    a = K[np.minimum(flat, K.size - 1)]            # odd k2: q2 may be k2/2+1, one past the row (clamped at the very end)
    a = np.where((q0 + q1 + q2) % 2 == 1, -a, a)                                                      # :113-117
    out = np.fft.irfftn(a * F, s=(d0, d1, d2), axes=(0, 1, 2))     # c = 1/N (:118-121) is numpy's inverse normalisation
    return out.reshape(-1)


def l2norm(reference, data):
    """tests/test_utils.hpp:75-88: sqrt(sum (a-b)^2) / N  (N, not sqrt(N))."""
    a = np.asarray(reference, dtype=np.float32).reshape(-1).astype(np.float64)
    b = np.asarray(data, dtype=np.float32).reshape(-1).astype(np.float64)
    return float(np.sqrt(np.sum((b - a) ** 2)) / a.size)
