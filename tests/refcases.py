"""Boost-free port of the reference's test CASES (inputs, call sequence, thresholds).

Sources: /root/reference/tests/test_gpu_convolve.cpp, tests/test_gpu_numerical_stability.cpp,
tests/test_fixtures.hpp:123-292, tests/padd_utils.h.  Arrays are [z][y][x] (boost::multi_array default
order); `dims_reversed` says whether the reference passes the extents reversed ({x,y,z}) or in
storage order (the 8^3 fixture passes padder.extents_ as is -- cubic, so it does not matter).

Each case: dict(name, stack, kernel, factor, check) where check(convolved_crop, case) -> (value, limit, ok).
`convolve(padded_flat, imDim, kernel_flat, kernelDim)` is supplied by the caller (product or oracle).
"""
import numpy as np

from oracle import fc_oracle as fo


def _fixture8():
    img = np.arange(512, dtype=np.float32).reshape(8, 8, 8)        # test_fixtures.hpp:197-204
    k = {}
    k["trivial"] = np.zeros((3, 3, 3), np.float32)
    k["identity"] = np.zeros((3, 3, 3), np.float32)
    k["identity"].reshape(-1)[27 // 2] = 1                        # :145
    for name in ("horizontal", "vertical", "depth"):
        k[name] = np.zeros((3, 3, 3), np.float32)
    for i in range(3):                                            # :147-151
        k["horizontal"][1, 1, i] = i + 1
        k["vertical"][1, i, 1] = i + 1
        k["depth"][i, 1, 1] = i + 1
    k["all1"] = np.ones((3, 3, 3), np.float32)
    return img, k


def legacy_convolution_cases():
    """test_gpu_convolve.cpp:15-199 -- sums compared with BOOST_*_CLOSE at 1e-5 percent."""
    img, kernels = _fixture8()
    cases = []
    for name in ("horizontal", "vertical", "depth", "all1"):
        k = kernels[name]
        padded, off = fo.zero_padd(img, k.shape)
        # expectation: CPU direct convolution of the padded image, cropped, summed in float
        cases.append(dict(name=f"legacy_{name}", image=img, kernel=k, padded=padded, off=off,
                          imDim=list(padded.shape), kernelDim=[3, 3, 3], kind="sum"))
    return cases, img, kernels


def run_case(convolve, stack, kernel, factor=1, reverse_dims=True):
    padded, off = fo.zero_padd(stack, kernel.shape, factor)
    imDim = list(padded.shape[::-1]) if reverse_dims else list(padded.shape)
    kDim = list(kernel.shape[::-1]) if reverse_dims else list(kernel.shape)
    out = convolve(padded.reshape(-1).copy(), imDim, kernel.reshape(-1).copy(), kDim)
    return fo.crop(np.asarray(out).reshape(padded.shape), off, stack.shape)


def asymmetric_cases():
    """test_gpu_convolve.cpp:202-466: x=13, y=17, z=19, 3^3 kernels, dims passed reversed."""
    shape = (19, 17, 13)
    out = []
    k = np.zeros((3, 3, 3), np.float32); k[1, 1, 1] = 1
    stack = np.arange(np.prod(shape), dtype=np.float32).reshape(shape)
    out.append(("identity_convolve_of_prime_shape", stack, k, stack.copy(), 1e-4))
    ones = np.ones(shape, np.float32)
    k = np.zeros((3, 3, 3), np.float32); k[1, 1, 0:3] = 1
    out.append(("horizontal_convolve_of_prime_shape", ones, k, np.full(shape, 3, np.float32), 1e-2))
    k = np.zeros((3, 3, 3), np.float32); k[1, 0:3, 1] = 1
    out.append(("vertical_convolve_of_prime_shape", ones, k, np.full(shape, 3, np.float32), 2e-2))
    k = np.zeros((3, 3, 3), np.float32)
    for i in range(3):
        k[i, i, i] = 1
    out.append(("diagonal_convolve_of_prime_shape", ones, k, np.full(shape, 3, np.float32), 2e-2))
    return out


def stability_cases(max_edge=256):
    """test_gpu_numerical_stability.cpp: (name, stack, kernel, padding factor, expected, threshold)."""
    out = []

    def delta(kshape, val):
        k = np.zeros(kshape, np.float32)
        k[kshape[0] // 2, kshape[1] // 2, kshape[2] // 2] = val
        return k

    s128 = (128, 128, 128)
    c42 = np.full(s128, 42, np.float32)
    out.append(("times_two_128", c42, delta((91, 31, 31), 2), 1, 2 * c42, 1e-3))                       # :40-96
    ramp = np.arange(128 ** 3, dtype=np.float32).reshape(s128)
    two_ramp = (2 * np.arange(128 ** 3)).astype(np.float32).reshape(s128)
    out.append(("ramp_with_tiny_kernel_times_two", ramp, delta((5, 3, 3), 2), 1, two_ramp, 1e-3))     # :98-155
    out.append(("ramp_with_tiny_kernel_times_two_padd_by_10fold", ramp, delta((5, 3, 3), 2), 10,
                two_ramp, 1e-3))                                                                      # :159-215
    for edge, thr in ((16, 1e-2), (128, 1e-3), (256, 1e-3)):                                          # :222-383
        if edge > max_edge:
            continue
        n = edge ** 3
        scale = np.float32(1.0) / np.float32(n)
        st = (np.arange(n, dtype=np.float32) * scale).reshape(edge, edge, edge)
        # the reference's expectation is 2*stack although the delta is 1.0; kept as is
        out.append((f"ramp_normalized_{edge}", st, delta((91, 31, 31), 1), 1, 2 * st, thr))
    return out
