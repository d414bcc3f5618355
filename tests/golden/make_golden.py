#!/usr/bin/env python
"""Generates the golden fixtures in this directory by running the REFERENCE's own cuFFT build
(oracle/_ref, compiled from /root/reference/src by oracle/Makefile) on a GPU box:

    python tests/golden/make_golden.py gpurun_out/golden      # then copy *.npz into tests/golden/

Each fixture holds seeded inputs and the reference's output of convolution3DfftCUDAInPlace.  Cases
cover cubic and non-cubic volumes (where the reference's PSF placement mixes axis conventions),
odd / even / prime-factor extents and even-sized kernels."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
import reflib

CASES = [  # name, imDim (d0 fastest), kernelDim
    ("cubic16_k3", (16, 16, 16), (3, 3, 3)),
    ("noncubic_15x19x21_k3", (15, 19, 21), (3, 3, 3)),
    ("noncubic_24x20x12_k5x3x7", (24, 20, 12), (5, 3, 7)),
    ("even_kernel_18x14x10_k4x2x6", (18, 14, 10), (4, 2, 6)),
    ("primes_22x26x14_k3x5x3", (22, 26, 14), (3, 5, 3)),
    ("pow2_32x16x8_k7x5x3", (32, 16, 8), (7, 5, 3)),
]


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name, imDim, kDim in CASES:
        rng = np.random.default_rng(abs(hash(name)) % (2 ** 31))
        rng = np.random.default_rng(sum(ord(c) for c in name))
        im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 100).astype(np.float32)
        k = rng.random(int(np.prod(kDim)), dtype=np.float32)
        k = (k / k.sum()).astype(np.float32)
        out = reflib.convolve_inplace(im, imDim, k, kDim, 0)
        np.savez_compressed(os.path.join(outdir, name + ".npz"), im=im, kernel=k, imDim=np.array(imDim, np.int32),
                            kernelDim=np.array(kDim, np.int32), out=out.astype(np.float32))
        print(name, float(np.abs(out).max()))


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(HERE))
