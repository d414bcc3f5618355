#!/usr/bin/env python
"""Small shapes that route through every kernel family added in round 2 (TMA pipeline incl. 64- / 32-byte rows, fused and
on-the-fly fused z, register butterflies 11..23, symmetric direct sum, Rader stage, persistent row-wise x kernels, tap-row PSF
pass, slab mode with both exchanges, multi-device batch) -- meant to be run under `compute-sanitizer --tool memcheck`; every
result is also checked against the numpy oracle so that the run proves the intended kernels really executed correctly."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np

import fourierconvolutioncudalib_b200 as fc
from oracle import fc_oracle as fo

CASES = [
    ((64, 512, 16), (5, 5, 5), "TMA plain y passes, L = 512"),
    ((64, 16, 256), (5, 5, 5), "TMA fused z, L = 256 (and on the fly through SaveMemory)"),
    ((64, 16, 384), (5, 5, 9), "TMA fused z, L = 384, two slots"),
    ((48, 384, 12), (3, 5, 5), "TMA plain y, L = 384"),
    ((32, 1024, 8), (3, 3, 3), "TMA plain y, L = 1024, 64-byte rows"),
    ((32, 2048, 4), (3, 3, 3), "TMA plain y, L = 2048, 32-byte rows"),
    ((64, 16, 512), (5, 3, 3), "TMA fused z, L = 512, 64-byte rows"),
    ((40, 560, 6), (3, 5, 3), "TMA plain y, L = 560"),
    ((30, 270, 270), (3, 3, 3), "TMA y and fused z, L = 270 = 18 * 15"),
    ((60, 300, 300), (3, 5, 3), "two-stage plans: TMA y and fused z with 300 = 20 * 15, masked PSF z pass"),
    ((300, 12, 10), (3, 3, 3), "tiled x kernel, half-length 150 = 10 * 15"),
    ((24, 420, 420), (3, 3, 5), "TMA y and fused z, L = 420 = 20 * 21"),
    ((24, 448, 560), (3, 3, 3), "TMA y with 448 = 16 * 28, fused z with 560 = 28 * 20 (register kernel)"),
    ((270, 12, 10), (3, 3, 3), "tiled x kernel, half-length 135 = 9 * 15"),
    ((66, 66, 26), (3, 3, 3), "register butterflies 11, 13"),
    ((46, 34, 38), (3, 3, 3), "register butterflies 23, 17, 19"),
    ((148, 74, 106), (5, 3, 3), "symmetric direct sum, primes 37 and 53"),
    ((542, 8, 8), (3, 3, 3), "Rader stage in the x passes (half-length 271)"),
    ((32, 562, 4), (3, 3, 3), "Rader stage in the y passes (2 * 281)"),
    ((16, 8, 326), (3, 3, 3), "Rader stage in the fused z pass (2 * 163)"),
    ((2048, 8, 4), (5, 3, 3), "persistent row-wise x kernel, nx = 2048"),
    ((512, 24, 16), (31, 5, 9), "persistent row-wise x kernel nx = 512, PSF x pass on tap rows"),
]
worst = 0.0
for im_dim, k_dim, what in CASES:
    rng = np.random.default_rng(int(np.prod(im_dim)) % 9973)
    im = (rng.random(int(np.prod(im_dim)), dtype=np.float32) * 100).astype(np.float32)
    k = rng.random(int(np.prod(k_dim)), dtype=np.float32)
    k /= k.sum()
    want = fo.convolve_inplace_ref(im, im_dim, k, k_dim)
    for entry in (fc.convolution3DfftCUDAInPlace, fc.convolution3DfftCUDAInPlaceSaveMemory):
        got = im.copy()
        entry(got, im_dim, k, k_dim, 0)
        err = float(np.abs(got - want).max() / np.abs(want).max())
        worst = max(worst, err)
        assert err <= 1e-4, (what, entry.__name__, err)
    print(f"ok  {str(im_dim):18s} {what}", flush=True)
    fc.release()

# slab mode (emulated ranks on device 0), both forward exchanges, and the multi-device batch
im_dim, k_dim = (64, 48, 40), (5, 7, 3)
rng = np.random.default_rng(3)
im = (rng.random(int(np.prod(im_dim)), dtype=np.float32) * 100).astype(np.float32)
k = rng.random(int(np.prod(k_dim)), dtype=np.float32)
want = fo.convolve_inplace_ref(im, im_dim, k, k_dim)
for x in ("1", "0", "2"):
    os.environ["FCB200_SLAB_EXCHANGE"] = x
    got = im.copy()
    fc.convolve_slab(got, im_dim, k, k_dim, [0, 0, 0, 0])
    assert float(np.abs(got - want).max() / np.abs(want).max()) <= 1e-4
    print(f"ok  slab, 4 ranks, exchange {x}", flush=True)
blocks = [im.copy() for _ in range(5)]
fc.convolve_batch_multi(blocks, im_dim, k, k_dim, [0])
for b in blocks:
    assert float(np.abs(b - want).max() / np.abs(want).max()) <= 1e-4
print("ok  batch_multi")
fc.release()
print(f"all cases passed, worst max error {worst:.2e} of max|out|")
