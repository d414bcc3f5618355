#!/usr/bin/env python
"""One call of the REFERENCE build (oracle/_ref) on the bench workload -- for an ncu launch list of
its device work (cuFFT kernels + modulateAndNormalize_kernel).  Checker/baseline only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import numpy as np

import bench
import reflib

n = int(np.prod(bench.IM_DIM))
im = (np.random.default_rng(0).random(n, dtype=np.float32) * 1000).astype(np.float32)
psf = bench.gaussian_psf(bench.K_DIM).reshape(-1)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 1):
    out = reflib.convolve_inplace(im, bench.IM_DIM, psf, bench.K_DIM, 0)
print("ref done", float(out[0]))
