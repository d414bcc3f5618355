#!/usr/bin/env python
"""Diagnostic: does a pageable call slow the pinned host paths that follow it?  Prints one JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("FCB200_PSF_CACHE", "0")
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

im_dim, k_dim = bench.IM_DIM, bench.K_DIM
n = int(np.prod(im_dim))
k = bench.gaussian_psf(k_dim).reshape(-1)
rng = np.random.default_rng(1)
base = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
pinned = [torch.from_numpy(base).pin_memory() for _ in range(6)]
pk = torch.from_numpy(k).pin_memory()


def t(fn, reps=3):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append(round((time.perf_counter() - t0) * 1e3, 2))
    return ts


out = {}
fc.convolve_batch(pinned[:3], im_dim, pk.numpy(), k_dim, 0)
out["single_pinned_before"] = t(lambda: fc.convolution3DfftCUDAInPlace(pinned[0].numpy(), im_dim, pk.numpy(), k_dim, 0))
out["batch6_pinned_before"] = t(lambda: fc.convolve_batch(pinned, im_dim, pk.numpy(), k_dim, 0))
pg = base.copy()
out["single_pageable"] = t(lambda: fc.convolution3DfftCUDAInPlace(pg, im_dim, k, k_dim, 0))
out["single_pinned_after"] = t(lambda: fc.convolution3DfftCUDAInPlace(pinned[0].numpy(), im_dim, pk.numpy(), k_dim, 0))
out["batch6_pinned_after"] = t(lambda: fc.convolve_batch(pinned, im_dim, pk.numpy(), k_dim, 0))
del pg
out["batch6_pinned_after_del"] = t(lambda: fc.convolve_batch(pinned, im_dim, pk.numpy(), k_dim, 0))
fc.release()
out["batch6_pinned_after_release"] = t(lambda: fc.convolve_batch(pinned, im_dim, pk.numpy(), k_dim, 0))
print(json.dumps(out))
