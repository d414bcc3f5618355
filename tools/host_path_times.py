#!/usr/bin/env python
"""End-to-end times of the host-pointer entry points (wall clock around blocking calls, median of `reps`):
  python tools/host_path_times.py d0 d1 d2 k0 k1 k2 [nblocks] [reps]
single call pageable / pinned, batch of nblocks pageable / pinned.  FCB200_STAGING=0 shows the driver's own
pageable copies for comparison.  One JSON line."""
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

im_dim = tuple(int(v) for v in sys.argv[1:4])
k_dim = tuple(int(v) for v in sys.argv[4:7])
nblocks = int(sys.argv[7]) if len(sys.argv) > 7 else 6
reps = int(sys.argv[8]) if len(sys.argv) > 8 else 5
n = int(np.prod(im_dim))
k = bench.gaussian_psf(k_dim).reshape(-1)
rng = np.random.default_rng(1234)
base = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)


def med(fn, reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


out = {"dims": im_dim + k_dim, "nblocks": nblocks, "copy_threads_env": os.environ.get("FCB200_COPY_THREADS"),
       "staging": os.environ.get("FCB200_STAGING", "1"), "host_cores": os.cpu_count()}
pageable = [base.copy() for _ in range(nblocks)]
pinned = [torch.from_numpy(base).pin_memory() for _ in range(nblocks)]
fc.convolution3DfftCUDAInPlace(pageable[0], im_dim, k, k_dim, 0)       # warm-up: plan, buffers, threads
fc.convolve_batch(pinned[:2], im_dim, k, k_dim, 0)
fc.convolve_batch(pageable[:2], im_dim, k, k_dim, 0)
out["single_pageable_ms"] = round(med(lambda: fc.convolution3DfftCUDAInPlace(pageable[0], im_dim, k, k_dim, 0), reps), 3)
out["single_pinned_ms"] = round(med(lambda: fc.convolution3DfftCUDAInPlace(pinned[0], im_dim, k, k_dim, 0), reps), 3)
out["batch_pageable_ms_per_block"] = round(med(lambda: fc.convolve_batch(pageable, im_dim, k, k_dim, 0), reps) / nblocks, 3)
out["batch_pinned_ms_per_block"] = round(med(lambda: fc.convolve_batch(pinned, im_dim, k, k_dim, 0), reps) / nblocks, 3)
for key in list(out):
    if key.endswith("_ms") or key.endswith("_ms_per_block"):
        out[key.replace("_ms_per_block", "_Mvox_s").replace("_ms", "_Mvox_s")] = round(n / out[key] / 1e3, 0)
print(json.dumps(out))
