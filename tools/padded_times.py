#!/usr/bin/env python
"""In-library padding against caller-side padding (SURVEY 8(d) "caller-padded" flavour), per named config:
  python tools/padded_times.py d0 d1 d2 k0 k1 k2 [reps]
device-resident (CUDA events on the launching stream, median) and end to end from pinned host memory (wall clock
around the blocking call).  Caller-side = numpy zero padding to the same 7-smooth grid, convolution3DfftCUDAInPlace
on it, crop on the host (what the reference's tests do, tests/padd_utils.h + tests/test_fixtures.hpp:254-268);
only the library call is timed for it, i.e. the host-side pad/crop loops are NOT charged.  One JSON line."""
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

os.environ.setdefault("FCB200_PSF_CACHE", "0")
im_dim = tuple(int(v) for v in sys.argv[1:4])
k_dim = tuple(int(v) for v in sys.argv[4:7])
reps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
n = int(np.prod(im_dim))
p_dim = fc.padded_extents(im_dim, k_dim, fc.PAD_SMOOTH)
n_pad = int(np.prod(p_dim))
k = bench.gaussian_psf(k_dim).reshape(-1)
rng = np.random.default_rng(1234)
base = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
dev = 0
stream = torch.cuda.current_stream(dev).cuda_stream


def med_events(fn, reps):
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.median(ts))


def med_wall(fn, reps):
    ts = []
    for _ in range(reps):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        fn()
        torch.cuda.synchronize()
        ts.append((time.perf_counter() - t0) * 1e3)
    return float(np.median(ts))


out = {"imDim": im_dim, "kernelDim": k_dim, "padDim": p_dim}
d_k = torch.from_numpy(k).to(f"cuda:{dev}")
d_im = torch.from_numpy(base).to(f"cuda:{dev}")
d_pad = torch.zeros(n_pad, dtype=torch.float32, device=f"cuda:{dev}")
for _ in range(3):
    fc.convolve_device_async(d_pad, p_dim, d_k, k_dim, dev, stream)
    fc.convolve_padded_device_async(d_im, im_dim, d_k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH, stream)
    fc.convolve_padded_device_async(d_im, im_dim, d_k, k_dim, dev, fc.PAD_MIRROR, fc.PAD_SMOOTH, stream)
out["device_caller_padded_ms"] = round(med_events(lambda: fc.convolve_device_async(d_pad, p_dim, d_k, k_dim, dev, stream), reps), 4)
out["device_library_zero_ms"] = round(med_events(lambda: fc.convolve_padded_device_async(
    d_im, im_dim, d_k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH, stream), reps), 4)
out["device_library_mirror_ms"] = round(med_events(lambda: fc.convolve_padded_device_async(
    d_im, im_dim, d_k, k_dim, dev, fc.PAD_MIRROR, fc.PAD_SMOOTH, stream), reps), 4)
fc.profile_enable(True)
fc.convolve_padded_device_async(d_im, im_dim, d_k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH, stream)
torch.cuda.synchronize()
out["library_zero_passes_ms"] = {kk: round(v[0], 4) for kk, v in fc.profile_read().items() if v[1]}
fc.profile_enable(False)
del d_pad, d_im
torch.cuda.empty_cache()

h_pad = torch.zeros(n_pad, dtype=torch.float32).pin_memory()
h_im = torch.from_numpy(base.copy()).pin_memory()
ereps = max(3, reps // 4)
for _ in range(2):
    fc.convolution3DfftCUDAInPlace(h_pad, p_dim, k, k_dim, dev)
    fc.convolve_padded(h_im, im_dim, k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH)
    fc.convolve_padded(h_im, im_dim, k, k_dim, dev, fc.PAD_MIRROR, fc.PAD_SMOOTH)
out["e2e_pinned_caller_padded_ms"] = round(med_wall(lambda: fc.convolution3DfftCUDAInPlace(h_pad, p_dim, k, k_dim, dev), ereps), 3)
out["e2e_pinned_library_zero_ms"] = round(med_wall(lambda: fc.convolve_padded(h_im, im_dim, k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH), ereps), 3)
out["e2e_pinned_library_mirror_ms"] = round(med_wall(lambda: fc.convolve_padded(h_im, im_dim, k, k_dim, dev, fc.PAD_MIRROR, fc.PAD_SMOOTH), ereps), 3)
pg = base.copy()
fc.convolve_padded(pg, im_dim, k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH)
out["e2e_pageable_library_zero_ms"] = round(med_wall(lambda: fc.convolve_padded(pg, im_dim, k, k_dim, dev, fc.PAD_ZERO, fc.PAD_SMOOTH), ereps), 3)
for key in list(out):
    if key.endswith("_ms") and isinstance(out[key], float):
        out[key[:-3] + "_Mvox_s"] = round(n / out[key] / 1e3, 0)      # named (unpadded) voxels, SURVEY 8(d)
print(json.dumps(out))
