#!/usr/bin/env python
"""Per-pass device times (CUDA events inside the library) for one workload; one JSON line.
usage: pass_times.py [steps] [d0 d1 d2 k0 k1 k2]   (env knobs: FCB200_COL_VARIANT, FCB200_COL_THREADS, ...)"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
im_dim = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 8 else bench.IM_DIM
k_dim = tuple(int(v) for v in sys.argv[5:8]) if len(sys.argv) >= 8 else bench.K_DIM
n = int(np.prod(im_dim))
d_im = torch.rand(n, device="cuda:0") * 1000
d_k = torch.from_numpy(bench.gaussian_psf(k_dim).reshape(-1)).cuda()
st = torch.cuda.current_stream().cuda_stream
if os.environ.get("TOOL_SAVEMEMORY") == "1":       # time convolution3DfftCUDAInPlaceSaveMemory's path instead
    fc.convolve_device_async = fc.convolve_device_async_savememory
for _ in range(3):
    fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
e1.record()
torch.cuda.synchronize()
total = e0.elapsed_time(e1) / steps
fc.profile_enable(True)
fc.profile_read()
for _ in range(steps):
    fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
torch.cuda.synchronize()
prof = fc.profile_read()
out = {k: round(ms / c, 4) for k, (ms, c) in prof.items() if c}
env = {k: v for k, v in os.environ.items() if k.startswith("FCB200_") or k.startswith("TOOL_")}
print(json.dumps({"dims": im_dim + k_dim, "ms_step": round(total, 4), "Mvox_s": round(n / total / 1e3, 0), "passes": out, "env": env}))
