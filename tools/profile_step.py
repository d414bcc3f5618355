#!/usr/bin/env python
"""Runs a few device-resident convolutions of the bench workload (for ncu captures; prints nothing
that should be read as a benchmark number).  usage: profile_step.py [steps] [d0 d1 d2 k0 k1 k2]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

steps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
im_dim = tuple(int(v) for v in sys.argv[2:5]) if len(sys.argv) >= 8 else bench.IM_DIM
k_dim = tuple(int(v) for v in sys.argv[5:8]) if len(sys.argv) >= 8 else bench.K_DIM
n = int(np.prod(im_dim))
d_im = torch.rand(n, device="cuda:0") * 1000
d_k = torch.from_numpy(bench.gaussian_psf(k_dim).reshape(-1)).cuda()
st = torch.cuda.current_stream().cuda_stream
for _ in range(steps):
    fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
torch.cuda.synchronize()
print("done", steps, im_dim, k_dim)
