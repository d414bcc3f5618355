#!/usr/bin/env python
"""Device-resident throughput on the awkward (non-7-smooth) extents the reference's own tests pad to
(/root/reference/tests/test_gpu_numerical_stability.cpp:43-51,101-109,225-233: image + kernel - 1), next to their
7-smooth neighbours.  One JSON line per shape: ms per call (CUDA events), Mvoxel/s, radices per axis."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

SHAPES = [  # (imDim as passed {x,y,z}, kernelDim, note)
    ((158, 158, 218), (31, 31, 91), "128^3 + 31x31x91 - 1: 2*79, 2*109"),
    ((160, 160, 224), (31, 31, 91), "7-smooth neighbour"),
    ((130, 130, 132), (3, 3, 5), "128^3 + 3x3x5 - 1: 2*5*13, 4*3*11"),
    ((135, 135, 135), (3, 3, 5), "7-smooth neighbour"),
    ((148, 148, 168), (21, 21, 41), "4*37"),
    ((150, 150, 168), (21, 21, 41), "7-smooth neighbour"),
    ((46, 46, 106), (31, 31, 91), "16^3 + 31x31x91 - 1: 2*23, 2*53"),
    ((48, 48, 108), (31, 31, 91), "7-smooth neighbour"),
    ((286, 286, 346), (31, 31, 91), "256^3 + 31x31x91 - 1: 2*11*13, 2*173"),
    ((288, 288, 350), (31, 31, 91), "7-smooth neighbour"),
    ((66, 66, 66), (3, 3, 3), "64^3 + 3^3 - 1: 2*3*11"),
    ((70, 70, 70), (3, 3, 3), "7-smooth neighbour"),
    ((542, 542, 296), (31, 31, 41), "C3 + k - 1: 2*271, 8*37"),
    ((560, 560, 300), (31, 31, 41), "7-smooth neighbour"),
    ((408, 408, 444), (25, 25, 61), "C4 + k - 1: 8*3*17, 4*3*37"),
    ((420, 420, 448), (25, 25, 61), "7-smooth neighbour"),
]
steps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
st = torch.cuda.current_stream().cuda_stream
for im_dim, k_dim, note in SHAPES:
    n = int(np.prod(im_dim))
    d_im = torch.rand(n, device="cuda:0") * 1000
    d_k = torch.from_numpy(bench.gaussian_psf(k_dim).reshape(-1)).cuda()
    for _ in range(3):
        fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    rad = [fc.plan_radices(im_dim[0] // 2 if im_dim[0] % 2 == 0 else im_dim[0], 2)[0], fc.plan_radices(im_dim[1], 0)[0],
           fc.plan_radices(im_dim[2], 1)[0]]
    print(json.dumps({"dims": im_dim, "kernel": k_dim, "note": note, "ms": round(ms, 4), "Mvox_s": round(n / ms / 1e3, 0),
                      "radices_xyz": rad}))
    fc.release()
