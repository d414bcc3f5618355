#!/usr/bin/env python
"""Times config 5 in slab mode over all GPUs of the box through fcb200_convolve_slab_device (one JSON line).
Environment knobs are passed through (FCB200_SLAB_EXCHANGE, FCB200_SLAB_XSTREAMS, FCB200_SLAB_XCHUNKS);
TOOL_HOST_PSF=1 hands the PSF over as a host pointer so that the PSF-spectrum slabs are cached (phases without PSF work)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

P = int(sys.argv[1]) if len(sys.argv) > 1 else torch.cuda.device_count()
devs = list(range(P))
dims, kd = bench.C5_DIM, bench.C5_K
plane = dims[0] * dims[1]
nzp, nyl, planes = fc.slab_partition(dims, P)
slabs = [torch.rand(planes[r] * plane, device=f"cuda:{r}") * 1000 for r in range(P)]
psf = bench.gaussian_psf(kd).reshape(-1)
host_psf = os.environ.get("TOOL_HOST_PSF") == "1"
k = psf if host_psf else torch.from_numpy(psf).to("cuda:0")
if host_psf:
    os.environ["FCB200_PSF_CACHE"] = "1"
for _ in range(3):
    fc.convolve_slab_device(slabs, dims, k, kd, devs)
tot, ph = 0.0, np.zeros(3)
steps = 8
for _ in range(steps):
    fc.convolve_slab_device(slabs, dims, k, kd, devs)
    t = np.array(fc.slab_last_timing(dims, devs))
    tot += t[:, 3].max()
    ph += t[:, :3].max(axis=0)
print(json.dumps({"gpus": P, "ms": round(tot / steps, 3), "phases": [round(v / steps, 3) for v in ph], "psf_cached": host_psf,
                  "env": {k: v for k, v in os.environ.items() if k.startswith("FCB200_SLAB")}}))
