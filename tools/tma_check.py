#!/usr/bin/env python
"""A/B of the TMA-staged column kernels (FCB200_TMA=1) against the cp.async / register kernels (FCB200_TMA=0):
same input, both results, max difference and per-pass times.  usage: tma_check.py d0 d1 d2 k0 k1 k2 [steps]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch

import bench
import fourierconvolutioncudalib_b200 as fc

im_dim = tuple(int(v) for v in sys.argv[1:4])
k_dim = tuple(int(v) for v in sys.argv[4:7])
steps = int(sys.argv[7]) if len(sys.argv) > 7 else 20
n = int(np.prod(im_dim))
base = torch.rand(n, device="cuda:0") * 1000
d_k = torch.from_numpy(bench.gaussian_psf(k_dim).reshape(-1)).cuda()
st = torch.cuda.current_stream().cuda_stream
out = {"dims": im_dim + k_dim}
res = {}
# variants: "0" = no TMA kernels, materialised PSF spectrum; "1"/"2" = TMA kernels (FCB200_TMA value), materialised;
#           "otf" = TMA kernels + PSF spectrum derived on the fly in the fused z pass (the default InPlace path)
variants = os.environ.get("TMA_VARIANTS", "0,1,otf").split(",")
for mode in variants:
    os.environ["FCB200_TMA"] = "1" if mode == "otf" else mode
    os.environ["FCB200_OTF_INPLACE"] = "1" if mode == "otf" else "0"
    x = base.clone()
    fc.convolve_device_async(x, im_dim, d_k, k_dim, 0, st)
    torch.cuda.synchronize()
    res[mode] = x
    x = x.clone()
    for _ in range(3):
        fc.convolve_device_async(x, im_dim, d_k, k_dim, 0, st)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        fc.convolve_device_async(x, im_dim, d_k, k_dim, 0, st)
    e1.record()
    torch.cuda.synchronize()
    fc.profile_enable(True)
    fc.profile_read()
    for _ in range(steps):
        fc.convolve_device_async(x, im_dim, d_k, k_dim, 0, st)
    torch.cuda.synchronize()
    prof = fc.profile_read()
    fc.profile_enable(False)
    out["tma=" + mode] = {"ms_step": round(e0.elapsed_time(e1) / steps, 4),
                          "passes": {k: round(ms / c, 4) for k, (ms, c) in prof.items() if c}}
vals = list(res.values())
out["max_rel_diff_vs_first"] = [float((v - vals[0]).abs().max() / vals[0].abs().max()) for v in vals[1:]]
out["window_planes"] = fc.psf_window_planes(im_dim, k_dim, 0)
out["env"] = {k: v for k, v in os.environ.items() if k.startswith("FCB200_") and k not in ("FCB200_TMA", "FCB200_OTF_INPLACE")}
print(json.dumps(out))
