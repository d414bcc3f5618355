#!/usr/bin/env python
"""A/B of radix sequences (FCB200_PLAN overrides, fc_plan.cu: plan_override) for one volume, in one process:
  plan_ab.py d0 d1 d2 k0 k1 k2 steps  "<override 1>" "<override 2>" ...        ("" = the planner's own choice)
Each variant: plans released, the override set, result compared with the planner's own (max abs diff / max abs),
per-pass device times.  One JSON line per variant."""
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
steps = int(sys.argv[7])
variants = [""] + [v for v in sys.argv[8:] if v]
os.environ.setdefault("FCB200_PSF_CACHE", "0")
n = int(np.prod(im_dim))
src = torch.rand(n, device="cuda:0") * 1000
d_k = torch.from_numpy(bench.gaussian_psf(k_dim).reshape(-1)).cuda()
st = torch.cuda.current_stream().cuda_stream
want = None
for v in variants:
    fc.release()
    if v:
        os.environ["FCB200_PLAN"] = v
    else:
        os.environ.pop("FCB200_PLAN", None)
    d_im = src.clone()
    fc.convolve_device_async(d_im, im_dim, d_k, k_dim, 0, st)
    torch.cuda.synchronize()
    if want is None:
        want = d_im.clone()
        err = 0.0
    else:
        err = float((d_im - want).abs().max() / want.abs().max())
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
    fc.profile_enable(False)
    out = {k: round(ms / c, 4) for k, (ms, c) in prof.items() if c}
    rad = {ax: fc.plan_radices(L, style) for ax, L, style in (("x", im_dim[2] // 2, 2), ("y", im_dim[1], 0), ("z", im_dim[0], 1))}
    print(json.dumps({"dims": im_dim, "plan": v, "radices": rad, "ms_step": round(total, 4), "max_rel_diff": err, "passes": out}), flush=True)
