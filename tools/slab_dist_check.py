#!/usr/bin/env python
"""torchrun check + timing of the slab-decomposed path over real GPUs (NCCL all-to-all):
  python -m torch.distributed.run --nproc-per-node P tools/slab_dist_check.py d0 d1 d2 k0 k1 k2 [steps] [check]
With check=1 every rank also convolves the whole volume alone (small shapes only) and compares its
z slab with the slab result.  Prints one JSON line on rank 0 (max over ranks, CUDA events)."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import numpy as np
import torch
import torch.distributed as dist

import bench
import fourierconvolutioncudalib_b200 as fc
from fourierconvolutioncudalib_b200 import slab, tiles

im_dim = tuple(int(v) for v in sys.argv[1:4])
k_dim = tuple(int(v) for v in sys.argv[4:7])
steps = int(sys.argv[7]) if len(sys.argv) > 7 else 5
check = int(sys.argv[8]) if len(sys.argv) > 8 else 0
rank, local, world = tiles.rank_info()
os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
torch.cuda.set_device(local)
device = torch.device(f"cuda:{local}")
dist.init_process_group("nccl", device_id=device)

d0, d1, d2 = im_dim
psf = torch.from_numpy(bench.gaussian_psf(k_dim).reshape(-1)).to(device)
mode = os.environ.get("SLAB_MODE", "nccl")      # nccl: all-to-all; peer: kernels store into the peers' buffers
if mode == "peer":
    conv = slab.PeerSlabConvolver(im_dim, k_dim, rank, world, local, barrier=slab.DistBarrier(device))
    conv.connect_ipc()
else:
    conv = slab.SlabConvolver(im_dim, k_dim, rank, world, local, slab.DistExchange())
conv.prepare_psf(psf)
gen = torch.Generator(device=device)
gen.manual_seed(1234 + rank)
nzl, nzp = conv.nzl, conv.nzp                   # ragged slabs: the last rank owns fewer planes than the pitch
my = torch.rand(nzl * d1 * d0, device=device, generator=gen) * 1000
err = None
if check:
    # assemble the full volume on every rank, convolve it alone, compare this rank's slab
    padded = torch.zeros(nzp * d1 * d0, device=device)
    padded[:my.numel()] = my
    parts = [torch.empty_like(padded) for _ in range(world)]
    dist.all_gather(parts, padded)
    full = torch.cat(parts)[:d0 * d1 * d2].contiguous()
    fc.convolution3DfftCUDAInPlace(full, im_dim, psf, k_dim, local)
    want = full[rank * nzp * d1 * d0:rank * nzp * d1 * d0 + my.numel()].clone()
    got = my.clone()
    conv.convolve(got)
    torch.cuda.synchronize()
    err = float((got - want).abs().max() / want.abs().max())
    l2 = float(torch.linalg.norm(got - want) / torch.linalg.norm(want))
    err = tiles.max_over_ranks(err, device)
    l2 = tiles.max_over_ranks(l2, device)

for _ in range(2):
    conv.convolve(my)
dist.barrier()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(steps):
    conv.convolve(my)
e1.record()
dist.barrier()
torch.cuda.synchronize()
ms = tiles.max_over_ranks(e0.elapsed_time(e1) / steps, device)
if rank == 0:
    n = d0 * d1 * d2
    out = {"dims": im_dim + k_dim, "world": world, "mode": mode, "ms_step": round(ms, 3), "Mvox_s": round(n / ms / 1e3, 0)}
    if check:
        out.update({"max_rel_err_vs_single_gpu": err, "rel_l2_vs_single_gpu": l2})
    print(json.dumps(out))
dist.barrier()
if mode == "peer":
    conv.close()
dist.destroy_process_group()
