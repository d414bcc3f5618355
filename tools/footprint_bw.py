#!/usr/bin/env python
"""Does streaming bandwidth depend on the footprint?  Copy (cudaMemcpyAsync D2D and an elementwise kernel) over buffers of
growing size, one JSON line each.  Context for the pass times of very large volumes (tools/pass_times.py)."""
import json
import torch

dev = torch.device("cuda:0")
for gb in (0.25, 1, 4, 16, 32):
    n = int(gb * (1 << 30)) // 4
    a = torch.empty(n, dtype=torch.float32, device=dev).fill_(1.0)
    b = torch.empty_like(a)
    res = {"GiB_per_buffer": gb}
    for name, fn in (("memcpy", lambda: b.copy_(a)), ("kernel_add", lambda: torch.add(a, 1.0, out=b)),
                     ("kernel_read", lambda: a.sum()), ("kernel_write", lambda: b.fill_(2.0))):
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        reps = max(2, int(8 / gb))
        e0.record()
        for _ in range(reps):
            fn()
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / reps
        traffic = n * 4 * (1 if name in ("kernel_read", "kernel_write") else 2)
        res[name + "_GBs"] = round(traffic / ms / 1e6, 0)
    print(json.dumps(res), flush=True)
    del a, b
    torch.cuda.empty_cache()
