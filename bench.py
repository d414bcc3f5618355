#!/usr/bin/env python
"""Benchmark of the hot path: Mvoxel/s of convolution3DfftCUDAInPlace on BASELINE.json config 3
(512x512x256 fp32 tile (x) 31x31x41 PSF), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one full convolution (PSF spectrum + image path).  `value` is device-resident
(inputs in HBM, async C-ABI extension on the current stream, CUDA events); `e2e` goes through the
reference-facing convolution3DfftCUDAInPlace with HOST (pinned) buffers, copies inside the timed
region.  N > 1 replicates the workload per GPU (independent tiles, no collective: weak scaling).
"""
import argparse
import json
import os
import subprocess
import sys
import tempfile
import time

# the benchmark times the STATELESS path (as the reference's ABI is): the library's transparent PSF-spectrum
# cache for repeated host-pointer PSFs is switched off, so every e2e step recomputes the PSF spectrum too
os.environ.setdefault("FCB200_PSF_CACHE", "0")

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

IM_DIM = (512, 512, 256)        # {d0 fastest, d1, d2}
K_DIM = (31, 31, 41)
METRIC = "Mvoxels/s 3D FFT convolution (512x512x256 (x) 31x31x41)"
UNIT = "Mvoxel/s"


def gaussian_psf(kDim):
    import numpy as np
    ax = [np.exp(-0.5 * ((np.arange(k) - k // 2) / (k / 6.0)) ** 2) for k in kDim]
    psf = ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]
    return (psf / psf.sum()).astype(np.float32)


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--id={self.gpu}", f"--query-gpu={self.Q}",
                                       "--format=csv,noheader,nounits", "-lms", "50"],
                                      stdout=self.f, stderr=subprocess.DEVNULL)
        except OSError:
            self.p = None
            return
        # nvidia-smi needs a moment before its first line; wait so the timed region is covered
        t0 = time.time()
        while time.time() - t0 < 5.0 and os.path.getsize(self.f.name) == 0:
            time.sleep(0.05)

    def mark(self):
        """byte offset of the log now (samples after it were taken after this call)"""
        return os.path.getsize(self.f.name) if self.p is not None else 0

    def stop(self, start_offset=0):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.p.kill()
        self.f.flush()
        self.f.seek(start_offset)
        sm, mx, reasons = [], [], set()
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1]))
                mx.append(float(c[2]))
            except ValueError:
                continue
            for n, v in zip(names, c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        os.unlink(self.f.name)
        if sm:
            s = sorted(sm)
            out = {"sm_mhz": s[len(s) // 2], "sm_max_mhz": max(mx), "reasons": sorted(reasons), "samples": len(sm)}
        return out


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            return float(json.load(open(p))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic(pass_name):
    """per-launch DRAM bytes of a pass from the committed ncu summary (profiles/ncu_traffic.json), or None"""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    if os.path.exists(p):
        try:
            return json.load(open(p)).get(pass_name)
        except Exception:
            return None
    return None


NUMA_BINDING = None


def dist_setup(n_gpus):
    import torch
    from fourierconvolutioncudalib_b200 import tiles
    rank, local, world = tiles.rank_info()
    # before any pinned allocation: host threads of this rank run next to its GPU (FCB200_BENCH_NUMA=0 turns it off)
    global NUMA_BINDING
    if os.environ.get("FCB200_BENCH_NUMA", "1") != "0":
        NUMA_BINDING = tiles.bind_to_gpu_numa_node(local)
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local}"))
    return rank, local, world


def barrier_sync(world):
    import torch
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
    torch.cuda.synchronize()


def max_over_ranks(x, world, device):
    from fourierconvolutioncudalib_b200 import tiles
    return tiles.max_over_ranks(x, device)


def cpu_baseline_sample(max_seconds=30.0):
    """Reference test-suite direct convolution (oracle/direct_convolve.c, OpenMP over z) on a bounded
    sample of the workload: as many interior z planes as host threads, extrapolated per voxel."""
    import numpy as np
    from oracle import c_oracle as co
    d0, d1, d2 = IM_DIM
    # reference convention: kernel array is [k0][k1][k2]; image [z][y][x]; kernel z extent = k0
    k = gaussian_psf(K_DIM)
    kz = K_DIM[0]
    threads = co.lib().fc_oracle_num_threads()
    planes = max(1, min(threads, 128))
    rng = np.random.default_rng(1234)
    slab = (rng.random((planes + 2 * (kz // 2), d1, d0), dtype=np.float32) * 1000).astype(np.float32)
    off = (kz // 2, K_DIM[1] // 2, K_DIM[2] // 2)
    # calibrate on a thin strip first so the sample stays bounded
    t0 = time.perf_counter()
    strip = slab[:, : 2 * off[1] + 8, :].copy()
    co.direct_convolve(strip, k, off, threads="all")
    t_strip = time.perf_counter() - t0
    vox_strip = planes * 8 * (d0 - 2 * off[2])
    rate = vox_strip / max(t_strip, 1e-9)
    rows = int(min(d1 - 2 * off[1], max(8, rate * max_seconds * 0.5 / (planes * (d0 - 2 * off[2])))))
    sample = slab[:, : 2 * off[1] + rows, :].copy()
    t0 = time.perf_counter()
    _, used = co.direct_convolve(sample, k, off, threads="all")
    dt = time.perf_counter() - t0
    vox = planes * rows * (d0 - 2 * off[2])
    return {"value": vox / dt / 1e6, "unit": UNIT, "cores": int(used), "kind": "port",
            "sample": f"direct convolution (tests/test_algorithms.hpp:10-58 restated in C, OpenMP over z) of "
                      f"{planes} z-planes x {rows} rows x {d0 - 2 * off[2]} voxels with the 31x31x41 PSF, {dt:.1f} s"}


def run_ours(args):
    import numpy as np
    import torch
    import fourierconvolutioncudalib_b200 as fc

    rank, local, world = dist_setup(args.gpus)
    dev = local
    torch.cuda.set_device(dev)
    device = torch.device(f"cuda:{dev}")
    n = int(np.prod(IM_DIM))
    rng = np.random.default_rng(1234 + rank)
    im_host = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
    psf = gaussian_psf(K_DIM).reshape(-1)
    d_im = torch.from_numpy(im_host).to(device)
    d_k = torch.from_numpy(psf).to(device)
    stream = torch.cuda.current_stream().cuda_stream

    def step():
        fc.convolve_device_async(d_im, IM_DIM, d_k, K_DIM, dev, stream)

    for _ in range(max(args.warmup, 3)):
        d_im.copy_(torch.from_numpy(im_host).to(device))
        step()
    torch.cuda.synchronize()

    # ---- device-resident timed region: K steps, CUDA events on the launching stream
    sampler = ClockSampler(dev)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier_sync(world)
    clock_mark = sampler.mark() if rank == 0 else 0
    launches0 = fc.launch_count()
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    barrier_sync(world)
    launches = fc.launch_count() - launches0
    ms_total = max_over_ranks(e0.elapsed_time(e1), world, device)
    ms_step = ms_total / args.steps
    value = world * n / (ms_step * 1e-3) / 1e6

    # ---- the SaveMemory entry point's path (PSF spectrum derived on the fly), reported as an extra
    def step_sm():
        fc.convolve_device_async_savememory(d_im, IM_DIM, d_k, K_DIM, dev, stream)
    for _ in range(3):
        step_sm()
    torch.cuda.synchronize()
    s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s0.record()
    for _ in range(args.steps):
        step_sm()
    s1.record()
    torch.cuda.synchronize()
    ms_sm = max_over_ranks(s0.elapsed_time(s1), world, device) / args.steps

    # ---- per-pass device times (CUDA events around every pass, same stream), separate loop
    fc.profile_enable(True)
    fc.profile_read()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    prof = fc.profile_read()
    fc.profile_enable(False)

    g_xc = IM_DIM[0] // 2 + 1
    N, Nc = n, IM_DIM[2] * IM_DIM[1] * g_xc
    alg_bytes = {"x_fwd": 4 * N + 8 * Nc, "y_fwd": 16 * Nc, "z_fused": 24 * Nc, "y_inv": 16 * Nc,
                 "x_inv": 8 * Nc + 4 * N}
    peak, peak_src = measured_peak()
    passes = {}
    for name, (ms, cnt) in prof.items():
        if cnt:
            avg = ms / cnt
            entry = {"ms": round(avg, 4)}
            if name in alg_bytes:
                entry["GBps"] = round(alg_bytes[name] / (avg * 1e-3) / 1e9, 1)
                entry["frac"] = round(entry["GBps"] / peak, 4)
            passes[name] = entry
    dom = max((k for k in passes if k in alg_bytes), key=lambda k: passes[k]["ms"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": passes[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": passes[dom]["frac"], "traffic": ncu_traffic(dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "passes": passes,
                "image_path_bytes": 8 * N + 72 * Nc,
                "image_path_frac": round((8 * N + 72 * Nc) / (sum(passes[k]["ms"] for k in alg_bytes) * 1e-3) / 1e9 / peak, 4)}

    # ---- end to end through the reference-facing ABI with pinned HOST buffers
    h_im = torch.from_numpy(im_host).pin_memory()
    h_k = torch.from_numpy(psf).pin_memory()
    src = torch.from_numpy(im_host)
    e2e_steps = max(3, min(args.steps, 10))
    for _ in range(2):
        h_im.copy_(src)
        fc.convolution3DfftCUDAInPlace(h_im.numpy(), IM_DIM, h_k.numpy(), K_DIM, dev)
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fc.convolution3DfftCUDAInPlace(h_im.numpy(), IM_DIM, h_k.numpy(), K_DIM, dev)
        checksum = float(h_im[0])          # device->host result is already in the caller's buffer
    barrier_sync(world)
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, device) / e2e_steps
    e2e = {"value": world * n / (e2e_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_ms,
           "h2d_bytes_per_step": 4 * n + 4 * psf.size, "d2h_bytes_per_step": 4 * n,
           "api": "convolution3DfftCUDAInPlace(host pinned buffers)", "checksum": checksum,
           "numa_binding": ({"node": NUMA_BINDING[0], "cpus": NUMA_BINDING[1]} if NUMA_BINDING else None)}

    # ---- the same tiles as a pipelined batch (fcb200_convolve_batch: upload b+1 | convolve b | download b-1),
    # reported next to `e2e`, not instead of it: the reference ABI is one volume per call
    nb = 6
    batch = [h_im] + [torch.from_numpy(im_host).pin_memory() for _ in range(nb - 1)]
    fc.convolve_batch(batch[:3], IM_DIM, h_k.numpy(), K_DIM, dev)
    barrier_sync(world)
    t0 = time.perf_counter()
    fc.convolve_batch(batch, IM_DIM, h_k.numpy(), K_DIM, dev)
    barrier_sync(world)
    batch_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, device) / nb
    e2e_batch = {"value": world * n / (batch_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_tile": batch_ms, "tiles": nb,
                 "api": "fcb200_convolve_batch(host pinned buffers): one PSF spectrum per batch, transfers of "
                        "neighbouring tiles overlap the convolution"}
    del batch

    # ---- the same tile with the caller-side padding done in the library (fcb200_convolve_padded, zero padding to the
    # 7-smooth grid 560x560x300; reference callers pad on the host, tests/padd_utils.h:99-171), reported as an extra:
    # named (unpadded) voxels per second, device-resident and end to end from pinned memory
    padded = None
    try:
        pad_dim = fc.padded_extents(IM_DIM, K_DIM, fc.PAD_SMOOTH)
        for _ in range(3):
            fc.convolve_padded_device_async(d_im, IM_DIM, d_k, K_DIM, dev, fc.PAD_ZERO, fc.PAD_SMOOTH, stream)
        torch.cuda.synchronize()
        p0, p1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        psteps = max(3, min(args.steps, 20))
        p0.record()
        for _ in range(psteps):
            fc.convolve_padded_device_async(d_im, IM_DIM, d_k, K_DIM, dev, fc.PAD_ZERO, fc.PAD_SMOOTH, stream)
        p1.record()
        torch.cuda.synchronize()
        pad_ms = max_over_ranks(p0.elapsed_time(p1), world, device) / psteps
        fc.convolve_padded(h_im.numpy(), IM_DIM, h_k.numpy(), K_DIM, dev, fc.PAD_ZERO, fc.PAD_SMOOTH)
        barrier_sync(world)
        t0 = time.perf_counter()
        for _ in range(3):
            fc.convolve_padded(h_im.numpy(), IM_DIM, h_k.numpy(), K_DIM, dev, fc.PAD_ZERO, fc.PAD_SMOOTH)
        barrier_sync(world)
        pad_e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, device) / 3
        padded = {"value": world * n / (pad_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": pad_ms,
                  "e2e_ms_per_step": pad_e2e_ms, "padded_grid": list(pad_dim),
                  "api": "fcb200_convolve_padded(zero padding, 7-smooth grid): padding fused into the x passes, only the "
                         "unpadded bytes cross PCIe"}
    except Exception as exc:      # an extra must never take the benchmark line down
        padded = {"error": str(exc)[:200]}

    # clocks were sampled every 50 ms from the start of the timed loop to the end of the e2e loop
    clocks = sampler.stop(clock_mark) if rank == 0 else None
    line = None
    if rank == 0:
        cpu = cpu_baseline_sample() if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": "512x512x256 fp32 tile (x) 31x31x41 Gaussian PSF (BASELINE config 3), "
                                   "imDim={512,512,256} kernelDim={31,31,41}, PSF spectrum recomputed every step "
                                   "(value and e2e; FCB200_PSF_CACHE=0)",
                       "l2": "inputs larger than L2 (256 MiB image, 260 MiB spectrum)",
                       "parallelism": f"independent tiles, one per GPU x{world}"},
            "e2e": e2e, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "e2e_batch": e2e_batch, "padded": padded,
            "savememory": {"value": world * n / (ms_sm * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_sm,
                           "api": "convolution3DfftCUDAInPlaceSaveMemory path (device-resident): PSF spectrum "
                                  "derived on the fly in the fused z kernel, no image-sized PSF buffer"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
    if world > 1:
        import torch.distributed as dist
        dist.barrier()
        dist.destroy_process_group()
    if line is not None:
        print(json.dumps(line))


def run_reference(args):
    """The reference arm: the UNMODIFIED reference (its own cuFFT build, oracle/_ref, compiled from
    /root/reference/src by oracle/Makefile) through its own convolution3DfftCUDAInPlace with host
    buffers -- the reference's only implementation of this path.  Falls back to the CPU port of the
    reference test-suite's direct convolution when the reference build cannot be loaded."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    import numpy as np
    n = int(np.prod(IM_DIM))
    line = {"impl": "reference", "metric": METRIC, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": "512x512x256 fp32 tile (x) 31x31x41 Gaussian PSF (BASELINE config 3), "
                                   "imDim={512,512,256} kernelDim={31,31,41}"}}
    try:
        import reflib
        lib = reflib.load()
        rng = np.random.default_rng(1234)
        im = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
        psf = gaussian_psf(K_DIM).reshape(-1)
        import ctypes
        idim = (ctypes.c_int * 3)(*IM_DIM)
        kdim = (ctypes.c_int * 3)(*K_DIM)
        buf = im.copy()

        def step():
            lib.convolution3DfftCUDAInPlace(ctypes.c_void_p(buf.ctypes.data), idim, ctypes.c_void_p(psf.ctypes.data),
                                            kdim, 0)
        t0 = time.perf_counter()
        step()
        first = time.perf_counter() - t0
        # bounded: keep the whole run within a few minutes
        steps = max(1, min(args.steps, int(150.0 / max(first, 1e-3))))
        warm = max(0, min(args.warmup, int(30.0 / max(first, 1e-3))) - 1)
        for _ in range(warm):
            step()
        t0 = time.perf_counter()
        for _ in range(steps):
            step()
        dt = (time.perf_counter() - t0) / steps
        v = n / dt / 1e6
        line.update({"value": v, "ms_per_step": dt * 1e3, "steps": steps, "warmup": warm + 1,
                     "cpu_baseline": {"value": v, "unit": UNIT, "cores": 1, "kind": "reference",
                                      "sample": "reference cuFFT build (oracle/_ref) convolution3DfftCUDAInPlace, "
                                                "host buffers, full 512x512x256 volume per step; single host thread "
                                                "drives the GPU (the reference has no CPU-only path)"},
                     "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}})
    except Exception as exc:   # reference build not loadable: time the CPU port instead
        cpu = cpu_baseline_sample()
        line.update({"value": cpu["value"], "ms_per_step": None, "cpu_baseline": cpu,
                     "e2e": {"value": cpu["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                     "note": f"oracle/_ref not loadable ({exc}); CPU port of the test-suite direct convolution timed"})
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
