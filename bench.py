#!/usr/bin/env python
"""Benchmark of the hot path: Mvoxel/s of convolution3DfftCUDAInPlace on BASELINE.json config 3
(512x512x256 fp32 tile (x) 31x31x41 PSF), one process per GPU.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference]

One step = one full convolution (PSF spectrum + image path).  `value` is device-resident
(inputs in HBM, async C-ABI extension on the current stream, CUDA events); `e2e` goes through the
reference-facing convolution3DfftCUDAInPlace with HOST (pinned) buffers, copies inside the timed
region; `e2e_pageable` is the same call with pageable numpy buffers (what a JNA caller hands over and what the
reference arm uses).  N > 1 replicates the workload per GPU (independent tiles, no collective: weak scaling).

The two SHARDED configurations of BASELINE.json ride along as sub-records of the same JSON line, measured by rank 0
through the library's single-process multi-GPU entry points over the N GPUs of the run (the other ranks wait):
  "slab"     config 5, one 2048x2048x1024 volume (x) 63x63x101 in z slabs over N GPUs (fcb200_convolve_slab_device;
             PSF spectrum slabs rebuilt every step, inside the timed region), strong scaling, max error against the
             1-GPU result, exchange bytes and effective NVLink rate per GPU
  "batch_c4" config 4, 64 host blocks of 384^3 (x) 25x25x61 dealt over N GPUs (fcb200_convolve_batch_multi), next to
             the measured ceiling of concurrent pinned H2D + D2H copies on the same GPUs ("pcie_ceiling")
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
# identical in both arms (ours / reference)
CONFIG = {"workload": "512x512x256 fp32 tile (x) 31x31x41 Gaussian PSF (BASELINE config 3), imDim={512,512,256} "
                      "kernelDim={31,31,41}, PSF spectrum recomputed every step",
          "l2": "inputs larger than L2 (256 MiB image, 260 MiB spectrum)"}
C4_DIM, C4_K, C4_BLOCKS = (384, 384, 384), (25, 25, 61), 64
C5_DIM, C5_K = (2048, 2048, 1024), (63, 63, 101)


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
HOST_GROUP = None


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
        global HOST_GROUP
        HOST_GROUP = dist.new_group(backend="gloo")    # host-side waits must not park a spinning NCCL kernel on the GPUs
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


# ------------------------------------------------------------------------------------------------
# the sharded configurations (BASELINE configs 4 and 5), measured by ONE process over the N GPUs of the run through
# the library's single-process multi-GPU entry points (include/fcb200_ext.h)
# ------------------------------------------------------------------------------------------------
def pcie_ceiling(devs, mib=256, reps=6):
    """aggregate GB/s of concurrent pinned host->device AND device->host cudaMemcpyAsync on all `devs` at once
    (what the host-pointer pipelines are bounded by on this box)"""
    import torch
    n = mib << 18
    bufs = []
    for d in devs:
        up, down = torch.empty(n, dtype=torch.float32, pin_memory=True), torch.empty(n, dtype=torch.float32, pin_memory=True)
        up.fill_(1.0)
        dv = [torch.empty(n, dtype=torch.float32, device=f"cuda:{d}") for _ in range(2)]
        bufs.append((d, up, down, dv, torch.cuda.Stream(device=d), torch.cuda.Stream(device=d)))

    def run(reps, h2d=True, d2h=True):
        for d, up, down, dv, s0, s1 in bufs:
            for _ in range(reps):
                if h2d:
                    with torch.cuda.stream(s0):
                        dv[0].copy_(up, non_blocking=True)
                if d2h:
                    with torch.cuda.stream(s1):
                        down.copy_(dv[1], non_blocking=True)
        for d in devs:
            torch.cuda.synchronize(d)
    run(2)
    out = {}
    for name, kw in (("bidir", {}), ("h2d_only", {"d2h": False}), ("d2h_only", {"h2d": False})):
        t0 = time.perf_counter()
        run(reps, **kw)
        dt = time.perf_counter() - t0
        per_dir = len(devs) * reps * n * 4 / dt / 1e9
        out[name + "_GBps_per_direction"] = round(per_dir, 1)
    out["devices"] = len(devs)
    out["how"] = f"{mib} MiB pinned buffers, {reps} copies per direction per GPU, all GPUs at once, wall clock"
    return out


def bench_batch_c4(fc, devs, ceiling):
    """config 4: 64 host (pinned) blocks of 384^3 with one 25x25x61 PSF over the devices, one shared queue"""
    import numpy as np
    import torch
    n = int(np.prod(C4_DIM))
    psf = gaussian_psf(C4_K).reshape(-1)
    first = torch.empty(n, dtype=torch.float32, pin_memory=True)
    first.copy_(torch.rand(n, device=f"cuda:{devs[0]}") * 1000)
    blocks = [first]
    for _ in range(C4_BLOCKS - 1):
        b = torch.empty(n, dtype=torch.float32, pin_memory=True)
        b.copy_(first)
        blocks.append(b)
    fc.convolve_batch_multi(blocks[:min(C4_BLOCKS, 3 * len(devs))], C4_DIM, psf, C4_K, devs)      # warm-up: plans, rings
    t0 = time.perf_counter()
    taken = fc.convolve_batch_multi(blocks, C4_DIM, psf, C4_K, devs)
    dt = time.perf_counter() - t0
    checksum = float(blocks[-1][n // 2])
    moved = 2.0 * C4_BLOCKS * n * 4
    rec = {"value": C4_BLOCKS * n / dt / 1e6, "unit": UNIT, "ms_per_block": dt * 1e3 / C4_BLOCKS, "blocks": C4_BLOCKS,
           "n_gpus": len(devs), "blocks_per_gpu": taken, "host_buffers": "pinned",
           "h2d_plus_d2h_GBps": round(moved / dt / 1e9, 1), "checksum": checksum,
           "api": "fcb200_convolve_batch_multi: one pipelined batch (upload | convolve | download) per GPU, blocks taken "
                  "from one shared counter; PSF spectrum once per GPU; wall clock around the call"}
    if ceiling:
        rec["frac_of_pcie_ceiling"] = round(moved / 2 / dt / 1e9 / ceiling["bidir_GBps_per_direction"], 3)
    return rec


def bench_slab_c5(fc, devs, steps=5):
    """config 5: one 2048x2048x1024 volume (x) 63x63x101, device-resident, in z slabs over the devices"""
    import numpy as np
    import torch
    P = len(devs)
    d0, d1, d2 = C5_DIM
    n = d0 * d1 * d2
    plane = d0 * d1
    nzp, nyl, planes = fc.slab_partition(C5_DIM, P)
    xcp = fc.spectrum_pitch(d0)
    d_k = torch.from_numpy(gaussian_psf(C5_K).reshape(-1)).to(f"cuda:{devs[0]}")

    def make_slabs():
        out = []
        for r, d in enumerate(devs):
            g = torch.Generator(device=f"cuda:{d}")
            g.manual_seed(77 + r)
            out.append(torch.rand(planes[r] * plane, device=f"cuda:{d}", generator=g) * 1000)
        return out

    slabs = make_slabs()
    rec = {"workload": "2048x2048x1024 fp32 volume (x) 63x63x101 Gaussian PSF (BASELINE config 5), device-resident, "
                       "PSF spectrum rebuilt every step", "n_gpus": P, "scaling": "strong", "unit": UNIT}
    if P == 1:
        stream = torch.cuda.current_stream(devs[0]).cuda_stream
        for _ in range(2):
            fc.convolve_device_async(slabs[0], C5_DIM, d_k, C5_K, devs[0], stream)
        torch.cuda.synchronize(devs[0])
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(steps):
            fc.convolve_device_async(slabs[0], C5_DIM, d_k, C5_K, devs[0], stream)
        e1.record()
        torch.cuda.synchronize(devs[0])
        ms = e0.elapsed_time(e1) / steps
        rec.update({"ms_per_step": ms, "value": n / ms / 1e3, "api": "fcb200_convolve_device_async (one GPU)",
                    "max_rel_err_vs_1gpu": 0.0})
        del slabs
        fc.release()
        torch.cuda.empty_cache()
        return rec
    for _ in range(2):
        fc.convolve_slab_device(slabs, C5_DIM, d_k, C5_K, devs)
    tot, phases, wall = 0.0, np.zeros(3), 0.0
    for _ in range(steps):
        t0 = time.perf_counter()
        fc.convolve_slab_device(slabs, C5_DIM, d_k, C5_K, devs)
        wall += time.perf_counter() - t0
        t = np.array(fc.slab_last_timing(C5_DIM, devs))      # [rank][forward, z, inverse, total], CUDA events per rank
        tot += t[:, 3].max()
        phases += t[:, :3].max(axis=0)
    ms = tot / steps
    # bytes a rank stores into its peers per exchange: all rows / planes it does not own itself
    sent = (P - 1) * nzp * nyl * xcp * 8
    rec.update({"ms_per_step": ms, "value": n / ms / 1e3, "wall_ms_per_step": wall * 1e3 / steps,
                "phase_ms": {"xy_forward+exchange": phases[0] / steps, "z_fused+exchange (incl. waiting for peers)": phases[1] / steps,
                             "yx_inverse (incl. waiting for peers)": phases[2] / steps},
                "exchange_bytes_per_gpu_per_exchange": int(sent),
                "nvlink_GBps_per_gpu": {"forward_exchange": round(sent / (phases[0] / steps) / 1e6, 1),
                                        "backward_exchange": round(sent / (phases[1] / steps) / 1e6, 1),
                                        "note": "bytes stored into peers / whole phase time (the stores overlap the butterflies)"},
                "api": "fcb200_convolve_slab_device: single process, one worker thread per GPU, the y pass and the fused z "
                       "pass store straight into the peers' buffers (NVLink), phases ordered by CUDA events; timed with "
                       "CUDA events per rank, max over ranks"})
    # ---- parity against the single-GPU result of the same input
    try:
        del slabs
        slabs = make_slabs()
        full = torch.cat([s.to(f"cuda:{devs[0]}") for s in slabs])
        fc.convolve_slab_device(slabs, C5_DIM, d_k, C5_K, devs)
        fc.release()                                     # slab buffers go, the single-GPU plan (48 GiB) comes
        torch.cuda.empty_cache()
        fc.convolve_device_async(full, C5_DIM, d_k, C5_K, devs[0], torch.cuda.current_stream(devs[0]).cuda_stream)
        torch.cuda.synchronize(devs[0])
        scale = float(full.abs().max())
        worst, num, den = 0.0, 0.0, 0.0
        for r in range(P):
            got = slabs[r].to(f"cuda:{devs[0]}")
            want = full[r * nzp * plane:(r * nzp + planes[r]) * plane]
            diff = got - want
            worst = max(worst, float(diff.abs().max()))
            num += float((diff.double() ** 2).sum())
            den += float((want.double() ** 2).sum())
            del got, diff
        rec["max_rel_err_vs_1gpu"] = worst / scale
        rec["rel_l2_vs_1gpu"] = (num / den) ** 0.5
        del full
    except Exception as exc:
        rec["max_rel_err_vs_1gpu"] = None
        rec["parity_error"] = str(exc)[:200]
    del slabs
    fc.release()
    torch.cuda.empty_cache()
    return rec


def bench_slab_c5_e2e(fc, devs):
    """config 5 end to end: the 16 GiB volume in pinned HOST memory through the C ABI"""
    import numpy as np
    import torch
    n = int(np.prod(C5_DIM))
    psf = gaussian_psf(C5_K).reshape(-1)
    host = torch.empty(n, dtype=torch.float32, pin_memory=True)
    q = n // 8
    for i in range(8):
        host[i * q:(i + 1) * q].copy_(torch.rand(q, device=f"cuda:{devs[0]}") * 1000)

    def call():
        if len(devs) == 1:
            fc.convolution3DfftCUDAInPlace(host, C5_DIM, psf, C5_K, devs[0])
        else:
            fc.convolve_slab(host, C5_DIM, psf, C5_K, devs)
    call()
    t0 = time.perf_counter()
    reps = 2
    for _ in range(reps):
        call()
    dt = (time.perf_counter() - t0) / reps
    rec = {"ms_per_step": dt * 1e3, "value": n / dt / 1e6, "unit": UNIT, "h2d_plus_d2h_GBps": round(2.0 * n * 4 / dt / 1e9, 1),
           "checksum": float(host[n // 3]),
           "api": ("convolution3DfftCUDAInPlace" if len(devs) == 1 else "fcb200_convolve_slab") + "(host pinned volume), wall clock"}
    del host
    fc.release()
    torch.cuda.empty_cache()
    return rec


def sharded_records(fc, n_gpus):
    """rank 0 only; the other ranks of a torchrun launch wait on the host (gloo) meanwhile"""
    devs = list(range(n_gpus))
    out = {}
    for name, fn in (("pcie_ceiling", lambda: pcie_ceiling(devs)),
                     ("batch_c4", lambda: bench_batch_c4(fc, devs, out.get("pcie_ceiling"))),
                     ("slab", lambda: bench_slab_c5(fc, devs)),
                     ("slab_e2e", lambda: bench_slab_c5_e2e(fc, devs))):
        if os.environ.get("FCB200_BENCH_SKIP_" + name.upper()):
            continue
        try:
            fc.release()
            out[name] = fn()
        except Exception as exc:      # an extra must never take the benchmark line down
            out[name] = {"error": str(exc)[:300]}
            try:
                fc.release()
            except Exception:
                pass
    if isinstance(out.get("pcie_ceiling"), dict) and "error" in out["pcie_ceiling"]:
        out["pcie_ceiling"] = None
    return out


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

    # the same loop with the PSF spectrum MATERIALISED (the default call derives it on the fly inside the fused z kernel
    # for this shape, fc_api.cu: prepare_psf): the HBM-bound form of the fused pass, reported next to the default one
    otf_before = os.environ.get("FCB200_OTF_INPLACE")
    os.environ["FCB200_OTF_INPLACE"] = "0"
    for _ in range(2):
        step()
    torch.cuda.synchronize()
    fc.profile_enable(True)
    fc.profile_read()
    for _ in range(args.steps):
        step()
    torch.cuda.synchronize()
    prof_mat = fc.profile_read()
    fc.profile_enable(False)
    if otf_before is None:
        del os.environ["FCB200_OTF_INPLACE"]
    else:
        os.environ["FCB200_OTF_INPLACE"] = otf_before
    for _ in range(2):
        step()
    torch.cuda.synchronize()

    g_xc = IM_DIM[0] // 2 + 1
    N, Nc = n, IM_DIM[2] * IM_DIM[1] * g_xc
    # algorithmic bytes of the five passes of the image path (the fused z pass reads the spectrum and H and
    # writes the spectrum)
    alg_bytes = {"x_fwd": 4 * N + 8 * Nc, "y_fwd": 16 * Nc, "z_fused": 24 * Nc, "y_inv": 16 * Nc,
                 "x_inv": 8 * Nc + 4 * N}
    peak, peak_src = measured_peak()

    def per_pass(profile):
        out = {}
        for name, (ms, cnt) in profile.items():
            if cnt:
                avg = ms / cnt
                entry = {"ms": round(avg, 4)}
                if name in alg_bytes:
                    entry["GBps"] = round(alg_bytes[name] / (avg * 1e-3) / 1e9, 1)
                    entry["frac"] = round(entry["GBps"] / peak, 4)
                out[name] = entry
        return out

    passes, passes_mat = per_pass(prof), per_pass(prof_mat)
    on_the_fly = "psf_z" not in passes     # no PSF z pass ran: H was derived inside the fused z kernel
    dom = max((k for k in passes if k in alg_bytes), key=lambda k: passes[k]["ms"])
    roofline = {"bound": "hbm", "kernel": dom, "achieved": passes[dom]["GBps"], "peak": peak, "unit": "GB/s",
                "frac": passes[dom]["frac"],
                "traffic": ncu_traffic("z_fused_otf" if (dom == "z_fused" and on_the_fly) else dom), "peak_source": peak_src,
                "algorithmic_bytes_per_launch": alg_bytes[dom], "passes": passes,
                "image_path_bytes": 8 * N + 72 * Nc,
                "image_path_frac": round((8 * N + 72 * Nc) / (sum(passes[k]["ms"] for k in alg_bytes) * 1e-3) / 1e9 / peak, 4)}
    if on_the_fly:
        # the default fused z kernel does NOT read H (it derives its H tile from 16 PSF planes in shared memory): it moves
        # 16 Nc bytes and pays for that with butterflies, i.e. it is no longer a pure HBM-bound kernel.  `achieved` above is
        # SURVEY 8(d)'s algorithmic figure (24 Nc) over its time; these are the bytes it really has to move, and the
        # materialised form of the same pass (FCB200_OTF_INPLACE=0), which is the HBM-bound one
        zf = passes["z_fused"]["ms"]
        roofline["z_fused_on_the_fly"] = {
            "bytes_by_design": 16 * Nc, "GBps_by_design": round(16 * Nc / (zf * 1e-3) / 1e9, 1),
            "frac_by_design": round(16 * Nc / (zf * 1e-3) / 1e9 / peak, 4),
            "note": "H derived in the kernel from a 16-plane PSF window: less traffic than the algorithmic 24 Nc, more "
                    "arithmetic; faster than the materialised form below"}
        roofline["materialised"] = {"passes": passes_mat,
                                    "z_fused_frac": passes_mat.get("z_fused", {}).get("frac"),
                                    "z_fused_traffic": ncu_traffic("z_fused")}

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

    # ---- the same call with PAGEABLE buffers (numpy): what a JNA caller hands over, and what the reference arm uses
    p_im = im_host.copy()
    for _ in range(2):
        fc.convolution3DfftCUDAInPlace(p_im, IM_DIM, psf, K_DIM, dev)
    barrier_sync(world)
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        fc.convolution3DfftCUDAInPlace(p_im, IM_DIM, psf, K_DIM, dev)
        checksum_p = float(p_im[0])
    barrier_sync(world)
    e2e_p_ms = max_over_ranks((time.perf_counter() - t0) * 1e3, world, device) / e2e_steps
    e2e_pageable = {"value": world * n / (e2e_p_ms * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": e2e_p_ms,
                    "vs_pinned": round(e2e_p_ms / e2e_ms, 3), "checksum": checksum_p,
                    "api": "convolution3DfftCUDAInPlace(host pageable numpy buffers): staged through pinned slots by "
                           "the library's copy threads"}
    del p_im

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

    # ---- the sharded configurations (rank 0 drives all N GPUs in one process; the others wait on the host)
    sharded = {}
    if not args.no_sharded:
        del d_im, h_im
        torch.cuda.empty_cache()
        if world > 1:
            import torch.distributed as dist
            barrier_sync(world)
            if rank == 0:
                sharded = sharded_records(fc, world)
            dist.barrier(group=HOST_GROUP)
        else:
            sharded = sharded_records(fc, 1)
    line = None
    if rank == 0:
        cpu = cpu_baseline_sample() if (world == 1 and not args.no_cpu_baseline) else None
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": max(args.warmup, 3), "ms_per_step": ms_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": dict(CONFIG), "parallelism": f"independent tiles, one per GPU x{world}",
            "e2e": e2e, "e2e_pageable": e2e_pageable, "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline,
            "e2e_batch": e2e_batch, "padded": padded,
            "savememory": {"value": world * n / (ms_sm * 1e-3) / 1e6, "unit": UNIT, "ms_per_step": ms_sm,
                           "api": "convolution3DfftCUDAInPlaceSaveMemory path (device-resident): PSF spectrum "
                                  "derived on the fly in the fused z kernel, no image-sized PSF buffer"},
        }
        if cpu is not None:
            line["cpu_baseline"] = cpu
        line.update(sharded)
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
            "config": dict(CONFIG), "parallelism": "one GPU (the reference has no multi-GPU path: rank 0 only)"}
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
        # the DEVICE work of the same call with device-resident inputs (oracle/ref_device_harness.cu: the reference's
        # own kernels and cuFFT plans under CUDA events), for the like-for-like comparison with our `value`
        try:
            dev_so = os.path.join(ROOT, "oracle", "_ref", "libref_device.so")
            h = ctypes.CDLL(dev_so)
            parts = (ctypes.c_float * 4)()
            rc = h.ref_device_time(idim, kdim, 10, 3, parts)
            if rc != 0:
                raise RuntimeError(f"ref_device_time rc={rc}")
            tot = float(sum(parts))
            line["reference_device_ms"] = tot
            line["reference_device"] = {"ms_per_step": tot, "value": n / tot / 1e3, "unit": UNIT,
                                        "parts_ms": {"psf_pad_shift": parts[0], "r2c_image+r2c_psf": parts[1],
                                                     "modulateAndNormalize": parts[2], "c2r": parts[3]},
                                        "how": "oracle/ref_device_harness.cu: reference kernels + its cuFFT plans, inputs "
                                               "resident in HBM, plans created once, the 131072 row copies of "
                                               ":474-486 issued as one cudaMemcpy2DAsync (all in the reference's favour)"}
        except Exception as exc:
            line["reference_device_ms"] = None
            line["reference_device"] = {"error": str(exc)[:200]}
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
    ap.add_argument("--no-sharded", action="store_true", help="skip the config 4 / config 5 sub-records")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
