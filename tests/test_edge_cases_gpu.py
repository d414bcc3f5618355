"""Edge cases through the C ABI: degenerate extents, kernel as large as the image, legacy entry points,
concurrent callers, failures that must raise instead of exiting.  Tolerance 1e-4 / 1e-5 as everywhere."""
import threading

import numpy as np
import pytest

from oracle import fc_oracle as fo

pytestmark = pytest.mark.gpu


def check(got, want, max_rel=1e-4, l2_rel=1e-5):
    got = np.asarray(got, np.float64).ravel()
    want = np.asarray(want, np.float64).ravel()
    scale = max(np.abs(want).max(), 1e-30)
    assert np.abs(got - want).max() <= max_rel * scale
    assert np.linalg.norm(got - want) <= l2_rel * max(np.linalg.norm(want), 1e-30)


DEGENERATE = [  # imDim, kernelDim
    ((1, 1, 1), (1, 1, 1)), ((8, 1, 1), (3, 1, 1)), ((1, 8, 1), (1, 3, 1)), ((1, 1, 8), (1, 1, 3)),
    ((2, 3, 1), (1, 1, 1)), ((7, 1, 5), (3, 1, 3)), ((16, 16, 1), (5, 5, 1)), ((3, 5, 7), (3, 5, 7)),
    ((8, 8, 8), (8, 8, 8)), ((12, 10, 6), (12, 1, 1)), ((5, 4, 3), (1, 4, 1)), ((64, 2, 2), (9, 1, 1)),
    ((2, 2, 64), (1, 1, 9)), ((31, 29, 23), (5, 3, 7)), ((97, 4, 4), (3, 3, 3)),
]


@pytest.mark.parametrize("imDim,kDim", DEGENERATE, ids=lambda v: "x".join(map(str, v)))
def test_degenerate_and_full_size_kernels(fc, dev, imDim, kDim):
    rng = np.random.default_rng(sum(imDim) * 31 + sum(kDim))
    im = rng.random(int(np.prod(imDim)), dtype=np.float32) + 0.5
    k = rng.random(int(np.prod(kDim)), dtype=np.float32)
    want = fo.convolve_inplace_ref(im, imDim, k, kDim)
    for entry in (fc.convolution3DfftCUDAInPlace, fc.convolution3DfftCUDAInPlaceSaveMemory):
        got = im.copy()
        entry(got, imDim, k.copy(), kDim, dev)
        check(got, want)


def test_legacy_out_of_place_entry_points(fc, dev):
    """convolution3DfftCUDA / _test (reference src/convolution3Dfft.cu:216-391): legacy convention, imDim[2]
    fastest for image, kernel and placement, i.e. a plain centred circular convolution of [d0][d1][d2] arrays."""
    rng = np.random.default_rng(8)
    imDim, kDim = (12, 20, 16), (3, 5, 7)
    im = rng.random(imDim).astype(np.float32)
    k = rng.random(kDim).astype(np.float32)
    S = np.zeros(imDim)
    for a in range(kDim[0]):
        for b in range(kDim[1]):
            for c in range(kDim[2]):
                S[(a - kDim[0] // 2) % imDim[0], (b - kDim[1] // 2) % imDim[1], (c - kDim[2] // 2) % imDim[2]] = k[a, b, c]
    want = np.fft.irfftn(np.fft.rfftn(im.astype(np.float64)) * np.fft.rfftn(S), s=imDim, axes=(0, 1, 2))
    before = im.copy()
    got = fc.convolution3DfftCUDA(im.reshape(-1), imDim, k.reshape(-1), kDim, dev)
    assert np.array_equal(im, before)                      # out of place: the input is untouched
    check(got, want)
    got2 = fc.convolution3DfftCUDA_test(im.reshape(-1), imDim, S.astype(np.float32).reshape(-1), dev)
    check(got2, want)


def test_concurrent_host_threads(fc, dev):
    """the reference is stateless and thread-safe by construction; the plan cache must keep that property"""
    shapes = [((64, 48, 40), (5, 5, 5)), ((64, 48, 40), (7, 3, 5)), ((96, 32, 16), (3, 3, 3)), ((30, 20, 50), (4, 6, 2))]
    rng = np.random.default_rng(0)
    jobs = []
    for imDim, kDim in shapes * 2:
        im = rng.random(int(np.prod(imDim)), dtype=np.float32)
        k = rng.random(int(np.prod(kDim)), dtype=np.float32)
        jobs.append((imDim, kDim, im, k, fo.convolve_inplace_ref(im, imDim, k, kDim)))
    errors = []

    def work(job):
        imDim, kDim, im, k, want = job
        try:
            for _ in range(3):
                got = im.copy()
                fc.convolution3DfftCUDAInPlace(got, imDim, k.copy(), kDim, dev)
                check(got, want)
        except Exception as exc:          # noqa: BLE001
            errors.append(repr(exc))

    threads = [threading.Thread(target=work, args=(j,)) for j in jobs]
    for t in threads:
        t.start()
    for t in threads:
        t.join()
    assert not errors, errors


def test_out_of_memory_raises_and_library_stays_usable(fc, dev):
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolution3DfftCUDAInPlace(np.zeros(8, np.float32), (4096, 4096, 4096), np.ones(1, np.float32), (1, 1, 1), dev)
    im = np.arange(64, dtype=np.float32)
    k = np.zeros(27, np.float32)
    k[13] = 1
    got = im.copy()
    fc.convolution3DfftCUDAInPlace(got, (4, 4, 4), k, (3, 3, 3), dev)
    check(got, im)


def test_release_frees_the_plan_cache(fc, dev):
    import torch
    im = np.ones(128 * 128 * 64, np.float32)
    fc.convolution3DfftCUDAInPlace(im, (128, 128, 64), np.ones(27, np.float32) / 27, (3, 3, 3), dev)
    free_before, _ = torch.cuda.mem_get_info(dev)
    fc.release()
    free_after, _ = torch.cuda.mem_get_info(dev)
    assert free_after > free_before
    fc.convolution3DfftCUDAInPlace(im, (128, 128, 64), np.ones(27, np.float32) / 27, (3, 3, 3), dev)   # rebuilds


def test_async_calls_of_one_shape_on_two_streams_do_not_race(fc, dev):
    """the stream-ordered entry point shares one workspace per (device, shape): a call on a second stream must be
    ordered after the call still running on the first (per-plan "workspace busy" event), not race with it"""
    import torch
    imDim, kDim = (256, 256, 128), (9, 9, 9)
    n = int(np.prod(imDim))
    device = torch.device(f"cuda:{dev}")
    rng = np.random.default_rng(31)
    k = torch.from_numpy(rng.random(int(np.prod(kDim)), dtype=np.float32)).to(device)
    ims = [torch.rand(n, device=device) * 100 for _ in range(4)]
    want = []
    for im in ims:
        w = im.clone()
        fc.convolve_device_async(w, imDim, k, kDim, dev, torch.cuda.current_stream().cuda_stream)
        torch.cuda.synchronize()
        want.append(w)
    streams = [torch.cuda.Stream(device=device) for _ in range(2)]
    torch.cuda.synchronize()
    got = [im.clone() for im in ims]
    for rep in range(5):
        for i, g in enumerate(got):
            g.copy_(ims[i])
        torch.cuda.synchronize()
        for i, g in enumerate(got):
            fc.convolve_device_async(g, imDim, k, kDim, dev, streams[i % 2].cuda_stream)
        torch.cuda.synchronize()
        for g, w in zip(got, want):
            assert torch.equal(g, w)


def test_misaligned_device_pointers_are_rejected_not_faulted(fc, dev):
    import torch
    imDim, kDim = (64, 64, 8), (3, 3, 3)
    n = int(np.prod(imDim))
    buf = torch.zeros(n + 1, device=f"cuda:{dev}")
    k = torch.ones(27, device=f"cuda:{dev}")
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_batch([buf[1:]], imDim, k, kDim, dev)         # 4-byte aligned slice
    fc.convolve_batch([buf[:n]], imDim, k, kDim, dev)             # the context is still healthy
    torch.cuda.synchronize()


@pytest.mark.parametrize("plan", ["y300=12.5.5", "z300=4.3.5.5,y300=4.3.5.5", "x150=2.3.5.5", "y300=20.15,z300=15.20"])
def test_plan_override_changes_the_radix_sequence_not_the_result(fc, dev, monkeypatch, plan):
    """FCB200_PLAN (tuning runs) picks the radix sequence of an axis and length; sequences without a compile-time kernel
    run on the run-time-radix kernels -- the result must be the planner's own to fp32 round-off"""
    imDim, kDim = (300, 300, 300), (5, 3, 7)
    rng = np.random.default_rng(77)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 1000).astype(np.float32)
    k = rng.random(int(np.prod(kDim)), dtype=np.float32)
    k /= k.sum()
    fc.release()
    want = im.copy()
    fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)
    assert fc.plan_radices(300, 0)[0] == [20, 15] and fc.plan_radices(150, 2)[0] == [10, 15]
    fc.release()
    monkeypatch.setenv("FCB200_PLAN", plan)
    try:
        first = plan.split(",")[0]
        style = {"y": 0, "z": 1, "x": 2}[first[0]]
        L, rad = first[1:].split("=")
        assert fc.plan_radices(int(L), style)[0] == [int(r) for r in rad.split(".")]
        got = im.copy()
        fc.convolution3DfftCUDAInPlace(got, imDim, k, kDim, dev)
    finally:
        monkeypatch.delenv("FCB200_PLAN")
        fc.release()
    check(got, want, max_rel=2e-6, l2_rel=1e-6)


def test_run_time_radix_tma_pipeline_equals_the_one_tile_kernels(fc, dev, monkeypatch):
    """plain passes of long lengths without a compile-time plan: FCB200_TMA_DYN=0 routes them through the one-tile-per-CTA
    run-time-radix kernels -- same radix sequence, same result to fp32 round-off"""
    import torch
    imDim, kDim = (64, 400, 480), (5, 5, 5)     # y = 400 = (16,5,5), z = 480 = (8,4,15) (fused: not on the pipeline)
    n = int(np.prod(imDim))
    g = torch.Generator(device=f"cuda:{dev}")
    g.manual_seed(11)
    base = torch.rand(n, device=f"cuda:{dev}", generator=g) * 1000
    d_k = torch.rand(int(np.prod(kDim)), device=f"cuda:{dev}", generator=g)
    d_k /= d_k.sum()
    x = base.clone()
    fc.convolve_device_async(x, imDim, d_k, kDim, dev, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    # FCB200_TMA_DYN is read once per process: the comparison runs in a child process
    import os
    import subprocess
    import sys
    import tempfile
    with tempfile.TemporaryDirectory() as tmp:
        np.save(os.path.join(tmp, "im.npy"), base.cpu().numpy())
        np.save(os.path.join(tmp, "k.npy"), d_k.cpu().numpy())
        code = (
            "import sys, numpy as np, torch\n"
            f"sys.path.insert(0, {os.path.dirname(os.path.dirname(os.path.abspath(__file__)))!r})\n"
            "import fourierconvolutioncudalib_b200 as fc\n"
            f"im = torch.from_numpy(np.load({os.path.join(tmp, 'im.npy')!r})).cuda({dev})\n"
            f"k = torch.from_numpy(np.load({os.path.join(tmp, 'k.npy')!r})).cuda({dev})\n"
            f"fc.convolve_device_async(im, {imDim!r}, k, {kDim!r}, {dev}, torch.cuda.current_stream().cuda_stream)\n"
            "torch.cuda.synchronize()\n"
            f"np.save({os.path.join(tmp, 'out.npy')!r}, im.cpu().numpy())\n")
        env = dict(os.environ, FCB200_TMA_DYN="0")
        subprocess.run([sys.executable, "-c", code], check=True, env=env)
        other = np.load(os.path.join(tmp, "out.npy"))
    check(x.cpu().numpy(), other, max_rel=2e-6, l2_rel=1e-6)
