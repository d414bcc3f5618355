"""Single-process multi-GPU entry points (include/fcb200_ext.h: fcb200_convolve_slab, fcb200_convolve_slab_device,
fcb200_convolve_batch_multi, SaveMemory routing): results must equal the single-device convolution3DfftCUDAInPlace
result (north_star: slab / tile sharding parity, tolerance 1e-4 / 1e-5).  On a one-GPU box the ranks share device 0
(emulated ranks: same code, same kernels, peer stores into the same device)."""
import os
import shutil
import subprocess

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gaussian_psf(kDim):
    ax = [np.exp(-0.5 * ((np.arange(k) - k // 2) / (k / 6.0)) ** 2) for k in kDim]
    psf = ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]
    return (psf / psf.sum()).astype(np.float32)


def device_list(fc, world):
    have = fc.getNumDevicesCUDA()
    return list(range(world)) if have >= world else [0] * world


def check(got, want):
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-4 * scale
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5


@pytest.mark.parametrize("world", [2, 4])
def test_cpp_caller_of_the_multi_gpu_entry_points(fc, world, tmp_path):
    """tests/cpp/multi_gpu.cpp, built with g++ and linked against the drop-in library"""
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "multi_gpu")
    libdir = os.path.dirname(fc._lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "multi_gpu.cpp"), "-o", exe,
                           "-L", libdir, "-lFourierConvolutionCUDALib", "-Wl,-rpath," + libdir])
    res = subprocess.run([exe, str(world)], capture_output=True, text=True, timeout=600)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failure(s)" in res.stdout


@pytest.mark.parametrize("imDim,kDim,world", [((64, 64, 64), (7, 7, 7), 2), ((128, 96, 64), (9, 5, 7), 4),
                                              ((256, 256, 64), (15, 15, 15), 8), ((64, 2048, 64), (5, 7, 5), 2),
                                              # ragged slabs and non-power-of-two extents
                                              ((64, 70, 45), (5, 5, 5), 2), ((128, 100, 75), (7, 5, 5), 8),
                                              ((72, 135, 90), (5, 5, 5), 4), ((70, 66, 46), (3, 3, 5), 3),
                                              # two-stage plans with fat radices on the split-layout run-time kernels
                                              ((64, 270, 300), (5, 5, 5), 4)])
@pytest.mark.parametrize("kind", ["pageable", "pinned", "device"])
def test_slab_matches_single_device(fc, dev, imDim, kDim, world, kind):
    import torch
    rng = np.random.default_rng(21)
    n = int(np.prod(imDim))
    im = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    want = im.copy()
    fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)
    devs = device_list(fc, world)
    if kind == "pageable":
        got = im.copy()
        fc.convolve_slab(got, imDim, k, kDim, devs)
    elif kind == "pinned":
        t = torch.from_numpy(im.copy()).pin_memory()
        fc.convolve_slab(t, imDim, k, kDim, devs)
        got = t.numpy().copy()
    else:
        nzp, _, planes = fc.slab_partition(imDim, world)
        plane = imDim[0] * imDim[1]
        slabs = [torch.from_numpy(im[r * nzp * plane:(r * nzp + planes[r]) * plane]).to(f"cuda:{devs[r]}")
                 for r in range(world)]
        d_k = torch.from_numpy(k).to(f"cuda:{devs[0]}")
        for _ in range(2):        # twice: the second call waits for the first through the ev_done events
            for r in range(world):
                slabs[r].copy_(torch.from_numpy(im[r * nzp * plane:(r * nzp + planes[r]) * plane]))
            torch.cuda.synchronize()
            fc.convolve_slab_device(slabs, imDim, d_k, kDim, devs)
        got = np.concatenate([s.cpu().numpy() for s in slabs])
        ms = fc.slab_last_timing(imDim, devs)
        assert len(ms) == world and all(m[3] > 0 for m in ms)
    check(got, want)


@pytest.mark.parametrize("exchange", ["0", "1", "2"])
def test_both_forward_exchanges(fc, dev, monkeypatch, exchange):
    """FCB200_SLAB_EXCHANGE=1 (default): the y pass writes the exchange layout locally and the copy engines move the
    blocks, chunk by chunk; =0: the y pass stores straight into the peers' buffers; =2: the peers' fused z pass reads the
    planes out of the producer's exchange buffer (no separate exchange step).  Same result every way."""
    import torch
    monkeypatch.setenv("FCB200_SLAB_EXCHANGE", exchange)
    imDim, kDim, world = (128, 96, 160), (7, 5, 9), 4          # 40 planes per rank: the copy exchange runs in 4 chunks
    devs = device_list(fc, world)
    rng = np.random.default_rng(33)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 1000).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    want = im.copy()
    fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)
    nzp, _, planes = fc.slab_partition(imDim, world)
    plane = imDim[0] * imDim[1]
    slabs = [torch.from_numpy(im[r * nzp * plane:(r * nzp + planes[r]) * plane]).to(f"cuda:{devs[r]}") for r in range(world)]
    d_k = torch.from_numpy(k).to(f"cuda:{devs[0]}")
    for _ in range(2):
        for r in range(world):
            slabs[r].copy_(torch.from_numpy(im[r * nzp * plane:(r * nzp + planes[r]) * plane]))
        torch.cuda.synchronize()
        fc.convolve_slab_device(slabs, imDim, d_k, kDim, devs)
    check(np.concatenate([s.cpu().numpy() for s in slabs]), want)
    t = torch.from_numpy(im.copy()).pin_memory()
    fc.convolve_slab(t, imDim, k, kDim, devs)
    check(t.numpy(), want)


def test_slab_psf_cache_and_new_psf(fc, dev):
    """a second call with the same host PSF may reuse the PSF-spectrum slabs; a different PSF must not"""
    imDim, world = (96, 64, 48), 2
    devs = device_list(fc, world)
    rng = np.random.default_rng(5)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 10).astype(np.float32)
    for kDim in ((5, 5, 5), (5, 5, 5), (7, 3, 9)):
        k = rng.random(int(np.prod(kDim)), dtype=np.float32)
        for _ in range(2):
            want, got = im.copy(), im.copy()
            fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)
            fc.convolve_slab(got, imDim, k, kDim, devs)
            check(got, want)


def test_savememory_routes_to_slab_mode(fc, dev, monkeypatch):
    """FCB200_SLAB=1 makes convolution3DfftCUDAInPlaceSaveMemory spread a host volume over the devices"""
    imDim, kDim = (128, 64, 40), (5, 7, 3)
    rng = np.random.default_rng(6)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 100).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    want = im.copy()
    fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)
    monkeypatch.setenv("FCB200_SLAB", "1")
    if fc.getNumDevicesCUDA() < 2:
        monkeypatch.setenv("FCB200_SLAB_DEVICES", f"{dev},{dev}")
    assert len(fc.slab_devices(imDim, dev)) >= 2
    got = im.copy()
    fc.convolution3DfftCUDAInPlaceSaveMemory(got, imDim, k, kDim, dev)
    check(got, want)
    ms = fc.slab_last_timing(imDim, fc.slab_devices(imDim, dev))     # proves the slab context ran
    assert all(m[3] > 0 for m in ms)
    monkeypatch.setenv("FCB200_SLAB", "0")
    assert fc.slab_devices(imDim, dev) == []


@pytest.mark.parametrize("kind", ["pageable", "pinned"])
def test_batch_multi_matches_separate_calls(fc, dev, kind):
    import torch
    imDim, kDim, nb = (96, 64, 48), (7, 5, 9), 9
    rng = np.random.default_rng(7)
    n = int(np.prod(imDim))
    k = gaussian_psf(kDim).reshape(-1)
    blocks = [(rng.random(n, dtype=np.float32) * 1000).astype(np.float32) for _ in range(nb)]
    want = [b.copy() for b in blocks]
    for w in want:
        fc.convolution3DfftCUDAInPlace(w, imDim, k, kDim, dev)
    have = fc.getNumDevicesCUDA()
    devs = list(range(min(have, 4))) if have > 1 else [dev]
    if kind == "pinned":
        blocks = [torch.from_numpy(b).pin_memory() for b in blocks]
    taken = fc.convolve_batch_multi(blocks, imDim, k, kDim, devs)
    assert sum(taken) == nb and len(taken) == len(devs)
    for b, w in zip(blocks, want):
        got = b.numpy() if kind == "pinned" else b
        assert np.array_equal(got, w)       # same kernels, same order of operations: bit-identical


def test_slab_rejects_bad_arguments(fc, dev):
    im = np.zeros(64 * 64 * 4, np.float32)
    k = np.ones(27, np.float32)
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_slab(im, (64, 64, 4), k, (3, 3, 3), [dev] * 8)       # more ranks than planes
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_slab(im, (64, 64, 4), np.ones(5 * 5 * 5, np.float32), (5, 5, 5), [dev, dev])   # kernel larger than image
