"""Host-pointer pipeline: pageable staging, the batch entry point (BASELINE config 4 pattern: many blocks, one
PSF) and the PSF-spectrum cache must give exactly the numbers of plain single calls on device-resident data."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gaussian_psf(kDim):
    ax = [np.exp(-0.5 * ((np.arange(k) - k // 2) / (k / 6.0)) ** 2) for k in kDim]
    psf = ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]
    return (psf / psf.sum()).astype(np.float32)


def device_result(fc, dev, im, imDim, k, kDim):
    """reference result for bit-comparisons: device-resident call (no staging, no PSF cache involved)"""
    import torch
    d_im = torch.from_numpy(im).to(f"cuda:{dev}")
    d_k = torch.from_numpy(k).to(f"cuda:{dev}")
    fc.convolve_device_async(d_im, imDim, d_k, kDim, dev, torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    return d_im.cpu().numpy()


def test_pageable_multi_chunk_call_matches_device_call(fc, dev):
    imDim, kDim = (256, 200, 130), (9, 7, 5)        # 26.6 MB: four 8 MiB staging chunks, last one ragged
    rng = np.random.default_rng(3)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 1000).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    want = device_result(fc, dev, im, imDim, k, kDim)
    got = im.copy()
    fc.convolution3DfftCUDAInPlace(got, imDim, k, kDim, dev)
    assert np.array_equal(got, want)


@pytest.mark.parametrize("kind", ["pageable", "pinned", "device"])
@pytest.mark.parametrize("n", [1, 2, 7])
def test_batch_equals_single_calls(fc, dev, kind, n):
    import torch
    imDim, kDim = (96, 80, 64), (7, 5, 9)
    rng = np.random.default_rng(100 + n)
    k = gaussian_psf(kDim).reshape(-1)
    blocks = [(rng.random(int(np.prod(imDim)), dtype=np.float32) * 1000).astype(np.float32) for _ in range(n)]
    want = [device_result(fc, dev, b, imDim, k, kDim) for b in blocks]
    if kind == "pageable":
        ims = [b.copy() for b in blocks]
    elif kind == "pinned":
        ims = [torch.from_numpy(b).pin_memory() for b in blocks]
    else:
        ims = [torch.from_numpy(b).to(f"cuda:{dev}") for b in blocks]
    fc.convolve_batch(ims, imDim, k, kDim, dev)
    for got, w in zip(ims, want):
        g = got if isinstance(got, np.ndarray) else got.cpu().numpy()
        assert np.array_equal(g, w)


def test_large_pageable_batch_overlaps_and_matches(fc, dev):
    """blocks larger than one staging chunk; more blocks than ring buffers"""
    imDim, kDim = (192, 160, 128), (5, 5, 5)       # 15.7 MB per block
    rng = np.random.default_rng(9)
    k = gaussian_psf(kDim).reshape(-1)
    blocks = [(rng.random(int(np.prod(imDim)), dtype=np.float32) * 100).astype(np.float32) for _ in range(5)]
    want = [device_result(fc, dev, b, imDim, k, kDim) for b in blocks]
    ims = [b.copy() for b in blocks]
    fc.convolve_batch(ims, imDim, k, kDim, dev)
    for g, w in zip(ims, want):
        assert np.array_equal(g, w)


def test_psf_cache_is_transparent(fc, dev):
    """same taps -> cached spectrum, changed taps (same shape) -> recomputed, other shape -> recomputed"""
    imDim, kDim = (64, 64, 32), (5, 7, 3)
    rng = np.random.default_rng(21)
    im = rng.random(int(np.prod(imDim)), dtype=np.float32)
    k1 = gaussian_psf(kDim).reshape(-1)
    k2 = k1.copy()
    k2[7] *= 3.0
    for k in (k1, k1, k2, k1, k2, k2):
        want = device_result(fc, dev, im, imDim, k, kDim)
        got = im.copy()
        fc.convolution3DfftCUDAInPlace(got, imDim, k, kDim, dev)
        assert np.array_equal(got, want)
    kDim3 = (3, 7, 5)                                  # same number of taps, other shape
    k3 = gaussian_psf(kDim3).reshape(-1)
    want = device_result(fc, dev, im, imDim, k3, kDim3)
    got = im.copy()
    fc.convolution3DfftCUDAInPlace(got, imDim, k3, kDim3, dev)
    assert np.array_equal(got, want)
    # a device-pointer call in between overwrites the spectrum: the cache must notice
    other = gaussian_psf(kDim).reshape(-1)[::-1].copy()
    device_result(fc, dev, im, imDim, other, kDim)
    want = device_result(fc, dev, im, imDim, k3, kDim3)
    got = im.copy()
    fc.convolution3DfftCUDAInPlace(got, imDim, k3, kDim3, dev)
    assert np.array_equal(got, want)


def test_batch_rejects_mixed_pointer_kinds(fc, dev):
    import torch
    imDim, kDim = (32, 32, 32), (3, 3, 3)
    k = gaussian_psf(kDim).reshape(-1)
    a = np.zeros(32 ** 3, np.float32)
    b = torch.zeros(32 ** 3, device=f"cuda:{dev}")
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_batch([a, b], imDim, k, kDim, dev)


@pytest.mark.parametrize("imDim", [(256, 256, 261), (512, 512, 256)])
def test_pinned_call_with_overlapped_chunks_matches_device_call(fc, dev, imDim):
    """pinned host images >= 64 MB travel in z chunks whose x/y passes overlap the copies (fc_api.cu);
    the numbers must not depend on the chunking (ragged last chunk included)"""
    import torch
    kDim = (9, 7, 11)
    rng = np.random.default_rng(4)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 1000).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    want = device_result(fc, dev, im, imDim, k, kDim)
    h = torch.from_numpy(im).pin_memory()
    fc.convolution3DfftCUDAInPlace(h, imDim, k, kDim, dev)
    assert np.array_equal(h.numpy(), want)
    h2 = torch.from_numpy(im).pin_memory()
    fc.convolution3DfftCUDAInPlaceSaveMemory(h2, imDim, k, kDim, dev)
    scale = np.abs(want).max()
    assert np.abs(h2.numpy() - want).max() <= 1e-5 * scale
