"""In-library padding (fcb200_convolve_padded, include/fcb200_ext.h) through the C ABI.

The reference leaves padding to its callers (src/convolution3Dfft.h:39, :54); its tests pad on the host with
zero_padd::insert_at_offsets (tests/padd_utils.h:99-171), call convolution3DfftCUDAInPlace on the padded grid and
read the interior back (tests/test_fixtures.hpp:254-268).  The padded entry point must return exactly what that
sequence returns: checked against (1) the oracle, (2) the product's own InPlace on a host-padded volume
(bit-identical: same kernels, same grid), (3) the reference build driven the way its tests drive it, and (4) the
direct convolution of the reference's test suite for the cubic cases where the placement is a centred PSF.
Tolerance (north_star): max|err| <= 1e-4 max|out|, relative L2 <= 1e-5."""
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


def host_pad(im, imDim, kDim, mode, policy):
    """numpy version of what the library does on the device; returns (padded flat, padded dims, offsets)"""
    d0, d1, d2 = imDim
    p0, p1, p2 = fo.padded_extents(imDim, kDim, policy)
    o0, o1, o2 = fo.zero_padd_offsets(kDim)
    I3 = np.asarray(im, np.float32).reshape(d2, d1, d0)
    widths = ((o2, p2 - d2 - o2), (o1, p1 - d1 - o1), (o0, p0 - d0 - o0))
    P3 = np.pad(I3, widths, mode="constant" if mode == 0 else "reflect")
    return np.ascontiguousarray(P3).reshape(-1), (p0, p1, p2), (o0, o1, o2)


CASES = [  # imDim, kernelDim
    ((64, 64, 64), (3, 3, 3)),          # BASELINE config 1
    ((32, 32, 32), (7, 7, 7)),
    ((48, 40, 24), (5, 7, 9)),          # non-cubic: the placement quirk is live on the padded grid
    ((30, 20, 50), (4, 6, 2)),          # even kernel extents
    ((17, 5, 9), (4, 3, 5)),
    ((100, 36, 20), (9, 9, 5)),
    ((16, 16, 16), (15, 15, 15)),       # halo almost as wide as the volume
    ((120, 20, 12), (9, 3, 3)),         # padded rows of 128: InPlace takes the row-wise x kernels, the fused
    ((370, 12, 6), (15, 3, 3)),         # padded path the tiled ones (384 likewise) -> round-off, not bits
    ((33, 9, 7), (5, 5, 5)),            # odd padded row length (37)
]
ROW_WISE_NX = (128, 256, 384, 512, 1024, 2048)


@pytest.mark.parametrize("mode", [0, 1], ids=["zero", "mirror"])
@pytest.mark.parametrize("policy", [0, 1], ids=["exact", "smooth"])
@pytest.mark.parametrize("imDim,kDim", CASES, ids=lambda v: "x".join(map(str, v)))
def test_padded_matches_oracle_and_host_padded_inplace(fc, dev, imDim, kDim, mode, policy):
    rng = np.random.default_rng(sum(imDim) * 7 + sum(kDim) + mode)
    im = (rng.random(int(np.prod(imDim)), dtype=np.float32) * 1000).astype(np.float32)
    k = rng.random(int(np.prod(kDim)), dtype=np.float32)
    k /= k.sum()
    assert fc.padded_extents(imDim, kDim, policy) == fo.padded_extents(imDim, kDim, policy)

    got = im.copy()
    fc.convolve_padded(got, imDim, k.copy(), kDim, dev, mode=mode, policy=policy)
    check(got, fo.convolve_padded_ref(im, imDim, k, kDim, mode, policy))

    # same kernels on the same grid: bit-identical to InPlace on the host-padded volume, cropped
    padded, pDim, off = host_pad(im, imDim, kDim, mode, policy)
    fc.convolution3DfftCUDAInPlace(padded, pDim, k.copy(), kDim, dev)
    want = fo.crop(padded.reshape(pDim[2], pDim[1], pDim[0]), off[::-1], imDim[::-1]).reshape(-1)
    if pDim[0] in ROW_WISE_NX:
        check(got, want, max_rel=2e-6, l2_rel=1e-6)
    else:
        assert np.array_equal(got, want)


@pytest.mark.parametrize("imDim,kDim", [((64, 64, 64), (3, 3, 3)), ((48, 40, 24), (5, 7, 9)), ((40, 40, 40), (9, 9, 9))],
                         ids=lambda v: "x".join(map(str, v)))
def test_padded_matches_reference_build_driven_like_its_tests(fc, dev, reflib, imDim, kDim):
    import reflib as rl
    rng = np.random.default_rng(11)
    im = rng.random(int(np.prod(imDim)), dtype=np.float32)
    k = rng.random(int(np.prod(kDim)), dtype=np.float32)
    for policy in (0, 1):
        padded, pDim, off = host_pad(im, imDim, kDim, 0, policy)
        ref = rl.convolve_inplace(padded, pDim, k, kDim, dev)
        want = fo.crop(ref.reshape(pDim[2], pDim[1], pDim[0]), off[::-1], imDim[::-1]).reshape(-1)
        got = im.copy()
        fc.convolve_padded(got, imDim, k.copy(), kDim, dev, mode=0, policy=policy)
        check(got, want)


def test_zero_padded_cubic_equals_direct_convolution(fc, dev):
    """the reference's own acceptance pattern (tests/test_gpu_convolve.cpp with tests/test_fixtures.hpp:254-268):
    on a cubic padded grid the FFT result inside the sub-view equals the CPU `convolve` of the padded stack"""
    n, kk = 16, 5
    rng = np.random.default_rng(3)
    im3 = rng.random((n, n, n), dtype=np.float32)
    k3 = rng.random((kk, kk, kk), dtype=np.float32)
    padded, off = fo.zero_padd(im3, k3.shape)
    want = fo.crop(fo.direct_convolve(padded, k3, off), off, im3.shape)
    got = im3.reshape(-1).copy()
    fc.convolve_padded(got, (n, n, n), k3.reshape(-1).copy(), (kk, kk, kk), dev, mode=0, policy=0)
    check(got, want)


def test_padded_device_pointers_and_pinned_chunks(fc, dev):
    """device-resident call (stream-ordered) and the z-chunked pinned path give the pageable-path result"""
    import torch
    imDim, kDim = (160, 144, 136), (9, 7, 11)      # 12.5 MB: below the chunking threshold (one chunk per 32 MiB, >= 2)
    big, kbig = (448, 320, 168), (9, 9, 7)         # 92 MB unpadded -> pinned calls travel in 2 z chunks
    rng = np.random.default_rng(5)
    for (idim, kdim) in ((imDim, kDim), (big, kbig)):
        im = rng.random(int(np.prod(idim)), dtype=np.float32)
        k = rng.random(int(np.prod(kdim)), dtype=np.float32)
        k /= k.sum()
        for mode in (0, 1):
            base = im.copy()
            fc.convolve_padded(base, idim, k.copy(), kdim, dev, mode=mode, policy=1)            # pageable
            check(base, fo.convolve_padded_ref(im, idim, k, kdim, mode, 1))
            pinned = torch.from_numpy(im.copy()).pin_memory()
            fc.convolve_padded(pinned, idim, k.copy(), kdim, dev, mode=mode, policy=1)
            assert np.array_equal(pinned.numpy(), base)
            d_im = torch.from_numpy(im).to(f"cuda:{dev}")
            d_k = torch.from_numpy(k).to(f"cuda:{dev}")
            torch.cuda.synchronize(dev)
            fc.convolve_padded_device_async(d_im, idim, d_k, kdim, dev, mode=mode, policy=1,
                                            stream=torch.cuda.current_stream(dev).cuda_stream)
            torch.cuda.synchronize(dev)
            assert np.array_equal(d_im.cpu().numpy(), base)
            d_im2 = torch.from_numpy(im).to(f"cuda:{dev}")
            fc.convolve_padded(d_im2, idim, k.copy(), kdim, dev, mode=mode, policy=1)           # device image, host PSF
            assert np.array_equal(d_im2.cpu().numpy(), base)


def test_padded_rejects_bad_arguments(fc, dev):
    im = np.zeros(8 * 8 * 8, np.float32)
    k = np.ones(27, np.float32)
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_padded(im, (8, 8, 8), k, (3, 3, 3), dev, mode=2)
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_padded(im, (8, 8, 8), k, (3, 3, 3), dev, policy=5)
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.convolve_padded(im, (8, 0, 8), k, (3, 3, 3), dev)


@pytest.mark.parametrize("mode", [0, 1], ids=["zero", "mirror"])
def test_padded_batch_equals_separate_padded_calls(fc, dev, mode):
    """fcb200_convolve_batch_padded (blocks + halo in one call, transfers pipelined) == n separate padded calls"""
    import torch
    imDim, kDim, nblocks = (96, 80, 72), (9, 7, 11), 5
    rng = np.random.default_rng(21)
    k = rng.random(int(np.prod(kDim)), dtype=np.float32)
    k /= k.sum()
    blocks = [(rng.random(int(np.prod(imDim)), dtype=np.float32) * 100).astype(np.float32) for _ in range(nblocks)]
    want = []
    for b in blocks:
        w = b.copy()
        fc.convolve_padded(w, imDim, k.copy(), kDim, dev, mode=mode, policy=1)
        want.append(w)
    check(want[0], fo.convolve_padded_ref(blocks[0], imDim, k, kDim, mode, 1))
    pageable = [b.copy() for b in blocks]
    fc.convolve_batch_padded(pageable, imDim, k.copy(), kDim, dev, mode=mode, policy=1)
    pinned = [torch.from_numpy(b.copy()).pin_memory() for b in blocks]
    fc.convolve_batch_padded(pinned, imDim, k.copy(), kDim, dev, mode=mode, policy=1)
    device = [torch.from_numpy(b).to(f"cuda:{dev}") for b in blocks]
    fc.convolve_batch_padded(device, imDim, k.copy(), kDim, dev, mode=mode, policy=1)
    for i in range(nblocks):
        assert np.array_equal(pageable[i], want[i])
        assert np.array_equal(pinned[i].numpy(), want[i])
        assert np.array_equal(device[i].cpu().numpy(), want[i])
    fc.convolve_batch_padded([], imDim, k.copy(), kDim, dev)        # empty batch: nothing to do
