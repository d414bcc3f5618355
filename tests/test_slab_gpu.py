"""Slab-decomposed single-volume path: P emulated ranks on ONE GPU must reproduce the single-GPU
convolution3DfftCUDAInPlace result (north_star: slab mode parity) -- tolerance 1e-4 / 1e-5 as everywhere."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def gaussian_psf(kDim):
    ax = [np.exp(-0.5 * ((np.arange(k) - k // 2) / (k / 6.0)) ** 2) for k in kDim]
    psf = ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]
    return (psf / psf.sum()).astype(np.float32)


@pytest.mark.parametrize("imDim,kDim,world", [((64, 64, 64), (7, 7, 7), 2), ((128, 96, 64), (9, 5, 7), 4),
                                              ((70, 60, 48), (5, 5, 9), 2), ((256, 256, 64), (15, 15, 15), 8),
                                              ((64, 2048, 64), (5, 7, 5), 2),
                                              # ragged slabs: extents not divisible by the number of ranks
                                              ((64, 70, 45), (5, 5, 5), 2), ((64, 90, 50), (5, 5, 9), 4),
                                              ((128, 100, 75), (7, 5, 5), 8), ((64, 64, 75), (5, 5, 5), 8),
                                              ((72, 135, 90), (5, 5, 5), 4)])
def test_emulated_ranks_match_single_gpu(fc, dev, imDim, kDim, world):
    import torch
    from fourierconvolutioncudalib_b200 import slab
    rng = np.random.default_rng(11)
    n = int(np.prod(imDim))
    im = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    want = im.copy()
    fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)

    d_k = torch.from_numpy(k).to(f"cuda:{dev}")
    ex = slab.LocalExchange(world)
    convs = [slab.SlabConvolver(imDim, kDim, r, world, dev, ex) for r in range(world)]
    full = torch.from_numpy(im).to(f"cuda:{dev}")
    slabs = [c.slab_of(full) for c in convs]          # views into `full`: convolved in place
    for c in convs:
        c.prepare_psf(d_k)
    slab.run_lockstep(convs, slabs, ex)
    torch.cuda.synchronize()
    got = full.cpu().numpy()
    scale = np.abs(want).max()
    assert np.abs(got - want).max() <= 1e-4 * scale
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5


@pytest.mark.parametrize("imDim,kDim,world", [((64, 64, 64), (7, 7, 7), 2), ((128, 96, 64), (9, 5, 7), 4),
                                              ((70, 60, 48), (5, 5, 9), 2), ((256, 256, 64), (15, 15, 15), 8),
                                              ((64, 2048, 64), (5, 7, 5), 2),
                                              # ragged slabs: extents not divisible by the number of ranks
                                              ((64, 70, 45), (5, 5, 5), 2), ((64, 90, 50), (5, 5, 9), 4),
                                              ((128, 100, 75), (7, 5, 5), 8), ((64, 64, 75), (5, 5, 5), 8),
                                              ((72, 135, 90), (5, 5, 5), 4)])
@pytest.mark.parametrize("raw", [False, True])
def test_peer_store_exchange_matches_alltoall(fc, dev, imDim, kDim, world, raw):
    """fused compute+exchange (kernels store into the peers' buffers through a pointer table) == the
    all-to-all schedule: bit for bit where both run the same kernel family (power-of-two extents: only the
    destination addresses differ), to fp32 round-off where the z pass falls back to another kernel variant"""
    import torch
    from fourierconvolutioncudalib_b200 import slab
    rng = np.random.default_rng(12)
    n = int(np.prod(imDim))
    im = (rng.random(n, dtype=np.float32) * 1000).astype(np.float32)
    k = gaussian_psf(kDim).reshape(-1)
    d_k = torch.from_numpy(k).to(f"cuda:{dev}")

    ex = slab.LocalExchange(world)
    convs = [slab.SlabConvolver(imDim, kDim, r, world, dev, ex) for r in range(world)]
    full_a = torch.from_numpy(im).to(f"cuda:{dev}")
    for c in convs:
        c.prepare_psf(d_k)
    slab.run_lockstep(convs, [c.slab_of(full_a) for c in convs], ex)

    peers = [slab.PeerSlabConvolver(imDim, kDim, r, world, dev, raw=raw) for r in range(world)]
    full_b = torch.from_numpy(im).to(f"cuda:{dev}")
    for c in peers:
        c.connect_local(peers)
        c.prepare_psf(d_k)
    for _ in range(2):        # twice: buffers are reused across calls
        full_b.copy_(torch.from_numpy(im))
        slab.run_lockstep_peer(peers, [c.slab_of(full_b) for c in peers])
    torch.cuda.synchronize()
    for c in peers:
        c.close()
    pow2 = all(v >= 64 and (v & (v - 1)) == 0 for v in imDim[1:])
    if pow2:
        assert torch.equal(full_a, full_b)
    scale = float(full_a.abs().max())
    assert float((full_a - full_b).abs().max()) <= 2e-6 * scale
    want = im.copy()
    fc.convolution3DfftCUDAInPlace(want, imDim, k, kDim, dev)
    got = full_b.cpu().numpy()
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5
