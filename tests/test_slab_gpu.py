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
                                              ((70, 60, 48), (5, 5, 9), 2), ((256, 256, 64), (15, 15, 15), 8)])
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
