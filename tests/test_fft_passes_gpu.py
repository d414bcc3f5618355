"""Pass-level parity of the hand-written FFT kernels against numpy (float64) through the C ABI's
debug entry points.  Tolerance: relative L2 <= 2e-6 (fp32 FFT round-off), written per assert."""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu

SHAPES = [  # imDim = (d0 fastest, d1, d2)
    (8, 8, 8), (16, 4, 2), (64, 64, 64), (10, 10, 10), (12, 6, 10), (70, 70, 70),
    (15, 19, 21),            # reference asymmetric_volumes padded extents (odd x)
    (13, 17, 19), (46, 46, 106), (130, 130, 132), (66, 66, 66),
    (158, 22, 18),           # 2*79: generic radix
    (256, 256, 16), (512, 32, 8), (270, 30, 20), (300, 40, 28), (2, 2, 2), (4, 1, 1), (6, 2, 1),
    (560, 300, 8), (448, 420, 4), (420, 560, 3), (270, 448, 2), (32, 6, 300), (32, 4, 448), (300, 4, 560),
    (1024, 16, 4), (2048, 8, 6), (128, 34, 3), (256, 5, 3), (384, 12, 4), (32, 1024, 2), (16, 4, 1024),
    # composite register radices 6, 9, 10, 12, 15 (fft_butterflies.cuh: DftCT) on every axis
    (12, 6, 9), (30, 18, 12), (20, 10, 15), (90, 45, 36), (150, 135, 10), (24, 96, 75), (2160, 4, 2),
    (4, 1080, 2), (6, 4, 1125), (540, 12, 6), (36, 540, 3), (10, 6, 810),
    # two-stage plans with fat composite radices 18, 20, 21, 28 (fc_plan.cu: factorize) on the axes they are planned for
    (32, 270, 6), (16, 6, 420), (32, 8, 270), (272, 300, 3), (272, 420, 2),
    # lengths without a compile-time plan on the run-time-radix TMA pipeline (two and three stages, with and without planes)
    (32, 400, 6), (16, 6, 360), (32, 288, 4), (16, 4, 350), (16, 480, 3), (48, 160, 5), (20, 200, 1), (272, 640, 2), (16, 2, 600),
]


def rel_l2(a, b):
    return float(np.linalg.norm((a - b).ravel()) / max(np.linalg.norm(b.ravel()), 1e-30))


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_forward_passes_match_numpy(fc, dev, shape):
    d0, d1, d2 = shape
    rng = np.random.default_rng(d0 * 7 + d1 * 3 + d2)
    vol = rng.standard_normal((d2, d1, d0)).astype(np.float32)
    v64 = vol.astype(np.float64)
    want1 = np.fft.rfft(v64, axis=2)
    want2 = np.fft.fft(want1, axis=1)
    want3 = np.fft.fft(want2, axis=0)
    for passes, want in ((1, want1), (2, want2), (3, want3)):
        got = fc.debug_rfft3(vol, shape, passes, dev)
        assert rel_l2(got, want) < 2e-6, (shape, passes, rel_l2(got, want))


@pytest.mark.parametrize("shape", SHAPES, ids=lambda s: "x".join(map(str, s)))
def test_inverse_matches_numpy(fc, dev, shape):
    d0, d1, d2 = shape
    rng = np.random.default_rng(d0 + d1 * 5 + d2 * 11)
    vol = rng.standard_normal((d2, d1, d0))
    spec = np.fft.rfftn(vol, axes=(0, 1, 2))
    got = fc.debug_irfft3(spec.astype(np.complex64), shape, dev)
    want = vol * vol.size                       # unnormalised, like cufftExecC2R
    assert rel_l2(got, want) < 2e-6, (shape, rel_l2(got, want))


@pytest.mark.parametrize("imDim,kDim", [((16, 16, 16), (3, 3, 3)), ((15, 19, 21), (3, 3, 3)), ((46, 46, 106), (31, 31, 91)),
                                        ((20, 12, 10), (5, 4, 3)), ((64, 64, 64), (7, 9, 11)), ((64, 32, 16), (16, 8, 4))])
def test_fused_psf_placement_spectrum(fc, dev, imDim, kDim):
    """zero-pad + circular shift fused into the x pass == FFT of the reference's placed PSF"""
    from oracle import fc_oracle as fo
    rng = np.random.default_rng(5)
    k = rng.random(int(np.prod(kDim))).astype(np.float32)
    S = fo.place_psf(k, kDim, imDim).reshape(imDim[2], imDim[1], imDim[0])
    want = np.fft.rfftn(S, axes=(0, 1, 2))
    got = fc.debug_psf_spectrum(k, kDim, imDim, dev)
    assert rel_l2(got, want) < 2e-6
