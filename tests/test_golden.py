"""Golden fixtures = outputs of the reference's own cuFFT build (tests/golden/make_golden.py, run on the
B200 box).  CPU: the oracle must reproduce them (this pins the restatement, including the PSF-placement
quirk on non-cubic volumes, without a GPU).  GPU: the product must reproduce them through the C ABI.
Tolerance: max|err| <= 1e-4*max|out|, relative L2 <= 1e-5 (north_star), fp32."""
import glob
import os

import numpy as np
import pytest

from oracle import fc_oracle as fo

HERE = os.path.dirname(os.path.abspath(__file__))
FIXTURES = sorted(glob.glob(os.path.join(HERE, "golden", "*.npz")))


def _check(got, want):
    got = np.asarray(got, np.float64).ravel()
    want = np.asarray(want, np.float64).ravel()
    assert np.abs(got - want).max() <= 1e-4 * np.abs(want).max()
    assert np.linalg.norm(got - want) / np.linalg.norm(want) <= 1e-5


def test_fixtures_exist():
    assert len(FIXTURES) >= 6


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_oracle_reproduces_reference_output(path):
    f = np.load(path)
    _check(fo.convolve_inplace_ref(f["im"], f["imDim"], f["kernel"], f["kernelDim"]), f["out"])


@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_noncubic_fixtures_really_exercise_the_quirk(path):
    """a 'true' centred convolution (what the placement would be without the axis mix-up) must NOT
    match on the non-cubic fixtures -- otherwise they would not pin the quirk"""
    f = np.load(path)
    d0, d1, d2 = (int(v) for v in f["imDim"])
    if d0 == d1 == d2:
        pytest.skip("cubic volume: placement is a plain centred PSF")
    k0, k1, k2 = (int(v) for v in f["kernelDim"])
    K = f["kernel"].reshape(k0, k1, k2).astype(np.float64)
    S = np.zeros((d2, d1, d0))
    for a in range(k0):           # the consistent reading: kernelDim[i] pairs with imDim[i]
        for b in range(k1):
            for c in range(k2):
                S[(c - k2 // 2) % d2, (b - k1 // 2) % d1, (a - k0 // 2) % d0] = K[a, b, c]
    I3 = f["im"].astype(np.float64).reshape(d2, d1, d0)
    true_conv = np.fft.irfftn(np.fft.rfftn(I3) * np.fft.rfftn(S), s=I3.shape, axes=(0, 1, 2)).ravel()
    rel = np.linalg.norm(true_conv - f["out"]) / np.linalg.norm(f["out"])
    assert rel > 1e-3


@pytest.mark.gpu
@pytest.mark.parametrize("path", FIXTURES, ids=lambda p: os.path.basename(p)[:-4])
def test_product_reproduces_reference_output(fc, dev, path):
    f = np.load(path)
    got = f["im"].copy()
    fc.convolution3DfftCUDAInPlace(got, f["imDim"], f["kernel"].copy(), f["kernelDim"], dev)
    _check(got, f["out"])
