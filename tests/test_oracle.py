"""Pins the CPU oracle (oracle/) against the reference test-suite's own expectations (no GPU)."""
import numpy as np
import pytest

import refcases
from oracle import c_oracle as co
from oracle import fc_oracle as fo

# golden interior sums of the 8^3 ramp fixture (direct convolution, tests/test_fixtures.hpp:254-261)
GOLDEN_SUMS = {"identity": 130816.0, "horizontal": 719040.0, "vertical": 715904.0, "depth": 690816.0,
               "all1": 2720564.0}


def oracle_convolve(im, imDim, kernel, kernelDim):
    return fo.convolve_inplace_ref(im, imDim, kernel, kernelDim)


def test_direct_convolution_golden_sums():
    cases, img, kernels = refcases.legacy_convolution_cases()
    for name, gold in GOLDEN_SUMS.items():
        k = kernels[name]
        padded, off = fo.zero_padd(img, k.shape)
        res, _ = co.direct_convolve(padded, k, off)
        s = float(np.sum(fo.crop(res, off, img.shape).astype(np.float32), dtype=np.float32))
        assert s == gold, (name, s)


def test_c_and_numpy_direct_convolution_agree_bitwise():
    rng = np.random.default_rng(7)
    img = rng.random((6, 7, 9), dtype=np.float32)
    k = rng.random((3, 5, 3), dtype=np.float32)
    padded, off = fo.zero_padd(img, k.shape)
    a, _ = co.direct_convolve(padded, k, off)
    b = fo.direct_convolve(padded, k, off)
    assert np.array_equal(a, b)
    c, n = co.direct_convolve(padded, k, off, threads="all")
    assert np.array_equal(a, c) and n >= 1


def test_fft_model_equals_direct_convolution_on_cubic_volumes():
    """cubic volume => the FFT path is the true convolution."""
    cases, img, kernels = refcases.legacy_convolution_cases()
    for c in cases:
        padded, off = c["padded"], c["off"]
        direct, _ = co.direct_convolve(padded, c["kernel"], off)
        fft = oracle_convolve(padded.reshape(-1), c["imDim"], c["kernel"].reshape(-1), c["kernelDim"])
        got = fo.crop(fft.reshape(padded.shape), off, img.shape)
        exp = fo.crop(direct, off, img.shape)
        assert np.abs(got - exp).max() < 1e-8 * np.abs(exp).max()
        # the reference's own check: float sums within 1e-5 percent (test_gpu_convolve.cpp:88)
        s, se = np.float32(got.astype(np.float32).sum(dtype=np.float32)), exp.sum(dtype=np.float32)
        assert abs(s - se) <= 1e-7 * abs(se)


def test_trivial_kernel_gives_exact_zero():
    img = np.arange(512, dtype=np.float32)
    out = oracle_convolve(img, [8, 8, 8], np.zeros(27, np.float32), [3, 3, 3])
    assert float(np.sum(out)) == 0.0                      # test_gpu_convolve.cpp:18-30


@pytest.mark.parametrize("case", refcases.asymmetric_cases(), ids=lambda c: c[0])
def test_asymmetric_volume_cases_pass_with_reference_placement(case):
    """The reference's thresholds are met WITH its PSF placement quirk."""
    name, stack, kernel, expected, thr = case
    got = refcases.run_case(oracle_convolve, stack, kernel)
    assert fo.l2norm(expected, got.astype(np.float32)) < thr, name


@pytest.mark.parametrize("case", refcases.stability_cases(max_edge=128), ids=lambda c: c[0])
def test_numerical_stability_cases(case):
    name, stack, kernel, factor, expected, thr = case
    got = refcases.run_case(oracle_convolve, stack, kernel, factor)
    assert fo.l2norm(expected, got.astype(np.float32)) < thr, name


def test_place_psf_matches_reference_formula_bruteforce():
    """independent scalar re-implementation of src/convolution3Dfft.cu:139-165"""
    kd, d = (3, 4, 5), (7, 6, 9)
    k = np.arange(1, 61, dtype=np.float32)
    S = np.zeros(np.prod(d))
    for tid in range(60):
        z = tid % kd[2]
        aux = (tid - z) // kd[2]
        y = aux % kd[1]
        x = (aux - y) // kd[1]
        x -= kd[0] // 2; y -= kd[1] // 2; z -= kd[2] // 2
        if x < 0: x += d[0]
        if y < 0: y += d[1]
        if z < 0: z += d[2]
        S[z + d[2] * (y + d[1] * x)] = k[tid]
    assert np.array_equal(S, fo.place_psf(k, kd, d))


def test_center_tap_lands_on_origin():
    for kd, d in (((3, 3, 5), (10, 12, 14)), ((31, 31, 91), (46, 46, 106))):
        k = np.zeros(kd, np.float32)
        k[kd[0] // 2, kd[1] // 2, kd[2] // 2] = 1
        S = fo.place_psf(k.reshape(-1), kd, d)
        assert S[0] == 1 and S.sum() == 1


# ---- padded entry point's oracle (fo.convolve_padded_ref): what a reference caller does around the ABI -------
def test_padded_extents_follow_zero_padd_and_the_smooth_policy(fc):
    # caller-padded flavours of the BASELINE configs
    want = {((64, 64, 64), (3, 3, 3)): (70, 70, 70), ((256, 256, 256), (15, 15, 15)): (270, 270, 270),
            ((512, 512, 256), (31, 31, 41)): (560, 560, 300), ((384, 384, 384), (25, 25, 61)): (420, 420, 448),
            ((2048, 2048, 1024), (63, 63, 101)): (2160, 2160, 1125)}
    for (im, k), smooth in want.items():
        assert fo.padded_extents(im, k, 0) == fo.zero_padd_extents(im, k)
        assert fo.padded_extents(im, k, 1) == smooth
        for policy in (0, 1):                      # the library's planner agrees (pure host code, no GPU needed)
            assert fc.padded_extents(im, k, policy) == fo.padded_extents(im, k, policy)
    assert fo.padded_extents((17, 5, 9), (4, 3, 5), 1) == (24, 7, 14)     # fastest extent kept even


def test_padded_oracle_zero_mode_is_the_reference_test_sequence():
    """zero_padd + convolve + sub-view (tests/test_fixtures.hpp:254-268) == CPU convolve on the padded stack"""
    rng = np.random.default_rng(2)
    im3 = rng.random((10, 10, 10), dtype=np.float32)
    k3 = rng.random((3, 3, 3), dtype=np.float32)
    padded, off = fo.zero_padd(im3, k3.shape)
    want = fo.crop(fo.direct_convolve(padded, k3, off), off, im3.shape).reshape(-1)
    for policy in (0, 1):
        got = fo.convolve_padded_ref(im3.reshape(-1), (10, 10, 10), k3.reshape(-1), (3, 3, 3), 0, policy)
        assert np.abs(got - want).max() <= 1e-5 * np.abs(want).max()


def test_padded_oracle_mirror_mode_keeps_a_constant_image_constant():
    k = np.random.default_rng(4).random(5 * 3 * 7)
    k /= k.sum()
    im = np.full(12 * 9 * 8, 42.0)
    got = fo.convolve_padded_ref(im, (12, 9, 8), k, (5, 3, 7), 1, 1)
    assert np.allclose(got, 42.0, rtol=1e-12)
    # zero mode darkens the border instead
    assert fo.convolve_padded_ref(im, (12, 9, 8), k, (5, 3, 7), 0, 1).min() < 41.0


def test_legacy_subsampled_multiply_is_not_a_convolution():
    """the only surviving piece of the legacy "SaveMemory" idea
    (modulateAndNormalizeSubsampled_kernel, src/convolution3Dfft.cu:65-125; no host caller in the snapshot, so there is
    nothing to pin it against) is 2-8 % away from the convolution the entry point's header promises
    (src/convolution3Dfft.h:58-64: "like InPlace").  That is why convolution3DfftCUDAInPlaceSaveMemory here returns the
    InPlace result instead (tests/test_parity_gpu.py::test_savememory_entry_point holds it to 1e-4 / 1e-5)."""
    rng = np.random.default_rng(0)
    for d, k in (((32, 32, 32), (7, 7, 7)), ((48, 48, 48), (15, 15, 15))):
        im = rng.random(int(np.prod(d)))
        ax = [np.exp(-0.5 * ((np.arange(n) - n // 2) / (n / 6.0)) ** 2) for n in k]
        psf = ax[0][:, None, None] * ax[1][None, :, None] * ax[2][None, None, :]
        psf /= psf.sum()
        got = fo.modulate_subsampled_ref(im, d, psf.reshape(-1), k)
        # cubic volume: the reference InPlace placement is the plain centred PSF, in either axis convention
        want = fo.convolve_inplace_ref(im, d, psf.reshape(-1), k)
        err = np.linalg.norm(got - want) / np.linalg.norm(want)
        assert 1e-2 < err < 1e-1, err            # three orders of magnitude above the parity tolerance (1e-5)
