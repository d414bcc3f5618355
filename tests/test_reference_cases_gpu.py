"""The reference's own GPU test cases (inputs, call sequence, thresholds) run against the product
through the C ABI -- ported from tests/test_gpu_convolve.cpp and tests/test_gpu_numerical_stability.cpp."""
import numpy as np
import pytest

import refcases
from oracle import c_oracle as co
from oracle import fc_oracle as fo

pytestmark = pytest.mark.gpu


def make_convolve(fc, dev):
    def convolve(im, imDim, kernel, kernelDim):
        fc.convolution3DfftCUDAInPlace(im, imDim, kernel, kernelDim, dev)
        return im
    return convolve


def test_trivial_convolve(fc, dev):
    img = np.arange(512, dtype=np.float32)
    fc.convolution3DfftCUDAInPlace(img, [8, 8, 8], np.zeros(27, np.float32), [3, 3, 3], dev)
    assert float(np.sum(img, dtype=np.float32)) == 0.0          # BOOST_CHECK_CLOSE(sum, 0.f, .00001)


@pytest.mark.parametrize("name", ["horizontal", "vertical", "depth", "all1"])
def test_legacy_convolution_sums(fc, dev, name):
    cases, img, kernels = refcases.legacy_convolution_cases()
    c = [c for c in cases if c["name"] == f"legacy_{name}"][0]
    direct, _ = co.direct_convolve(c["padded"], c["kernel"], c["off"])
    expected = np.float32(fo.crop(direct, c["off"], img.shape).sum(dtype=np.float32))
    got = refcases.run_case(make_convolve(fc, dev), img, c["kernel"], reverse_dims=False)
    s = np.float32(0)
    for v in got.astype(np.float32).reshape(-1):                # std::accumulate in float
        s = np.float32(s + v)
    assert abs(float(s) - float(expected)) <= 1e-7 * abs(float(expected)), (s, expected)   # 1e-5 percent


@pytest.mark.parametrize("case", refcases.asymmetric_cases(), ids=lambda c: c[0])
def test_asymmetric_volumes(fc, dev, case):
    name, stack, kernel, expected, thr = case
    got = refcases.run_case(make_convolve(fc, dev), stack, kernel)
    assert fo.l2norm(expected, got) < thr


@pytest.mark.parametrize("case", refcases.stability_cases(max_edge=256), ids=lambda c: c[0])
def test_numerical_stability(fc, dev, case):
    name, stack, kernel, factor, expected, thr = case
    got = refcases.run_case(make_convolve(fc, dev), stack, kernel, factor)
    assert fo.l2norm(expected, got) < thr


def test_cpp_port_of_reference_cases_links_and_passes(fc, dev, tmp_path):
    """tests/cpp/reference_cases.cpp: the reference's asymmetric-volume and 8^3 identity cases as a C++ program linked
    against the drop-in library like the reference's own test executables (tests/CMakeLists.txt:23-27)"""
    import os
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "reference_cases")
    libdir = os.path.dirname(fc._lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(root, "include"),
                           os.path.join(root, "tests", "cpp", "reference_cases.cpp"), "-o", exe,
                           "-L", libdir, "-lFourierConvolutionCUDALib", "-Wl,-rpath," + libdir])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=300)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failure(s)" in res.stdout
