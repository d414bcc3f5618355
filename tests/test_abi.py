"""The C-ABI library loads and exports every symbol the headers declare (no compute, no GPU)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    names = []
    for h in ("convolution3Dfft.h", "fcb200_ext.h"):
        txt = open(os.path.join(ROOT, "include", h)).read()
        txt = re.sub(r"/\*.*?\*/", "", txt, flags=re.S)
        for m in re.finditer(r"FUNCTION_PREFIX\s+[^;(]*?(\w+)\s*\(", txt):
            names.append(m.group(1))
    return names


def test_headers_declare_the_reference_abi():
    names = declared_symbols()
    # exported symbols of a reference build + the SaveMemory entry north_star names
    for ref in ("convolution3DfftCUDA", "convolution3DfftCUDAInPlace", "convolution3DfftCUDA_test", "cuda_version",
                "getCUDAcomputeCapabilityMajorVersion", "getCUDAcomputeCapabilityMinorVersion", "getMemDeviceCUDA",
                "getNameDeviceCUDA", "getNumDevicesCUDA", "gpu_mem_needed_mb",
                "selectDeviceWithHighestComputeCapability", "convolution3DfftCUDAInPlaceSaveMemory"):
        assert ref in names


def test_library_exports_every_declared_symbol(fc):
    lib = ctypes.CDLL(fc._lib.LIB_PATH)
    for name in declared_symbols():
        assert hasattr(lib, name), name
    assert set(declared_symbols()) == set(fc._lib.ABI_SYMBOLS) | set(fc._lib.EXT_SYMBOLS)


def test_only_c_abi_is_exported(fc):
    import subprocess
    out = subprocess.check_output(["nm", "-D", "--defined-only", fc._lib.LIB_PATH], text=True)
    syms = [l.split()[-1] for l in out.splitlines() if " T " in l]
    assert sorted(syms) == sorted(declared_symbols())


def test_host_only_entry_points(fc):
    assert fc.cuda_version() >= 12000
    assert fc.spectrum_pitch(512) == 260 and fc.spectrum_pitch(13) == 8
    # spectrum + PSF spectrum for 512x512x256: 2 * 256*512*260*8 bytes
    assert fc.workspace_bytes([512, 512, 256]) >= 2 * 256 * 512 * 260 * 8
    assert fc.gpu_mem_needed_mb([256, 512, 512]) == fc.workspace_bytes([512, 512, 256]) // (1 << 20)


def test_errors_surface_as_exceptions(fc):
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.gpu_mem_needed_mb([1, 2, 3, 4])
    with pytest.raises(fc.api.FourierConvolutionError):
        fc.plan_radices(0)


def test_product_does_not_import_the_oracle():
    pkg = os.path.join(ROOT, "fourierconvolutioncudalib_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".cpp")):
                txt = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in txt.replace("no CPU fallback", ""), f


def test_cpp_caller_catches_runtime_error_across_the_c_boundary(fc, tmp_path):
    """the reference's C++ callers rely on std::runtime_error crossing the C ABI (tests/test_gpu_convolve.cpp:237-247);
    tests/cpp/abi_exceptions.cpp is such a caller, restricted to the entry points that need no GPU"""
    import shutil
    import subprocess
    if shutil.which("g++") is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "abi_exceptions")
    libdir = os.path.dirname(fc._lib.LIB_PATH)
    subprocess.check_call(["g++", "-std=c++17", "-O1", "-I", os.path.join(ROOT, "include"),
                           os.path.join(ROOT, "tests", "cpp", "abi_exceptions.cpp"), "-o", exe,
                           "-L", libdir, "-lFourierConvolutionCUDALib", "-Wl,-rpath," + libdir])
    res = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    assert res.returncode == 0, res.stdout + res.stderr
    assert "0 failure(s)" in res.stdout
