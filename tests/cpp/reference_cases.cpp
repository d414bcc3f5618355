// Boost-free C++ port of a subset of the reference's own GPU test cases, linked against the drop-in library exactly
// like the reference's test executables are linked against theirs (/root/reference/tests/CMakeLists.txt:23-27):
//   * tests/test_gpu_convolve.cpp:202-466  "asymmetric volumes": x=13, y=17, z=19 stacks, 3x3x3 kernels, zero_padd
//     (tests/padd_utils.h:99-171), extents passed reversed ({x,y,z}), result read back through the sub-view
//     (tests/test_fixtures.hpp:254-268), compared with l2norm (tests/test_utils.hpp:75-88: sqrt(sum d^2) / N)
//   * tests/test_gpu_convolve.cpp:15-199   8x8x8 ramp fixture with the identity kernel
// Exit code = number of failed cases.  Built with g++ and run by tests/test_reference_cases_gpu.py (-m gpu).
#include <cmath>
#include <cstdio>
#include <stdexcept>
#include <vector>

#include "convolution3Dfft.h"

namespace {

struct Stack {
    int nz, ny, nx;
    std::vector<float> v;
    Stack(int z, int y, int x, float fill = 0.f) : nz(z), ny(y), nx(x), v((size_t)z * y * x, fill) {}
    float& at(int z, int y, int x) { return v[((size_t)z * ny + y) * nx + x]; }
    float at(int z, int y, int x) const { return v[((size_t)z * ny + y) * nx + x]; }
};

// zero_padd::insert_at_offsets: extent = image + 2*(kernel/2), offset = kernel/2 (tests/padd_utils.h:12-38,157-171)
Stack zero_padd(const Stack& s, int kz, int ky, int kx)
{
    Stack p(s.nz + 2 * (kz / 2), s.ny + 2 * (ky / 2), s.nx + 2 * (kx / 2));
    for (int z = 0; z < s.nz; ++z)
        for (int y = 0; y < s.ny; ++y)
            for (int x = 0; x < s.nx; ++x) p.at(z + kz / 2, y + ky / 2, x + kx / 2) = s.at(z, y, x);
    return p;
}

// l2norm of tests/test_utils.hpp:75-88 between the expectation and the sub-view of the padded result
double l2norm_subview(const Stack& expected, const Stack& padded, int oz, int oy, int ox)
{
    double sum = 0.0;
    for (int z = 0; z < expected.nz; ++z)
        for (int y = 0; y < expected.ny; ++y)
            for (int x = 0; x < expected.nx; ++x) {
                const double d = (double)padded.at(z + oz, y + oy, x + ox) - (double)expected.at(z, y, x);
                sum += d * d;
            }
    return std::sqrt(sum) / (double)expected.v.size();
}

int run_case(const char* name, const Stack& stack, const Stack& kernel, const Stack& expected, double threshold, int dev)
{
    Stack padded = zero_padd(stack, kernel.nz, kernel.ny, kernel.nx);
    int imDim[3] = {padded.nx, padded.ny, padded.nz};          // reversed extents, as the reference tests pass them
    int kDim[3] = {kernel.nx, kernel.ny, kernel.nz};
    std::vector<float> k = kernel.v;
    try {
        convolution3DfftCUDAInPlace(padded.v.data(), imDim, k.data(), kDim, dev);
    } catch (const std::runtime_error& e) {                    // tests/test_gpu_convolve.cpp:237-247
        std::printf("%-40s EXCEPTION %s\n", name, e.what());
        return 1;
    }
    const double l2 = l2norm_subview(expected, padded, kernel.nz / 2, kernel.ny / 2, kernel.nx / 2);
    const bool ok = l2 < threshold;
    std::printf("%-40s l2norm %.3e  limit %.1e  %s\n", name, l2, threshold, ok ? "ok" : "FAILED");
    return ok ? 0 : 1;
}

}  // namespace

int main()
{
    const int dev = selectDeviceWithHighestComputeCapability();
    if (dev < 0 || getNumDevicesCUDA() < 1) {
        std::printf("no CUDA device\n");
        return 100;
    }
    char name[256];
    getNameDeviceCUDA(dev, name);
    std::printf("device %d: %s, %lld MB\n", dev, name, getMemDeviceCUDA(dev) >> 20);
    int fails = 0;

    // ---- asymmetric volumes: z=19, y=17, x=13 (test_gpu_convolve.cpp:202-466)
    const int Z = 19, Y = 17, X = 13;
    {
        Stack ramp(Z, Y, X), k(3, 3, 3);
        for (size_t i = 0; i < ramp.v.size(); ++i) ramp.v[i] = (float)i;
        k.at(1, 1, 1) = 1.f;
        fails += run_case("identity_convolve_of_prime_shape", ramp, k, ramp, 1e-4, dev);
    }
    const Stack ones(Z, Y, X, 1.f), threes(Z, Y, X, 3.f);
    {
        Stack k(3, 3, 3);
        for (int i = 0; i < 3; ++i) k.at(1, 1, i) = 1.f;
        fails += run_case("horizontal_convolve_of_prime_shape", ones, k, threes, 1e-2, dev);
    }
    {
        Stack k(3, 3, 3);
        for (int i = 0; i < 3; ++i) k.at(1, i, 1) = 1.f;
        fails += run_case("vertical_convolve_of_prime_shape", ones, k, threes, 2e-2, dev);
    }
    {
        Stack k(3, 3, 3);
        for (int i = 0; i < 3; ++i) k.at(i, i, i) = 1.f;
        fails += run_case("diagonal_convolve_of_prime_shape", ones, k, threes, 2e-2, dev);
    }
    // ---- 8x8x8 ramp fixture, identity kernel (test_fixtures.hpp:145,197-204; test_gpu_convolve.cpp:15-60)
    {
        Stack ramp(8, 8, 8), k(3, 3, 3);
        for (size_t i = 0; i < ramp.v.size(); ++i) ramp.v[i] = (float)i;
        k.v[27 / 2] = 1.f;
        fails += run_case("identity_convolve_8x8x8_fixture", ramp, k, ramp, 1e-5, dev);
    }
    std::printf("reference_cases: %d failure(s)\n", fails);
    return fails;
}
