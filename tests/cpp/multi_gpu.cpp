// C++ caller of the single-process multi-GPU entry points (include/fcb200_ext.h): fcb200_convolve_slab (one volume in
// z slabs), fcb200_convolve_slab_device and fcb200_convolve_batch_multi must reproduce what the reference-facing
// single-device call convolution3DfftCUDAInPlace (reference src/convolution3Dfft.h:56) returns for the same inputs.
//   usage: multi_gpu [ndev]      ndev ranks; when the box has fewer GPUs the ranks share device 0 (emulated ranks)
// Exit code = number of failed cases.  Built with g++ and run by tests/test_multi_gpu.py (-m gpu).
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <vector>

#include "fcb200_ext.h"

namespace {

std::vector<float> ramp_noise(size_t n, unsigned seed)
{
    std::vector<float> v(n);
    unsigned s = seed * 2654435761u + 12345u;
    for (size_t i = 0; i < n; ++i) {
        s = s * 1664525u + 1013904223u;
        v[i] = (float)(s >> 8) * (1000.0f / 16777216.0f);
    }
    return v;
}

std::vector<float> gaussian(const int* k)
{
    std::vector<float> v((size_t)k[0] * k[1] * k[2]);
    double sum = 0.0;
    for (int a = 0; a < k[0]; ++a)
        for (int b = 0; b < k[1]; ++b)
            for (int c = 0; c < k[2]; ++c) {
                const double da = (a - k[0] / 2) / (k[0] / 6.0), db = (b - k[1] / 2) / (k[1] / 6.0), dc = (c - k[2] / 2) / (k[2] / 6.0);
                const double g = std::exp(-0.5 * (da * da + db * db + dc * dc));
                v[((size_t)a * k[1] + b) * k[2] + c] = (float)g;
                sum += g;
            }
    for (float& x : v) x = (float)(x / sum);
    return v;
}

// max |a-b| / max |b| and relative L2 (north_star tolerances: 1e-4, 1e-5)
bool close(const std::vector<float>& a, const std::vector<float>& b, const char* what)
{
    double mx = 0.0, ref = 0.0, num = 0.0, den = 0.0;
    for (size_t i = 0; i < a.size(); ++i) {
        const double d = (double)a[i] - (double)b[i];
        mx = std::fmax(mx, std::fabs(d));
        ref = std::fmax(ref, std::fabs((double)b[i]));
        num += d * d;
        den += (double)b[i] * (double)b[i];
    }
    const double rel = mx / ref, l2 = std::sqrt(num / den);
    const bool ok = rel <= 1e-4 && l2 <= 1e-5;
    std::printf("%-58s max_rel %.2e  rel_l2 %.2e  %s\n", what, rel, l2, ok ? "ok" : "FAILED");
    return ok;
}

}  // namespace

int main(int argc, char** argv)
{
    int ndev = argc > 1 ? std::atoi(argv[1]) : 2;
    const int have = getNumDevicesCUDA();
    if (have < 1) {
        std::printf("no CUDA device\n");
        return 99;
    }
    std::vector<int> devs((size_t)ndev);
    for (int r = 0; r < ndev; ++r) devs[(size_t)r] = have >= ndev ? r : 0;
    std::printf("ranks %d on %d device(s)%s\n", ndev, have, have >= ndev ? "" : " (emulated ranks on device 0)");
    int failures = 0;

    // ---- one volume in slabs: divisible and ragged extents, host pointer
    const int shapes[3][6] = {{128, 96, 64, 9, 5, 7}, {70, 60, 45, 5, 5, 9}, {256, 128, 50, 7, 7, 7}};
    for (const auto& s : shapes) {
        int imDim[3] = {s[0], s[1], s[2]}, kDim[3] = {s[3], s[4], s[5]};
        const size_t n = (size_t)imDim[0] * imDim[1] * imDim[2];
        std::vector<float> im = ramp_noise(n, (unsigned)s[0]), k = gaussian(kDim);
        std::vector<float> want = im, got = im;
        try {
            convolution3DfftCUDAInPlace(want.data(), imDim, k.data(), kDim, devs[0]);
            for (int rep = 0; rep < 2; ++rep) {   // twice: the context and its buffers are reused
                got = im;
                fcb200_convolve_slab(got.data(), imDim, k.data(), kDim, devs.data(), ndev);
            }
        } catch (const std::runtime_error& e) {
            std::printf("EXCEPTION %s\n", e.what());
            ++failures;
            continue;
        }
        char what[128];
        std::snprintf(what, sizeof what, "slab %dx%dx%d (x) %dx%dx%d over %d ranks", s[0], s[1], s[2], s[3], s[4], s[5], ndev);
        failures += close(got, want, what) ? 0 : 1;
        float ms[64];
        const int nr = fcb200_slab_last_timing(imDim, devs.data(), ndev, ms, 64);
        if (nr != ndev || !(ms[3] > 0.f)) {
            std::printf("fcb200_slab_last_timing: %d ranks, total %.3f ms  FAILED\n", nr, ms[3]);
            ++failures;
        }
    }

    // ---- batch of independent blocks over the devices
    {
        int imDim[3] = {96, 64, 48}, kDim[3] = {7, 5, 9};
        const size_t n = (size_t)imDim[0] * imDim[1] * imDim[2];
        const int nb = 7;
        std::vector<float> k = gaussian(kDim);
        std::vector<std::vector<float>> blocks, want;
        for (int b = 0; b < nb; ++b) blocks.push_back(ramp_noise(n, 100u + (unsigned)b));
        want = blocks;
        std::vector<float*> ptrs;
        for (auto& b : blocks) ptrs.push_back(b.data());
        std::vector<int> taken((size_t)ndev, 0);
        try {
            for (auto& w : want) convolution3DfftCUDAInPlace(w.data(), imDim, k.data(), kDim, devs[0]);
            fcb200_convolve_batch_multi(ptrs.data(), nb, imDim, k.data(), kDim, devs.data(), ndev, taken.data());
            int total = 0;
            for (int t : taken) total += t;
            bool ok = total == nb;
            for (int b = 0; b < nb; ++b) {
                char what[128];
                std::snprintf(what, sizeof what, "batch_multi block %d of %d over %d devices", b, nb, ndev);
                ok = close(blocks[(size_t)b], want[(size_t)b], what) && ok;
            }
            std::printf("blocks per device:");
            for (int t : taken) std::printf(" %d", t);
            std::printf("\n");
            failures += ok ? 0 : 1;
        } catch (const std::runtime_error& e) {
            std::printf("EXCEPTION %s\n", e.what());
            ++failures;
        }
    }

    // ---- errors still cross the boundary as std::runtime_error
    {
        int imDim[3] = {64, 64, 4}, kDim[3] = {3, 3, 3};
        std::vector<float> im(64 * 64 * 4, 1.f), k(27, 1.f);
        std::vector<int> many(8, devs[0]);
        bool threw = false;
        try {
            fcb200_convolve_slab(im.data(), imDim, k.data(), kDim, many.data(), 8);   // 4 planes over 8 ranks
        } catch (const std::runtime_error&) {
            threw = true;
        }
        std::printf("%-58s %s\n", "slab with more ranks than planes throws", threw ? "ok" : "FAILED");
        failures += threw ? 0 : 1;
    }
    fcb200_release();
    std::printf("%d failure(s)\n", failures);
    return failures;
}
