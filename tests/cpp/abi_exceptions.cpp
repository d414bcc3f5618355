// C++ caller of the drop-in library, the way the reference's own test executables call it
// (/root/reference/tests/CMakeLists.txt:23-27 link the shared library; /root/reference/tests/test_gpu_convolve.cpp:237-247
// catch std::runtime_error thrown across the C boundary, /root/reference/src/book.h:112-123).  Only entry points that
// need no GPU are exercised, so this runs in the CPU test suite (tests/test_abi.py builds and runs it).
#include <cstdio>
#include <cstring>
#include <stdexcept>

#include "convolution3Dfft.h"
#include "fcb200_ext.h"

static int fails = 0;
#define EXPECT(cond)                                                        \
    do {                                                                    \
        if (!(cond)) {                                                      \
            std::printf("FAILED line %d: %s\n", __LINE__, #cond);           \
            ++fails;                                                        \
        }                                                                   \
    } while (0)

int main()
{
    // 1. a recoverable failure surfaces as std::runtime_error in the caller, and the message is kept
    int shape4[4] = {1, 2, 3, 4};
    bool caught = false;
    try {
        gpu_mem_needed_mb(shape4, 4);
    } catch (const std::runtime_error& e) {
        caught = std::strstr(e.what(), "len") != nullptr;
    }
    EXPECT(caught);
    EXPECT(std::strlen(fcb200_last_error()) > 0);

    // 2. record-only mode (JNA / ctypes callers cannot catch C++ exceptions): no throw, error string set, then cleared
    fcb200_set_error_mode(1);
    int v = -1;
    try {
        v = gpu_mem_needed_mb(shape4, 4);
    } catch (...) {
        v = -2;
    }
    EXPECT(v == 0);
    EXPECT(std::strlen(fcb200_last_error()) > 0);
    int shape3[3] = {256, 512, 512};
    EXPECT(gpu_mem_needed_mb(shape3, 3) > 0);
    EXPECT(std::strlen(fcb200_last_error()) == 0);
    fcb200_set_error_mode(0);

    // 3. host-only entry points
    EXPECT(cuda_version() >= 12000);
    int im[3] = {512, 512, 256}, k[3] = {31, 31, 41}, pad[3] = {0, 0, 0};
    fcb200_padded_extents(im, k, 0, pad);      // the reference's zero_padd: image + 2*(kernel/2)
    EXPECT(pad[0] == 542 && pad[1] == 542 && pad[2] == 296);
    fcb200_padded_extents(im, k, 1, pad);      // 7-smooth
    EXPECT(pad[0] == 560 && pad[1] == 560 && pad[2] == 300);
    int radix[16], generic = -1;
    EXPECT(fcb200_plan_radices(270, radix, &generic) == 2 && radix[0] == 18 && radix[1] == 15 && generic == 0);
    EXPECT(fcb200_plan_radices(79, radix, &generic) == 1 && generic == 1);
    EXPECT(fcb200_spectrum_pitch(512) == 260);
    bool bad = false;
    try {
        fcb200_padded_extents(im, k, 7, pad);
    } catch (const std::runtime_error&) {
        bad = true;
    }
    EXPECT(bad);

    std::printf("abi_exceptions: %d failure(s)\n", fails);
    return fails;
}
