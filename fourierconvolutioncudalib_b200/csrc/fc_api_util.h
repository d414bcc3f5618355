// Helpers shared by the files that implement extern "C" entry points (fc_api.cu, fc_multi.cu).
#pragma once
#include <atomic>
#include <cstdlib>
#include <stdexcept>
#include <string>

#include "fc_common.h"

namespace fcb200 {

extern thread_local std::string g_last_error;
extern std::atomic<int> g_error_mode;   // 0: throw std::runtime_error (reference convention); 1: record only

// Runs `fn`, records the message of any exception for fcb200_last_error() and rethrows it as
// std::runtime_error -- the reference's convention for recoverable failures
// (/root/reference/src/book.h:112-123; its tests catch it, tests/test_gpu_convolve.cpp:237-247).
template <typename F>
auto guarded(F&& fn) -> decltype(fn())
{
    try {
        g_last_error.clear();
        return fn();
    } catch (const std::exception& e) {
        g_last_error = e.what();
        cudaGetLastError();  // clear non-sticky errors so later calls can proceed
        if (g_error_mode.load() == 1) return decltype(fn())();
        throw std::runtime_error(g_last_error);
    }
}

inline void check_dims(const int* imDim, const int* kernelDim)
{
    if (!imDim) throw std::runtime_error("fcb200: imDim is NULL");
    for (int i = 0; i < 3; ++i)
        if (imDim[i] <= 0) throw std::runtime_error("fcb200: image extents must be positive");
    if (kernelDim)
        for (int i = 0; i < 3; ++i)
            if (kernelDim[i] <= 0) throw std::runtime_error("fcb200: kernel extents must be positive");
}

inline bool env_flag(const char* name, bool dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) != 0 : dflt;
}

struct DeviceGuard {
    int dev;
    explicit DeviceGuard(int d) : dev(d) { FC_CUDA(cudaSetDevice(d)); }  // like the reference, devCUDA stays current
};

// The x kernels move float2 / 8-byte cp.async on `real + row * nx`; the column kernels move 16 bytes on spectrum
// buffers.  A misaligned caller pointer would be a sticky misaligned-address fault that kills the context, so it
// is rejected here with a std::runtime_error instead.
inline void check_real_alignment(const void* p, int nx)
{
    (void)nx;
    if ((reinterpret_cast<uintptr_t>(p) & 7) != 0)
        throw std::runtime_error("fcb200: device image pointer must be 8-byte aligned");
}
inline void check_spec_alignment(const void* p)
{
    if ((reinterpret_cast<uintptr_t>(p) & 15) != 0)
        throw std::runtime_error("fcb200: device spectrum buffer must be 16-byte aligned");
}

}  // namespace fcb200
