// Shared declarations of the B200-native FFT convolution library (host + device).
#pragma once
#include <cuda_runtime.h>
#include <cstddef>
#include <cstdint>
#include <sstream>
#include <stdexcept>
#include <string>

namespace fcb200 {

constexpr int kMaxStages = 16;
constexpr int kMaxDynSmem = 227 * 1024;   // usable shared memory per CTA on sm_100a
constexpr int kColThreads = 256;          // CTA size of every FFT kernel

// Device view of a 1D transform plan (passed to kernels by value).
struct AxisPlanDev {
    int L;                    // transform length
    int ns;                   // number of stages
    int radix[kMaxStages];    // stage radices, product == L
    int generic;              // 1 when a prime factor above 23 is present (direct-sum stage, needs two tile buffers)
    int big;                  // 1 when a radix 11, 13, 17, 19 or 23 is present (kernels compiled with those butterflies)
    const float2* tw;         // L roots: exp(-2*pi*i*t/L), computed in double
    const int* rev;           // rev[p]  = frequency held at position p after the forward transform
    const int* pos;           // pos[k]  = position that holds frequency k (inverse permutation)
    // Rader stage for the LAST radix when it is a large prime p (fft_engine.cuh: stage_rader): the length-p DFT of a
    // contiguous block becomes a cyclic convolution of length n = p - 1, done with an n-point FFT of smooth radices.
    int rader_p;              // 0: none
    int rader_n;              // p - 1
    int rader_ns;             // stages of the n-point transform (all register radices)
    int rader_radix[8];
    const float2* rader_tw;   // n roots exp(-2*pi*i*t/n)
    const int* rader_perm;    // perm[m]  = g^m mod p,    m = 0 .. n-1
    const int* rader_iperm;   // iperm[q] = g^(-q) mod p, q = 0 .. n-1
    const float2* rader_bf;   // spectrum of b[t] = exp(-2*pi*i*g^(-t)/p), / n, in POSITION order of the n-point DIF
    const float2* rader_bi;   // the same for the inverse (conjugate roots)
};

// Geometry of one convolution problem as the kernels see it.
//   real volume  : [nz][ny][nx] floats, nx fastest (reference: imDim = {nx, ny, nz},
//                  /root/reference/src/convolution3Dfft.cu:417-421)
//   spectrum     : [nz][ny][xcp] float2, xc = nx/2+1 valid bins per row, xcp = xc rounded up to 4
//                  (32-byte sector alignment of every row); kx, ky and kz are all in natural order,
//                  i.e. the buffer is numpy.fft.rfftn of the volume with padded rows -- except that
//                  each 16 bytes hold two adjacent kx bins in PAIR-PLANAR form (re0, re1, im0, im1)
//                  so the strided passes can use packed fp32 arithmetic without any shuffles.
struct Geometry {
    int nx, ny, nz;
    int xc, xcp;
    int M;        // length of the complex transform used along x: nx/2 (even nx) or nx (odd nx)
    int odd;      // nx odd
};

// PSF placement parameters (fftShiftKernel as called by the reference,
// /root/reference/src/convolution3Dfft.cu:128-166 with the arguments of :454-461).
struct PsfGather {
    const float* kernel;  // k0*k1*k2 taps (device)
    int k0, k1, k2;
    int d0, d1, d2;
};

// In-library padding: source volume [sz][sy][sx] embedded at offsets (ox,oy,oz) in the padded volume
// [pz][py][px] (x fastest).  mode 0: zeros outside (reference tests/padd_utils.h:157-171); 1: mirror.
struct PadGeom {
    int sx, sy, sz;
    int px, py, pz;
    int ox, oy, oz;
    int mode;
};

inline void throw_cuda(cudaError_t err, const char* what, const char* file, int line)
{
    if (err != cudaSuccess) {
        std::ostringstream msg;
        msg << cudaGetErrorString(err) << " (" << what << ") in " << file << " at line " << line;
        throw std::runtime_error(msg.str());
    }
}
// ---- programmatic dependent launch (PDL, sm_90+) ------------------------------------------------------------
// Consecutive passes on one stream are launched with cudaLaunchAttributeProgrammaticStreamSerialization: every
// kernel calls pdl_launch_dependents() first and pdl_wait() after its prologue (twiddle tables into shared memory:
// plan constants, never produced by the previous pass), so the next pass's CTAs are scheduled and run their prologue
// while the last wave of the previous pass drains; griddepcontrol.wait then blocks until the previous grid has
// completed and its writes are visible.  Both instructions are no-ops in a normally launched kernel.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
#endif
bool pdl_enabled();   // FCB200_PDL (default 1); fc_plan.cu

#define FC_CUDA(expr) ::fcb200::throw_cuda((expr), #expr, __FILE__, __LINE__)
#define FC_CUDA_KERNEL() ::fcb200::throw_cuda(cudaPeekAtLastError(), "kernel launch", __FILE__, __LINE__)

#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
inline void launch_pdl(bool allow, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                       Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = allow ? 1 : 0;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    throw_cuda(cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...), "kernel launch", __FILE__, __LINE__);
}
#endif

}  // namespace fcb200
