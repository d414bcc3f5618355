// In-library padding for the caller-side step the reference leaves to its users
// (/root/reference/src/convolution3Dfft.h:39, :54 "TODO: pad data"; its tests do it on the host with
// zero_padd::insert_at_offsets, /root/reference/tests/padd_utils.h:99-171, and read the result back through the
// same sub-view, /root/reference/tests/test_fixtures.hpp:254-268).
//
// Two small HBM-bound kernels around the FFT passes:
//   pad_embed_kernel : padded planes [pz0, pz0+pn) <- source volume, zero or mirror extension
//   pad_crop_kernel  : source-sized planes [z0, z0+n) <- interior of the padded volume
// Both move rows with one thread per voxel along x (coalesced 4-byte accesses, 4 rows per CTA).
#include "fc_plan.h"

namespace fcb200 {

namespace {

constexpr int kPadThreadsX = 128;
constexpr int kPadRows = 4;

// numpy.pad(..., mode="reflect") index folding: ... 2 1 | 0 1 2 ... n-1 | n-2 n-3 ...
__device__ __forceinline__ int fold_reflect(int s, int n)
{
    if (n == 1) return 0;
    const int T = 2 * (n - 1);
    s %= T;
    if (s < 0) s += T;
    return s < n ? s : T - s;
}

__global__ void __launch_bounds__(kPadThreadsX* kPadRows)
pad_embed_kernel(const float* __restrict__ src, float* __restrict__ dst, PadGeom g, int pz0, long long nrows)
{
    const long long row = (long long)blockIdx.x * kPadRows + threadIdx.y;
    if (row >= nrows) return;
    const int z = pz0 + (int)(row / g.py), y = (int)(row % g.py);
    int sz = z - g.oz, sy = y - g.oy;
    bool zero_row = false;
    if (g.mode == 1) {
        sz = fold_reflect(sz, g.sz);
        sy = fold_reflect(sy, g.sy);
    } else {
        zero_row = sz < 0 || sz >= g.sz || sy < 0 || sy >= g.sy;
    }
    float* out = dst + ((size_t)z * g.py + y) * g.px;
    const float* in = src + ((size_t)(zero_row ? 0 : sz) * g.sy + (zero_row ? 0 : sy)) * g.sx;
    for (int x = threadIdx.x; x < g.px; x += kPadThreadsX) {
        int sxp = x - g.ox;
        float v = 0.f;
        if (g.mode == 1) v = in[fold_reflect(sxp, g.sx)];
        else if (!zero_row && sxp >= 0 && sxp < g.sx) v = in[sxp];
        out[x] = v;
    }
}

__global__ void __launch_bounds__(kPadThreadsX* kPadRows)
pad_crop_kernel(const float* __restrict__ pad, float* __restrict__ dst, PadGeom g, int z0, long long nrows)
{
    const long long row = (long long)blockIdx.x * kPadRows + threadIdx.y;
    if (row >= nrows) return;
    const int z = z0 + (int)(row / g.sy), y = (int)(row % g.sy);
    const float* in = pad + ((size_t)(z + g.oz) * g.py + (y + g.oy)) * g.px + g.ox;
    float* out = dst + ((size_t)z * g.sy + y) * g.sx;
    for (int x = threadIdx.x; x < g.sx; x += kPadThreadsX) out[x] = in[x];
}

}  // namespace

void run_pad_embed(const float* d_src, float* d_pad, const PadGeom& g, int pz0, int pn, cudaStream_t st)
{
    if (pn <= 0) return;
    const long long nrows = (long long)pn * g.py;
    const dim3 block(kPadThreadsX, kPadRows);
    const unsigned grid = (unsigned)((nrows + kPadRows - 1) / kPadRows);
    pad_embed_kernel<<<grid, block, 0, st>>>(d_src, d_pad, g, pz0, nrows);
    FC_CUDA_KERNEL();
    count_launches(1);
}

void run_pad_crop(const float* d_pad, float* d_dst, const PadGeom& g, int z0, int n, cudaStream_t st)
{
    if (n <= 0) return;
    const long long nrows = (long long)n * g.sy;
    const dim3 block(kPadThreadsX, kPadRows);
    const unsigned grid = (unsigned)((nrows + kPadRows - 1) / kPadRows);
    pad_crop_kernel<<<grid, block, 0, st>>>(d_pad, d_dst, g, z0, nrows);
    FC_CUDA_KERNEL();
    count_launches(1);
}

// ---- padded extents (pure host) -------------------------------------------------------------------
static bool seven_smooth(int v)
{
    for (int f : {2, 3, 5, 7})
        while (v % f == 0) v /= f;
    return v == 1;
}

void padded_extents(const int* imDim, const int* kernelDim, int policy, int* padDim)
{
    for (int i = 0; i < 3; ++i) {
        // reference zero_padd: extent = image + 2*(kernel/2), tests/padd_utils.h:12-24,99-108
        int e = imDim[i] + 2 * (kernelDim[i] / 2);
        if (policy == 1) {
            // next 7-smooth size every kernel of this library runs without the generic O(p) radix; the
            // fastest extent is kept even so that the x pass is a half-length complex transform
            while (!seven_smooth(e) || (i == 0 && (e & 1))) ++e;
        }
        padDim[i] = e;
    }
}

}  // namespace fcb200
