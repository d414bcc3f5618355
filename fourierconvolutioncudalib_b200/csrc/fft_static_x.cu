// Compile-time specialised x-pass kernels: the transposing two-tile kernels (fft_xpass.cuh) for static
// plans and the register-resident row-wise kernels (fft_xrow.cuh, fft_xrowg.cuh), with their dispatch.
#include "fft_xpass.cuh"
#include "fft_xrow.cuh"
#include "fft_xrowg.cuh"
#include "fft_static_plans.h"

namespace fcb200 {

namespace {

// FCB200_XTILE_PERSIST=n > 0: the tiled x kernels are launched with n x (resident CTAs) CTAs that walk over the tiles
// (twiddle table loaded once per CTA); 0 (default): one tile of 16 rows per CTA.  Measured: no clear gain, unlike the
// row-wise kernels (560: x forward 0.238 -> 0.248 ms, 420: 0.220 -> 0.212, 270: 0.063 -> 0.060) -- these kernels are
// bound by their two-tile structure, not by the table loads.
template <typename K>
long long x_tile_grid(K kernel, int threads, size_t smem, long long tiles, bool psf)
{
    static const int persist = env_int("FCB200_XTILE_PERSIST", 0);
    if (persist <= 0 || psf) return tiles;
    int per_sm = 1, dev = 0, sms = 148;
    FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, smem));
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    return std::min<long long>(tiles, (long long)sms * std::max(1, per_sm) * persist);
}

template <class P, int THREADS>
bool try_x_fwd(const XArgs& a, bool psf, long long tiles, cudaStream_t st)
{
    if (!plan_matches<P>(a.P)) return false;
    const size_t smem = x_smem_bytes(a.g, a.P);
    if (smem > (size_t)kMaxDynSmem) return false;
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024)
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        launch_pdl(a.pdl != 0, kernel, dim3((unsigned)x_tile_grid(kernel, THREADS, smem, tiles, psf)), dim3(THREADS), smem, st, a);
        FC_CUDA_KERNEL();
    };
    if (psf) go(x_fwd_kernel<1, P, THREADS>);
    else go(x_fwd_kernel<0, P, THREADS>);
    return true;
}

template <class P, int THREADS>
bool try_x_inv(const XArgs& a, long long tiles, cudaStream_t st)
{
    if (!plan_matches<P>(a.P)) return false;
    const size_t smem = x_smem_bytes(a.g, a.P);
    if (smem > (size_t)kMaxDynSmem) return false;
    auto kernel = x_inv_kernel<P, THREADS>;
    if (smem > 48 * 1024) FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(a.pdl != 0, kernel, dim3((unsigned)x_tile_grid(kernel, THREADS, smem, tiles, false)), dim3(THREADS), smem, st, a);
    FC_CUDA_KERNEL();
    return true;
}


}  // namespace

static int xt256()
{
    static const int v = env_int("FCB200_XT256", 256);
    return v;
}

// register-resident row-wise kernels (fft_xrow.cuh) for plans (R, R); FCB200_XROW=0 disables them
static bool xrow_enabled()
{
    static const bool on = env_int("FCB200_XROW", 1) != 0;
    return on;
}

template <int R, int THREADS>
static bool try_xrow(const XArgs& a, bool inverse, cudaStream_t st)
{
    if (a.P.L != R * R || a.P.ns != 2 || a.P.radix[0] != R || a.P.radix[1] != R || a.g.odd || a.rowList) return false;
    constexpr int RP = THREADS / R;
    const long long grid = (a.nrows + 2 * RP - 1) / (2 * RP);
    if (grid == 0) return true;
    if (grid > 0x7fffffffLL) return false;
    const size_t smem = (size_t)(R * R + RP * XRow<R>::PADM) * sizeof(float4);
    // FCB200_XROW_PERSIST=n > 0: persistent launch with n x (resident CTAs) CTAs, each walking over blocks of RP row pairs
    // and building its twiddle table once; 0: one block of rows per CTA.  C3 (nx = 512): x forward 0.1045 -> 0.1001 ms,
    // x inverse 0.1024 -> 0.0942 ms with n = 1 (n = 2: 0.1015 / 0.0957, n = 4: 0.1029 / 0.0995)
    static const int persist = env_int("FCB200_XROW_PERSIST", 1);
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024)
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long launch = grid;
        if (persist > 0) {
            int per_sm = 1, dev = 0, sms = 148;
            FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            launch = std::min<long long>(grid, (long long)sms * std::max(1, per_sm) * persist);
        }
        launch_pdl(a.pdl != 0, kernel, dim3((unsigned)launch), dim3(THREADS), smem, st, a);
        FC_CUDA_KERNEL();
    };
    if (inverse) go(xrow_inv_kernel<R, THREADS>);
    else go(xrow_fwd_kernel<R, THREADS>);
    return true;
}

template <int R0, int R1, int R2, int THREADS>
static bool try_xrowg(const XArgs& a, bool inverse, cudaStream_t st)
{
    typedef XRowG<R0, R1, R2> G;
    // three-stage plans read the planner's digit-reversal table, so the radices must be the planner's
    if (a.P.L != G::M || a.g.odd || a.rowList) return false;
    if (G::NS == 3 && (a.P.ns != 3 || a.P.radix[0] != R0 || a.P.radix[1] != R1 || a.P.radix[2] != R2)) return false;
    constexpr int RP = THREADS / G::TG;
    static_assert(RP >= 1 && (G::TG <= 32 || RP <= 15), "row pairs per CTA (named barriers 1..15)");
    const long long grid = (a.nrows + 2 * RP - 1) / (2 * RP);
    if (grid == 0) return true;
    if (grid > 0x7fffffffLL) return false;
    const size_t smem = (size_t)(G::TW1 + G::TW2 + RP * G::PADM) * sizeof(float4);
    // persistent for rows of 384 voxels and more (M >= 192; nx = 2048: the twiddle tables are as large as the rows of a CTA;
    // 384^3: x passes 0.106 / 0.099 -> 0.099 / 0.096 ms; no effect at M = 128): as many CTAs as fit,
    // each walking over blocks of RP row pairs.  FCB200_XROWG_PERSIST=0: one block of rows per CTA.
    static const int persist = env_int("FCB200_XROWG_PERSIST", 1);
    static const int persist_min_m = env_int("FCB200_XROWG_PERSIST_MINM", 192);
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024)
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        long long launch = grid;
        if (persist && G::M >= persist_min_m) {
            int per_sm = 1, dev = 0, sms = 148;
            FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
            cudaGetDevice(&dev);
            cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
            launch = std::min<long long>(grid, (long long)sms * std::max(1, per_sm));
        }
        launch_pdl(a.pdl != 0, kernel, dim3((unsigned)launch), dim3(THREADS), smem, st, a);
        FC_CUDA_KERNEL();
    };
    if (inverse) go(xrowg_inv_kernel<R0, R1, R2, THREADS>);
    else go(xrowg_fwd_kernel<R0, R1, R2, THREADS>);
    return true;
}

static bool try_xrowg_all(const XArgs& a, bool inverse, cudaStream_t st)
{
    static const int x384 = env_int("FCB200_XROW384", 1);
    if (x384 == 64 && try_xrowg<24, 8, 1, 64>(a, inverse, st)) return true;
    if (x384 == 256 && try_xrowg<24, 8, 1, 256>(a, inverse, st)) return true;
    if (x384 != 0 && try_xrowg<24, 8, 1, 128>(a, inverse, st)) return true;   // nx = 384
    // nx = 256: CTA sizes 32 / 128 / 256 and persistent launches were measured (256^3: x passes 0.045 / 0.033 ms with every
    // variant, profiles/r02_notes.md): the pass is L2-resident and launch-bound, 64 threads stay
    return try_xrowg<16, 8, 1, 64>(a, inverse, st) ||      // nx = 256
           try_xrowg<16, 4, 8, 128>(a, inverse, st) ||     // nx = 1024 (x-axis planning style)
           try_xrowg<8, 8, 8, 128>(a, inverse, st) ||
           try_xrowg<16, 8, 8, 128>(a, inverse, st);       // nx = 2048 (x-axis planning style)
}

static int xrow_threads()
{
    static const int v = env_int("FCB200_XROW_T", 128);
    return v;
}

// Fused z pass with on-the-fly PSF spectrum; probe = only report whether a kernel exists for the plan.
bool launch_x_fwd_static(const XArgs& a, bool psf, cudaStream_t st)
{
    if (!static_enabled()) return false;
    const long long tiles = (a.nrows + 15) / 16;
    if (tiles == 0) return true;
    if (!psf && xrow_enabled() && !a.padOn) {
        if (xrow_threads() == 64 && try_xrow<16, 64>(a, false, st)) return true;
        if (xrow_threads() == 256 && try_xrow<16, 256>(a, false, st)) return true;
        if (try_xrow<16, 128>(a, false, st) || try_xrow<8, 64>(a, false, st)) return true;
        if (try_xrowg_all(a, false, st)) return true;
    }
    return try_x_fwd<P32, 64>(a, psf, tiles, st) || try_x_fwd<P64, 64>(a, psf, tiles, st) ||
           try_x_fwd<P128, 128>(a, psf, tiles, st) || try_x_fwd<P192, 192>(a, psf, tiles, st) ||
           (xt256() == 128 && try_x_fwd<P256, 128>(a, psf, tiles, st)) ||
           (xt256() == 512 && try_x_fwd<P256, 512>(a, psf, tiles, st)) ||
           try_x_fwd<P256, 256>(a, psf, tiles, st) || try_x_fwd<P512, 256>(a, psf, tiles, st) ||
           try_x_fwd<P512x, 256>(a, psf, tiles, st) ||
           try_x_fwd<P1024, 512>(a, psf, tiles, st) || try_x_fwd<P280, 256>(a, psf, tiles, st) ||

           try_x_fwd<P224, 256>(a, psf, tiles, st) || try_x_fwd<P210, 256>(a, psf, tiles, st) ||
           try_x_fwd<P150, 128>(a, psf, tiles, st) || try_x_fwd<P135, 128>(a, psf, tiles, st);
}

bool launch_x_inv_static(const XArgs& a, cudaStream_t st)
{
    if (!static_enabled()) return false;
    const long long tiles = (a.nrows + 15) / 16;
    if (tiles == 0) return true;
    if (xrow_enabled() && !a.padOn) {
        if (xrow_threads() == 64 && try_xrow<16, 64>(a, true, st)) return true;
        if (xrow_threads() == 256 && try_xrow<16, 256>(a, true, st)) return true;
        if (try_xrow<16, 128>(a, true, st) || try_xrow<8, 64>(a, true, st)) return true;
        if (try_xrowg_all(a, true, st)) return true;
    }
    return try_x_inv<P32, 64>(a, tiles, st) || try_x_inv<P64, 64>(a, tiles, st) || try_x_inv<P128, 128>(a, tiles, st) ||
           try_x_inv<P192, 192>(a, tiles, st) || (xt256() == 128 && try_x_inv<P256, 128>(a, tiles, st)) ||
           (xt256() == 512 && try_x_inv<P256, 512>(a, tiles, st)) || try_x_inv<P256, 256>(a, tiles, st) ||
           try_x_inv<P512, 256>(a, tiles, st) || try_x_inv<P512x, 256>(a, tiles, st) || try_x_inv<P1024, 512>(a, tiles, st) ||
           try_x_inv<P280, 256>(a, tiles, st) || try_x_inv<P224, 256>(a, tiles, st) || try_x_inv<P210, 256>(a, tiles, st) ||

           try_x_inv<P150, 128>(a, tiles, st) || try_x_inv<P135, 128>(a, tiles, st);
}


}  // namespace fcb200
