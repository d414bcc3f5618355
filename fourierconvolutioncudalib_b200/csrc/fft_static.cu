// Compile-time specialised kernels (fft_static.cuh) for the hot transform lengths, and the
// dispatch that picks them when the run-time plan has exactly the same radix sequence.
// Anything else falls back to the run-time-radix kernels (fft_col_fast.cu / fft_kernels.cu).
#include "fft_xpass.cuh"

#include <cstdlib>

namespace fcb200 {

namespace {

// MODE 0 forward, 1 inverse, 2 fused forward x H x scale inverse (see fft_col_fast.cu)
template <int MODE, class P, int THREADS, int U, bool MASKED, int TXP>
__global__ void __launch_bounds__(THREADS) col_static_kernel(ColArgs a, int tilesPerGroup)
{
    constexpr int L = P::L, NW = THREADS / TXP;
    extern __shared__ float4 smem[];
    float4* sm = smem;
    float4* tw = sm + (size_t)L * TXP;

    const int t = threadIdx.x;
    const int cp = t % TXP, w = t / TXP;
    const int gi = blockIdx.x / tilesPerGroup;
    const int tt = blockIdx.x - gi * tilesPerGroup;
    const long long group = a.groupList ? (long long)a.groupList[gi] : (long long)gi;
    const int col0 = tt * 2 * TXP;
    const int npairs = min(TXP, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    const size_t off = (size_t)group * a.groupStride + col0 + 2 * cp;
    float2* base = a.data + off;
    const size_t stride = (size_t)a.stride;

    load_twiddles(tw, a.P.tw, L);
    __syncthreads();

    if (MODE == 0 || MODE == 2) {
        if (active) sfirst_fwd<P::R0, L, NW, U, MASKED, TXP>(base, stride, sm, tw, cp, w, a.rowMask);
        __syncthreads();
        if constexpr (P::ns >= 3) {
            if (active) sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
            __syncthreads();
        }
        if constexpr (P::ns >= 4) {
            if (active) sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, TXP>(sm, tw, cp, w);
            __syncthreads();
        }
        if (MODE == 0) {
            if (active) slast_fwd<P::RL, L, NW, TXP>(base, stride, sm, a.P.rev, cp, w);
            return;
        }
        if (active) smid_fused<P::RL, L, NW, U, TXP>(a.H + off, stride, sm, a.P.rev, cp, w, a.scale);
        __syncthreads();
    } else {
        if (active) sfirst_inv<P::RL, L, NW, U, TXP>(base, stride, sm, a.P.rev, cp, w);
        __syncthreads();
    }
    if constexpr (P::ns >= 4) {
        if (active) sstage<P::R2, L, P::R2 * P::R3, NW, true, TXP>(sm, tw, cp, w);
        __syncthreads();
    }
    if constexpr (P::ns >= 3) {
        if (active) sstage<P::R1, L, P::R1 * P::R2 * P::R3, NW, true, TXP>(sm, tw, cp, w);
        __syncthreads();
    }
    if (active) slast_inv<P::R0, L, NW, TXP>(base, stride, sm, tw, cp, w);
}

template <class P>
bool plan_matches(const AxisPlanDev& d)
{
    if (d.L != P::L || d.ns != P::ns || d.generic) return false;
    const int r[4] = {P::R0, P::R1, P::R2, P::R3};
    for (int i = 0; i < P::ns; ++i)
        if (d.radix[i] != r[i]) return false;
    return true;
}

template <typename K>
void launch(K kernel, long long grid, int threads, size_t smem, cudaStream_t st, const ColArgs& a, int tpg)
{
    if (smem > 48 * 1024) FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)grid, threads, smem, st>>>(a, tpg);
    FC_CUDA_KERNEL();
}

// One static configuration: plan P, CTA size, load batch U, tile width TXP (column pairs per row).
template <class P, int THREADS, int U, int TXP>
void run_col(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const int tpg = (a.rowLen + 2 * TXP - 1) / (2 * TXP);
    const long long grid = ngroups * tpg;
    if (grid == 0) return;
    if (grid > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    const size_t smem = (size_t)P::L * TXP * sizeof(float4) + (size_t)P::L * sizeof(float4);
    if (mode == 0 && a.rowMask) launch(col_static_kernel<0, P, THREADS, U, true, TXP>, grid, THREADS, smem, st, a, tpg);
    else if (mode == 0) launch(col_static_kernel<0, P, THREADS, U, false, TXP>, grid, THREADS, smem, st, a, tpg);
    else if (mode == 1) launch(col_static_kernel<1, P, THREADS, U, false, TXP>, grid, THREADS, smem, st, a, tpg);
    else launch(col_static_kernel<2, P, THREADS, U, false, TXP>, grid, THREADS, smem, st, a, tpg);
}

template <class P, int THREADS>
bool try_x_fwd(const XArgs& a, bool psf, long long tiles, cudaStream_t st)
{
    if (!plan_matches<P>(a.P)) return false;
    const size_t smem = x_smem_bytes(a.g, a.P);
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024)
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<(unsigned)tiles, THREADS, smem, st>>>(a);
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
    auto kernel = x_inv_kernel<P, THREADS>;
    if (smem > 48 * 1024) FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    kernel<<<(unsigned)tiles, THREADS, smem, st>>>(a);
    FC_CUDA_KERNEL();
    return true;
}

bool static_enabled()
{
    static const bool on = [] {
        const char* e = std::getenv("FCB200_STATIC");
        return !(e && std::atoi(e) == 0);
    }();
    return on;
}

// The radix sequences are exactly what the planner (fc_plan.cu: factorize) produces.
typedef SPlan<32, 8, 4> P32;
typedef SPlan<64, 8, 8> P64;
typedef SPlan<128, 16, 8> P128;
typedef SPlan<192, 8, 8, 3> P192;
typedef SPlan<256, 16, 16> P256;
typedef SPlan<384, 16, 8, 3> P384;
typedef SPlan<512, 8, 8, 8> P512;
typedef SPlan<1024, 16, 16, 4> P1024;
// planning style 1 (fused z axis): L = 256 as (8,8,4)
typedef SPlan<256, 8, 8, 4> P256b;

}  // namespace

static int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

bool launch_col_static(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    if (!static_enabled() || a.txp != 8) return false;
    // tuning knob for the longest pencils (profiles/): CTA shape / tile width of the L = 512 kernels
    static const int v512 = env_int("FCB200_V512", 0);
    if (plan_matches<P64>(a.P)) run_col<P64, 64, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P128>(a.P)) run_col<P128, 64, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P256>(a.P)) run_col<P256, 128, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P256b>(a.P)) run_col<P256b, 128, 2, 8>(a, mode, ngroups, st);
    else if (plan_matches<P384>(a.P)) run_col<P384, 192, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P512>(a.P)) {
        switch (v512) {
            case 1: run_col<P512, 256, 1, 8>(a, mode, ngroups, st); break;
            case 2: run_col<P512, 128, 2, 8>(a, mode, ngroups, st); break;
            case 3: run_col<P512, 256, 2, 8>(a, mode, ngroups, st); break;
            case 4: run_col<P512, 128, 2, 4>(a, mode, ngroups, st); break;
            case 5: run_col<P512, 256, 1, 4>(a, mode, ngroups, st); break;
            case 6: run_col<P512, 128, 1, 4>(a, mode, ngroups, st); break;
            default: run_col<P512, 512, 1, 8>(a, mode, ngroups, st); break;
        }
    } else if (plan_matches<P1024>(a.P)) run_col<P1024, 512, 1, 8>(a, mode, ngroups, st);
    else return false;
    return true;
}

static int xt256()
{
    static const int v = env_int("FCB200_XT256", 256);
    return v;
}

bool launch_x_fwd_static(const XArgs& a, bool psf, cudaStream_t st)
{
    if (!static_enabled()) return false;
    const long long tiles = (a.nrows + 15) / 16;
    if (tiles == 0) return true;
    return try_x_fwd<P32, 64>(a, psf, tiles, st) || try_x_fwd<P64, 64>(a, psf, tiles, st) ||
           try_x_fwd<P128, 128>(a, psf, tiles, st) || try_x_fwd<P192, 192>(a, psf, tiles, st) ||
           (xt256() == 128 && try_x_fwd<P256, 128>(a, psf, tiles, st)) ||
           (xt256() == 512 && try_x_fwd<P256, 512>(a, psf, tiles, st)) ||
           try_x_fwd<P256, 256>(a, psf, tiles, st) || try_x_fwd<P512, 256>(a, psf, tiles, st) ||
           try_x_fwd<P1024, 512>(a, psf, tiles, st);
}

bool launch_x_inv_static(const XArgs& a, cudaStream_t st)
{
    if (!static_enabled()) return false;
    const long long tiles = (a.nrows + 15) / 16;
    if (tiles == 0) return true;
    return try_x_inv<P32, 64>(a, tiles, st) || try_x_inv<P64, 64>(a, tiles, st) || try_x_inv<P128, 128>(a, tiles, st) ||
           try_x_inv<P192, 192>(a, tiles, st) || (xt256() == 128 && try_x_inv<P256, 128>(a, tiles, st)) ||
           (xt256() == 512 && try_x_inv<P256, 512>(a, tiles, st)) || try_x_inv<P256, 256>(a, tiles, st) ||
           try_x_inv<P512, 256>(a, tiles, st) || try_x_inv<P1024, 512>(a, tiles, st);
}

}  // namespace fcb200
