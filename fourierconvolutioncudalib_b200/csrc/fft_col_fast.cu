// Fast strided-axis (y / z) passes for sm_100a: the first stage reads its butterfly inputs straight
// from global memory into registers, the last stage writes its outputs straight to global memory,
// so the shared-memory tile is touched by only 2(ns-1) tile transfers instead of 2ns+2.
//
// Tile = L positions x 16 pencils (8 float4 column pairs, one 128-byte line per position).  Each
// quarter-warp (8 lanes = 8 column pairs) reads or writes one full 128-byte line of global memory
// per access, whatever the row order, so natural frequency order in global memory is free.
//
// Fused mode (image z pass): forward stages, then  last forward stage -> x H * 1/N -> first inverse
// stage  in registers (modulateAndNormalize_kernel of /root/reference/src/convolution3Dfft.cu:41-62
// fused between the two transforms), then the inverse stages.  The image spectrum makes no extra
// trip through HBM.
//
// Loads are issued in batches of ~16 float4 per thread before any arithmetic so that a CTA keeps
// its whole tile in flight (HBM latency x bandwidth needs ~44 KB in flight per SM).
#include "fft_engine.cuh"
#include "fft_kernels.h"

#include <algorithm>
#include <cstdlib>

namespace fcb200 {

namespace {

// U butterflies of radix R are loaded together: about LOADS float4 in flight per thread
template <int R, int LOADS>
struct Batch {
    static constexpr int U = (LOADS / R) < 1 ? 1 : (LOADS / R);
};

__device__ __forceinline__ float4 ldg_stream(const float2* p)
{
    return __ldcs(reinterpret_cast<const float4*>(p));   // read-once data: evict-first
}

template <int R>
__device__ __forceinline__ void split(const float4* v, p2* r, p2* i)
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        r[k] = make_float2(v[k].x, v[k].y);   // pair-planar: real parts of the two pencils
        i[k] = make_float2(v[k].z, v[k].w);
    }
}

__device__ __forceinline__ float4 join(p2 r, p2 i) { return make_float4(r.x, r.y, i.x, i.y); }

// ---- first forward stage: global (natural rows) -> registers -> smem -----------------------------
template <int R, int LOADS>
__device__ __forceinline__ void first_fwd(const float2* __restrict__ base, long long stride, float4* __restrict__ sm,
                                          const float4* __restrict__ tw, int L, int cp, int w, int W,
                                          const unsigned char* __restrict__ rowMask)
{
    constexpr int U = Batch<R, LOADS>::U;
    const int S = L / R;  // Li == L, beta == 0, j == b
    for (int b0 = w; b0 < S; b0 += U * W) {
        float4 v[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = b0 + u * W;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                const int row = j + k * S;
                v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                if (j < S && (rowMask == nullptr || rowMask[row])) v[u][k] = ldg_stream(base + (size_t)row * stride);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = b0 + u * W;
            if (j < S) {
                p2 r[R], i[R];
                split<R>(v[u], r, i);
                Dft<R>::run(r, i);
#pragma unroll
                for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * m]);
#pragma unroll
                for (int m = 0; m < R; ++m) sm[(j + m * S) * 8 + cp] = join(r[m], i[m]);
            }
        }
    }
}

// ---- first inverse stage: global (natural rows) -> registers -> smem (positions) ---------------
template <int R, int LOADS>
__device__ __forceinline__ void first_inv(const float2* __restrict__ base, long long stride, float4* __restrict__ sm,
                                          const int* __restrict__ rev, int L, int cp, int w, int W)
{
    constexpr int U = Batch<R, LOADS>::U;
    const int nb = L / R;
    const int fs = L / R;  // frequency step between the R inputs of one butterfly
    for (int b0 = w; b0 < nb; b0 += U * W) {
        float4 v[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * W;
            if (b < nb) {
                const int k0 = __ldg(rev + b * R);
#pragma unroll
                for (int k = 0; k < R; ++k) v[u][k] = ldg_stream(base + (size_t)(k0 + k * fs) * stride);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * W;
            if (b < nb) {
                p2 r[R], i[R];
                split<R>(v[u], r, i);
                Dft<R>::run(i, r);
#pragma unroll
                for (int m = 0; m < R; ++m) sm[(b * R + m) * 8 + cp] = join(r[m], i[m]);
            }
        }
    }
}

// ---- last forward stage: smem (positions) -> registers -> global (natural rows) ----------------
template <int R>
__device__ __forceinline__ void last_fwd(float2* __restrict__ base, long long stride, const float4* __restrict__ sm,
                                         const int* __restrict__ rev, int L, int cp, int w, int W)
{
    const int nb = L / R;
    const int fs = L / R;
    for (int b = w; b < nb; b += W) {
        p2 r[R], i[R];
        const int k0 = __ldg(rev + b * R);
        load_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
        Dft<R>::run(r, i);
#pragma unroll
        for (int m = 0; m < R; ++m)
            *reinterpret_cast<float4*>(base + (size_t)(k0 + m * fs) * stride) = join(r[m], i[m]);
    }
}

// ---- last inverse stage: smem -> registers -> global (natural rows) -----------------------------
template <int R>
__device__ __forceinline__ void last_inv(float2* __restrict__ base, long long stride, const float4* __restrict__ sm,
                                         const float4* __restrict__ tw, int L, int cp, int w, int W)
{
    const int S = L / R;  // Li == L, j == b
    for (int j = w; j < S; j += W) {
        p2 r[R], i[R];
        load_pairs<R>(sm, j * 8 + cp, S * 8, r, i);
#pragma unroll
        for (int k = 1; k < R; ++k) cmulc(r[k], i[k], tw[j * k]);
        Dft<R>::run(i, r);
#pragma unroll
        for (int m = 0; m < R; ++m)
            *reinterpret_cast<float4*>(base + (size_t)(j + m * S) * stride) = join(r[m], i[m]);
    }
}

// ---- fused middle: last forward stage, multiply by H and 1/N, first inverse stage --------------
template <int R, int LOADS>
__device__ __forceinline__ void mid_fused(const float2* __restrict__ hbase, long long stride, float4* __restrict__ sm,
                                          const int* __restrict__ rev, int L, int cp, int w, int W, float c)
{
    constexpr int U = Batch<R, LOADS>::U;
    const int nb = L / R;
    const int fs = L / R;
    for (int b0 = w; b0 < nb; b0 += U * W) {
        float4 h[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * W;
            if (b < nb) {
                const int k0 = __ldg(rev + b * R);
#pragma unroll
                for (int k = 0; k < R; ++k) h[u][k] = ldg_stream(hbase + (size_t)(k0 + k * fs) * stride);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = b0 + u * W;
            if (b < nb) {
                p2 r[R], i[R];
                load_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
                Dft<R>::run(r, i);
                // Dst = c * (Src * Dst), Src = PSF spectrum (reference mulAndScale, :41-45)
#pragma unroll
                for (int m = 0; m < R; ++m) {
                    const p2 hr = make_float2(h[u][m].x, h[u][m].y), hi = make_float2(h[u][m].z, h[u][m].w);
                    const p2 xr = pmuls(pfma(hr, r[m], pneg(pmul(hi, i[m]))), c);
                    const p2 xi = pmuls(pfma(hi, r[m], pmul(hr, i[m])), c);
                    r[m] = xr;
                    i[m] = xi;
                }
                Dft<R>::run(i, r);
                store_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
            }
        }
    }
}

#define FC_RADIX_SWITCH(R, CALL)            \
    switch (R) {                            \
        case 2: { constexpr int RR = 2; CALL; } break; \
        case 3: { constexpr int RR = 3; CALL; } break; \
        case 4: { constexpr int RR = 4; CALL; } break; \
        case 5: { constexpr int RR = 5; CALL; } break; \
        case 6: { constexpr int RR = 6; CALL; } break; \
        case 7: { constexpr int RR = 7; CALL; } break; \
        case 9: { constexpr int RR = 9; CALL; } break; \
        case 10: { constexpr int RR = 10; CALL; } break; \
        case 12: { constexpr int RR = 12; CALL; } break; \
        case 15: { constexpr int RR = 15; CALL; } break; \
        case 16: { constexpr int RR = 16; CALL; } break; \
        default: { constexpr int RR = 8; CALL; } break; \
    }

template <bool INV>
__device__ __forceinline__ void mid_stage(int R, float4* sm, const float4* tw, int L, int Li, int cp, int w, int W)
{
    FC_RADIX_SWITCH(R, (stage_smem<RR, INV>(sm, tw, L, Li, cp, w, W, 8)));
}

}  // namespace

// MODE 0 forward, 1 inverse, 2 fused (forward, x H x scale, inverse).  Requires a plan with
// ns >= 2 stages, all radices in {2,3,4,5,7,8}, and a tile of 8 column pairs.
template <int MODE, int MAXT, int MINB, int LOADS>
__global__ void __launch_bounds__(MAXT, MINB) col_fast_kernel(ColArgs a)
{
    extern __shared__ float4 smem[];
    const int L = a.P.L;
    const int ns = a.P.ns;
    float4* sm = smem;
    float4* tw_s = sm + (size_t)L * 8;

    const int t = threadIdx.x;
    const int cp = t & 7, w = t >> 3, W = blockDim.x >> 3;
    const int gi = blockIdx.x / a.tilesPerGroup;
    const int tt = blockIdx.x - gi * a.tilesPerGroup;
    const long long group = a.groupList ? (long long)a.groupList[gi] : (long long)gi;
    const int col0 = tt * 16;
    const int npairs = min(8, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    const size_t off = (size_t)group * a.groupStride + col0 + 2 * cp;
    float2* base = a.data + off;

    load_twiddles(tw_s, a.P.tw, L);
    __syncthreads();

    const int Rf = a.P.radix[0];
    const int Rl = a.P.radix[ns - 1];

    if (MODE == 0 || MODE == 2) {
        if (active) FC_RADIX_SWITCH(Rf, (first_fwd<RR, LOADS>(base, a.stride, sm, tw_s, L, cp, w, W, a.rowMask)));
        __syncthreads();
        int Li = L / Rf;
        for (int s = 1; s < ns - 1; ++s) {
            const int R = a.P.radix[s];
            if (active) mid_stage<false>(R, sm, tw_s, L, Li, cp, w, W);
            Li /= R;
            __syncthreads();
        }
        if (MODE == 0) {
            if (active) FC_RADIX_SWITCH(Rl, (last_fwd<RR>(base, a.stride, sm, a.P.rev, L, cp, w, W)));
            return;
        }
        if (active) FC_RADIX_SWITCH(Rl, (mid_fused<RR, LOADS>(a.H + off, a.stride, sm, a.P.rev, L, cp, w, W, a.scale)));
        __syncthreads();
    } else {
        if (active) FC_RADIX_SWITCH(Rl, (first_inv<RR, LOADS>(base, a.stride, sm, a.P.rev, L, cp, w, W)));
        __syncthreads();
    }
    // inverse stages ns-2 .. 1 in shared memory, then stage 0 to global
    int Li = Rl;
    for (int s = ns - 2; s >= 1; --s) {
        const int R = a.P.radix[s];
        Li *= R;
        if (active) mid_stage<true>(R, sm, tw_s, L, Li, cp, w, W);
        __syncthreads();
    }
    if (active) FC_RADIX_SWITCH(Rf, (last_inv<RR>(base, a.stride, sm, tw_s, L, cp, w, W)));
}

bool col_fast_supported(const AxisPlanDev& P)
{
    if (P.generic || P.ns < 2) return false;
    if (P.big) return false;   // the odd primes 11..23 and the fat composite radices run on the all-shared-memory kernel (fft_kernels.cu)
    const size_t need = (size_t)P.L * 8 * sizeof(float4) + (size_t)P.L * sizeof(float4);
    return need <= (size_t)kMaxDynSmem;
}

// Tuning knobs (read once from the environment; defaults chosen from the sweep in profiles/):
//   FCB200_COL_VARIANT  0: <=128 regs, 16 loads in flight   1: <=80 regs, 8 loads   2: <=64 regs, 8 loads (512 thr)
//   FCB200_COL_THREADS  CTA size (multiple of 32); 0 = derived from the transform length
static int env_int(const char* name, int dflt)
{
    const char* e = std::getenv(name);
    return e ? std::atoi(e) : dflt;
}

void launch_col_fast(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    static const int variant = env_int("FCB200_COL_VARIANT", 0);
    static const int threads_env = env_int("FCB200_COL_THREADS", 0);
    const size_t smem = (size_t)a.P.L * 8 * sizeof(float4) + (size_t)a.P.L * sizeof(float4);
    const long long grid = ngroups * a.tilesPerGroup;
    if (grid == 0) return;
    if (grid > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    int threads = threads_env;
    if (threads <= 0) {
        // one radix-8 butterfly per worker and stage: L/8 workers of 8 lanes, within [128, maxT]
        threads = ((a.P.L + 31) / 32) * 32;
    }
    (void)variant;
    const int maxT = 256;
    threads = std::max(64, std::min(threads, maxT));
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024)
            FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        kernel<<<(unsigned)grid, threads, smem, st>>>(a);
    };
    // (the <=80- and <=64-register variants of round 1 spilled in the radix-16 / 15 / 12 arms and never won a sweep:
    // removed; this build keeps at most 92 bytes of spill stores)
    switch (mode) {
        case 0: go(col_fast_kernel<0, 256, 2, 16>); break;
        case 1: go(col_fast_kernel<1, 256, 2, 16>); break;
        default: go(col_fast_kernel<2, 256, 2, 16>); break;
    }
    FC_CUDA_KERNEL();
}

// ------------------------------------------------------------------------------------------------
// Input-pruned PSF z pass: the placed PSF is non-zero on at most 16 consecutive z planes (mod L), so
//   X[k1*(L/16) + k2] = w16^(z0 k1) * sum_{n<16} [ x[z0+n] w_L^((z0+n) k2) ] w16^(n k1)
// i.e. one radix-16 butterfly per output residue k2 on the 16 twiddled inputs: no shared-memory
// stages, no barriers after the 2 KB window is staged.  L % 16 == 0.
// ------------------------------------------------------------------------------------------------
// MAXT: largest CTA (L / 2 threads): 512 for L <= 1024 leaves 128 registers per thread (the 64 of a 1024-thread bound
// spilled 150-380 bytes of the 16-point butterfly)
template <int NH, int MAXT>   // window of 16 * NH planes
__global__ void __launch_bounds__(MAXT) psf_z_pruned_kernel(ColArgs a, int z0)
{
    __shared__ float4 win[16 * NH * 8];
    const int L = a.P.L, Q = L / 16;
    const int t = threadIdx.x, cp = t & 7, w = t >> 3;   // w = residue k2 in [0, Q)
    const int col0 = blockIdx.x * 16;
    const int npairs = min(8, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    float2* base = a.data + col0 + 2 * cp;
    const size_t stride = (size_t)a.stride;

    for (int q = t; q < 128 * NH; q += blockDim.x) {
        const int n = q >> 3;
        int z = z0 + n;
        if (z >= L) z -= L;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((q & 7) < npairs && (a.rowMask == nullptr || a.rowMask[z]))
            v = *reinterpret_cast<const float4*>(a.data + col0 + 2 * (q & 7) + (size_t)z * stride);
        win[q] = v;
    }
    __syncthreads();
    if (!active) return;

    // sum over the 16-plane sub-windows first: w16^(n k1) only depends on n mod 16
    p2 r[16], i[16];
    int e = (int)(((long long)z0 * w) % L);   // exponent (z0 + n) * k2 mod L, advanced by k2 per input
#pragma unroll
    for (int n = 0; n < 16; ++n) {
        const float4 v = win[n * 8 + cp];
        const float2 tw = __ldg(a.P.tw + e);
        r[n] = make_float2(v.x, v.y);
        i[n] = make_float2(v.z, v.w);
        cmul(r[n], i[n], make_float4(tw.x, tw.x, tw.y, tw.y));
        e += w;
        if (e >= L) e -= L;
    }
#pragma unroll
    for (int h = 1; h < NH; ++h) {
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float4 v = win[(16 * h + n) * 8 + cp];
            const float2 tw = __ldg(a.P.tw + e);
            p2 tr = make_float2(v.x, v.y), ti = make_float2(v.z, v.w);
            cmul(tr, ti, make_float4(tw.x, tw.x, tw.y, tw.y));
            r[n] = padd(r[n], tr);
            i[n] = padd(i[n], ti);
            e += w;
            if (e >= L) e -= L;
        }
    }
    Dft<16>::run(r, i);
    const int s16 = z0 & 15;   // w16^(z0 k1) = w_L^((z0 k1 mod 16) * Q)
#pragma unroll
    for (int k1 = 0; k1 < 16; ++k1) {
        if (s16 != 0 && k1 != 0) {
            const float2 tw = __ldg(a.P.tw + ((s16 * k1) & 15) * Q);
            cmul(r[k1], i[k1], make_float4(tw.x, tw.x, tw.y, tw.y));
        }
        *reinterpret_cast<float4*>(base + (size_t)(k1 * Q + w) * stride) = make_float4(r[k1].x, r[k1].y, i[k1].x, i[k1].y);
    }
}

// The same for lengths that are not multiples of 16 (the caller-padded 300, 420, 270, 350 ...): L = R * Q with a register
// radix R (20, 18, 24, 28, 21, 25) and a window of WP = 16 / 32 / 64 planes; inputs that are R planes apart share their
// w_R^(n k1), so they are summed (twiddled) into one butterfly input first, inputs beyond the window are zero:
//   X[k1*Q + k2] = w_R^(z0 k1) * sum_{m<R} [ sum_{n = m mod R, n < WP} x[z0+n] w_L^((z0+n) k2) ] w_R^(m k1)
// (these lengths ran a full masked z pass before: 560x560x300 PSF z pass 0.134 ms for 382 MB written)
template <int R, int WP>
__global__ void __launch_bounds__(256) psf_z_pruned_r_kernel(ColArgs a, int z0)
{
    __shared__ float4 win[WP * 8];
    const int L = a.P.L, Q = L / R;
    const int t = threadIdx.x, cp = t & 7, w = t >> 3;   // w = residue k2 in [0, Q)
    const int col0 = blockIdx.x * 16;
    const int npairs = min(8, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    float2* base = a.data + col0 + 2 * cp;
    const size_t stride = (size_t)a.stride;
    for (int q = t; q < WP * 8; q += blockDim.x) {
        int z = z0 + (q >> 3);
        if (z >= L) z -= L;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if ((q & 7) < npairs && (a.rowMask == nullptr || a.rowMask[z]))
            v = *reinterpret_cast<const float4*>(a.data + col0 + 2 * (q & 7) + (size_t)z * stride);
        win[q] = v;
    }
    __syncthreads();
    if (!active || w >= Q) return;
    p2 r[R], i[R];
    int e = (int)(((long long)z0 * w) % L);
#pragma unroll
    for (int n = 0; n < (WP > R ? WP : R); ++n) {
        if (n < WP) {
            const float4 v = win[n * 8 + cp];
            const float2 tw = __ldg(a.P.tw + e);
            p2 tr = make_float2(v.x, v.y), ti = make_float2(v.z, v.w);
            cmul(tr, ti, make_float4(tw.x, tw.x, tw.y, tw.y));
            if (n < R) {
                r[n] = tr;
                i[n] = ti;
            } else {
                r[n % R] = padd(r[n % R], tr);
                i[n % R] = padd(i[n % R], ti);
            }
            e += w;
            if (e >= L) e -= L;
        } else {
            r[n] = make_float2(0.f, 0.f);
            i[n] = make_float2(0.f, 0.f);
        }
    }
    Dft<R>::run(r, i);
    const int sR = z0 % R;   // w_R^(z0 k1) = w_L^(((z0 k1) mod R) * Q)
#pragma unroll
    for (int k1 = 0; k1 < R; ++k1) {
        if (sR != 0 && k1 != 0) {
            const float2 tw = __ldg(a.P.tw + ((sR * k1) % R) * Q);
            cmul(r[k1], i[k1], make_float4(tw.x, tw.x, tw.y, tw.y));
        }
        *reinterpret_cast<float4*>(base + (size_t)(k1 * Q + w) * stride) = make_float4(r[k1].x, r[k1].y, i[k1].x, i[k1].y);
    }
}

template <int R>
static bool try_psf_z_pruned_r(const ColArgs& a, int z0, int planes, cudaStream_t st)
{
    const int L = a.P.L;
    if (L % R != 0 || planes > L) return false;
    const int threads = ((L / R * 8 + 31) / 32) * 32;
    if (threads > 256 || L / R < 2) return false;
    const int grid = (a.rowLen + 15) / 16;
    if (planes == 16) psf_z_pruned_r_kernel<R, 16><<<grid, threads, 0, st>>>(a, z0);
    else if (planes == 32) psf_z_pruned_r_kernel<R, 32><<<grid, threads, 0, st>>>(a, z0);
    else if (planes == 64) psf_z_pruned_r_kernel<R, 64><<<grid, threads, 0, st>>>(a, z0);
    else return false;
    FC_CUDA_KERNEL();
    return true;
}

// planes: size of the window (16, 32 or 64 consecutive planes from z0, mod L) that holds every non-zero input plane
bool launch_psf_z_pruned(const ColArgs& a, int z0, int planes, cudaStream_t st)
{
    static const bool on = env_int("FCB200_PSF_PRUNED", 1) != 0;
    const int L = a.P.L;
    if (on && L % 16 != 0 && a.groupStride == 0 && a.rowLen > 0)
        return try_psf_z_pruned_r<20>(a, z0, planes, st) || try_psf_z_pruned_r<18>(a, z0, planes, st) ||
               try_psf_z_pruned_r<24>(a, z0, planes, st) || try_psf_z_pruned_r<28>(a, z0, planes, st) ||
               try_psf_z_pruned_r<21>(a, z0, planes, st) || try_psf_z_pruned_r<25>(a, z0, planes, st);
    if (!on || L % 16 != 0 || L / 16 * 8 > 1024 || L / 16 * 8 < 128 || a.groupStride != 0) return false;
    if ((planes != 16 && planes != 32 && planes != 64) || planes > L) return false;
    const int tiles = (a.rowLen + 15) / 16;
    if (tiles == 0) return true;
    const int threads = L / 16 * 8;
    if (threads <= 512) {
        if (planes == 16) psf_z_pruned_kernel<1, 512><<<tiles, threads, 0, st>>>(a, z0);
        else if (planes == 32) psf_z_pruned_kernel<2, 512><<<tiles, threads, 0, st>>>(a, z0);
        else psf_z_pruned_kernel<4, 512><<<tiles, threads, 0, st>>>(a, z0);
    } else {
        if (planes == 16) psf_z_pruned_kernel<1, 1024><<<tiles, threads, 0, st>>>(a, z0);
        else if (planes == 32) psf_z_pruned_kernel<2, 1024><<<tiles, threads, 0, st>>>(a, z0);
        else psf_z_pruned_kernel<4, 1024><<<tiles, threads, 0, st>>>(a, z0);
    }
    FC_CUDA_KERNEL();
    return true;
}

}  // namespace fcb200
