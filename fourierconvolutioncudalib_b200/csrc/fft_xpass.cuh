// X-pass kernels (R2C / C2R along the fastest axis), shared by the run-time-radix build
// (fft_kernels.cu) and the compile-time specialised build (fft_static.cu).
#pragma once
#include "fft_static.cuh"
#include "fft_kernels.h"

namespace fcb200 {

struct DynPlan {};
struct DynPlanBig {};   // run-time radices including the register butterflies of the primes 11..23
template <class P> struct IsBigPlan { static constexpr bool value = false; };
template <> struct IsBigPlan<DynPlanBig> { static constexpr bool value = true; };

template <class P>
struct PlanLen {
    static __device__ __forceinline__ int get(const AxisPlanDev&) { return P::L; }
};
template <>
struct PlanLen<DynPlan> {
    static __device__ __forceinline__ int get(const AxisPlanDev& d) { return d.L; }
};
template <>
struct PlanLen<DynPlanBig> {
    static __device__ __forceinline__ int get(const AxisPlanDev& d) { return d.L; }
};

template <class P, int NW, bool INV>
struct PlanRun {
    static __device__ __forceinline__ float4* run(const AxisPlanDev&, float4* A, float4*, const float4* tw, int cp,
                                                  int w, int)
    {
        sengine<P, NW, INV>(A, tw, cp, w);
        return A;
    }
};
template <int NW, bool INV>
struct PlanRun<DynPlan, NW, INV> {
    static __device__ __forceinline__ float4* run(const AxisPlanDev& d, float4* A, float4* B, const float4* tw,
                                                  int cp, int w, int W)
    {
        return engine_run<INV>(d, A, B, tw, cp, w, W, 8, true);
    }
};
template <int NW, bool INV>
struct PlanRun<DynPlanBig, NW, INV> {
    static __device__ __forceinline__ float4* run(const AxisPlanDev& d, float4* A, float4* B, const float4* tw,
                                                  int cp, int w, int W)
    {
        return engine_run<INV, true>(d, A, B, tw, cp, w, W, 8, true);
    }
};

// ------------------------------------------------------------------------------------------------
// helpers
// ------------------------------------------------------------------------------------------------

// Value of the zero-padded, shifted PSF at flat index `flat` of the [d2][d1][d0] volume: inverse of
// the scatter in fftShiftKernel (reference :145-164) called with (k0,k1,k2,d0,d1,d2) (:454-461).
__device__ __forceinline__ float psf_tap(const PsfGather& g, long long flat)
{
    int cq, bq, aq;
    if (flat < 0x7fffffffLL) {
        unsigned f = (unsigned)flat;
        unsigned t = f / (unsigned)g.d2;
        cq = (int)(f - t * (unsigned)g.d2);
        unsigned a = t / (unsigned)g.d1;
        bq = (int)(t - a * (unsigned)g.d1);
        aq = (int)a;
    } else {
        long long t = flat / g.d2;
        cq = (int)(flat - t * g.d2);
        long long a = t / g.d1;
        bq = (int)(t - a * g.d1);
        aq = (int)a;
    }
    const int h0 = g.k0 / 2, h1 = g.k1 / 2, h2 = g.k2 / 2;
    int a, b, c;
    if (aq < g.k0 - h0) a = aq + h0;
    else if (aq >= g.d0 - h0) a = aq - g.d0 + h0;
    else return 0.f;
    if (bq < g.k1 - h1) b = bq + h1;
    else if (bq >= g.d1 - h1) b = bq - g.d1 + h1;
    else return 0.f;
    if (cq < g.k2 - h2) c = cq + h2;
    else if (cq >= g.d2 - h2) c = cq - g.d2 + h2;
    else return 0.f;
    return __ldg(g.kernel + (c + g.k2 * (b + g.k1 * a)));
}

// numpy.pad(mode="reflect") index: even about 0, period 2(n-1)
__device__ __forceinline__ int pad_fold(int s, int n)
{
    if (n == 1) return 0;
    const int T = 2 * (n - 1);
    s = s < 0 ? -s : s;
    if (s >= T) s %= T;
    return s < n ? s : T - s;
}
// source row feeding padded row (z, y); nullptr = the row is all zero
__device__ __forceinline__ const float* pad_src_row(const PadGeom& pg, const float* src, int z, int y)
{
    int sz = z - pg.oz, sy = y - pg.oy;
    if (pg.mode == 1) {
        sz = pad_fold(sz, pg.sz);
        sy = pad_fold(sy, pg.sy);
    } else if (sz < 0 || sz >= pg.sz || sy < 0 || sy >= pg.sy) {
        return nullptr;
    }
    return src + ((size_t)sz * pg.sy + sy) * pg.sx;
}
// source element feeding padded column x of that row; nullptr = zero
__device__ __forceinline__ const float* pad_src_elem(const PadGeom& pg, const float* row, int x)
{
    const int s = x - pg.ox;
    if (pg.mode == 1) return row + pad_fold(s, pg.sx);
    return (row != nullptr && s >= 0 && s < pg.sx) ? row + s : nullptr;
}

__device__ __forceinline__ void cp_async4(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async8(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all()
{
    asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;\n" ::: "memory");
}

// ------------------------------------------------------------------------------------------------
// X passes.  16 rows per CTA, two shared-memory tiles:
//   row tile    [16 rows][P float2], P odd: global rows as interleaved float2 (real rows: two
//               consecutive samples; spectrum rows: (re, im) per bin).  Filled / drained with lanes
//               running along x (coalesced, conflict-free); read / written column-wise with lanes
//               running along row pairs (conflict-free because P is odd);
//   engine tile [L positions][8 row pairs] float4, pair-planar: the layout of fft_engine.cuh.
// Moving data between the two tiles is the transposition; kx comes out in natural order.  Spectrum
// rows in GLOBAL memory are pair-planar (fc_common.h); the conversion from / to interleaved is one
// lane shuffle in the coalesced copy loops.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int x_row_pitch(int L, int xcp) { return (L > xcp ? L : xcp) | 1; }

constexpr int XLB = 10;   // spectrum-row loads in flight per thread in the inverse x pass

struct XSmem {
    float4* A;
    float4* B;
    float4* tw;
    float4* rtw;    // roots of the Rader n-point transform (run-time plans with a Rader stage)
    float2* rowt;
    int P;
};

__device__ __forceinline__ XSmem x_carve(float4* smem, const Geometry& g, const AxisPlanDev& pl, int L, int txp)
{
    XSmem s;
    s.P = x_row_pitch(L, g.xcp);
    s.A = smem;
    s.B = pl.generic ? (s.A + (size_t)L * txp) : nullptr;
    s.tw = s.A + (size_t)L * txp * (pl.generic ? 2 : 1);
    s.rtw = s.tw + L;
    s.rowt = reinterpret_cast<float2*>(s.rtw + pl.rader_n);
    return s;
}

// ---- static-plan stages that touch the row tile directly ---------------------------------------------
// forward first stage: row tile (rows 2cp, 2cp+1) -> registers -> engine tile
template <int R, int L, int NW>
__device__ __forceinline__ void sstage_rows_first(const float2* __restrict__ rowt, int P, float4* __restrict__ sm,
                                                  const float4* __restrict__ tw, int cp, int w)
{
    constexpr int S = L / R;
    constexpr int ITER = (S + NW - 1) / NW;
    const float2* ra = rowt + (2 * cp) * P;
    const float2* rb = ra + P;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int j = w + it * NW;
        if ((S % NW) != 0 && j >= S) break;
        p2 r[R], i[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            const float2 u = ra[j + k * S], v = rb[j + k * S];
            r[k] = make_float2(u.x, v.x);
            i[k] = make_float2(u.y, v.y);
        }
        Dft<R>::run(r, i);
#pragma unroll
        for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * m]);
        store_pairs<R>(sm, j * 8 + cp, S * 8, r, i);
    }
}

// inverse last stage: engine tile -> registers -> row tile (natural order)
template <int R, int L, int NW>
__device__ __forceinline__ void sstage_rows_last_inv(const float4* __restrict__ sm, float2* __restrict__ rowt, int P,
                                                     const float4* __restrict__ tw, int cp, int w)
{
    constexpr int S = L / R;
    constexpr int ITER = (S + NW - 1) / NW;
    float2* ra = rowt + (2 * cp) * P;
    float2* rb = ra + P;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int j = w + it * NW;
        if ((S % NW) != 0 && j >= S) break;
        p2 r[R], i[R];
        load_pairs<R>(sm, j * 8 + cp, S * 8, r, i);
#pragma unroll
        for (int k = 1; k < R; ++k) cmulc(r[k], i[k], tw[j * k]);
        Dft<R>::run(i, r);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            ra[j + m * S] = make_float2(r[m].x, i[m].x);
            rb[j + m * S] = make_float2(r[m].y, i[m].y);
        }
    }
}

template <class PL>
struct IsStaticPlan {
    static constexpr bool value = true;
};
template <>
struct IsStaticPlan<DynPlan> {
    static constexpr bool value = false;
};
template <>
struct IsStaticPlan<DynPlanBig> {
    static constexpr bool value = false;
};

// P = DynPlan: run-time radices (any length);  P = SPlan<...>: compile-time specialised stages.
template <int LOADER, class PL, int THREADS>  // LOADER 0: dense real rows, 1: PSF gather
__global__ void __launch_bounds__(THREADS, (IsBigPlan<PL>::value ? 1 : (THREADS <= 256 ? 3 : 1))) x_fwd_kernel(XArgs a)
{
    extern __shared__ float4 smem[];
    const Geometry g = a.g;
    const int L = PlanLen<PL>::get(a.P);  // complex transform length: nx/2 (even nx) or nx (odd nx)
    // rows per CTA = 2 * TXP: 16 for the static plans, fewer for very long rows (run-time plans)
    const int TXP = IsStaticPlan<PL>::value ? 8 : a.txp;
    const int NROWS = 2 * TXP;
    const XSmem sm = x_carve(smem, g, a.P, L, TXP);
    const int P = sm.P;
    float2* rowt = sm.rowt;
    constexpr int NW = THREADS / 8;

    const int t = threadIdx.x;
    const int cp = t % TXP, w = t / TXP, W = blockDim.x / TXP;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;

    pdl_launch_dependents();
    load_twiddles(sm.tw, a.P.tw, L);
    if constexpr (IsBigPlan<PL>::value) load_rader_twiddles(sm.rtw, a.P);
    pdl_wait();

    // grid-stride over tiles of NROWS rows (persistent launch: the twiddle table is loaded once per CTA)
    const long long ntile = (a.nrows + NROWS - 1) / NROWS;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {
    const long long row0 = tile * NROWS;
    // ---- global rows -> row tile (one warp per row at a time, lanes along x)
    int tile_has_data = !(LOADER == 0 && a.padOn);   // fused padding: a tile whose rows all lie in the zero border is not transformed
    for (int lrow = warp; lrow < NROWS; lrow += nwarps) {
        const long long li = row0 + lrow;
        const long long grow = (li < a.nrows) ? (a.rowList ? (long long)a.rowList[li] : li) : -1;
        float2* dst = rowt + lrow * P;
        if (grow < 0) {
            for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(0.f, 0.f);
        } else if (LOADER == 0 && a.padOn) {
            // padded row assembled from the caller's unpadded volume (lanes along x: coalesced source reads)
            const float* srow = pad_src_row(a.pad, a.in_real, a.padZ0 + (int)(grow / g.ny), (int)(grow % g.ny));
            // (4-byte cp.async: the source offset x - ox has no alignment; every load of the row is in flight at once)
            float* dstf = reinterpret_cast<float*>(dst);
            if (srow != nullptr) tile_has_data = 1;
            if (!g.odd && a.pad.mode == 0) {
                // zero padding: the interior [ox, ox + sx) is one contiguous piece of the source row -- 8-byte copies when
                // both ends are even and the source row is 8-byte aligned -- between two zero borders (per-element
                // addressing cost the fused pass 0.08 ms on the padded config 3)
                const int ox = a.pad.ox, sx = a.pad.sx;
                if (srow == nullptr) {
                    for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(0.f, 0.f);
                } else {
                    for (int x = lane; x < ox; x += 32) dstf[x] = 0.f;
                    for (int x = ox + sx + lane; x < g.nx; x += 32) dstf[x] = 0.f;
                    if ((((ox | sx) & 1) == 0) && ((reinterpret_cast<size_t>(srow) & 7) == 0)) {
                        for (int s2 = 2 * lane; s2 < sx; s2 += 64) cp_async8(dstf + ox + s2, srow + s2);
                    } else {
                        for (int s1 = lane; s1 < sx; s1 += 32) cp_async4(dstf + ox + s1, srow + s1);
                    }
                }
            } else if (!g.odd) {
                for (int x = lane; x < g.nx; x += 32) {
                    const float* e = pad_src_elem(a.pad, srow, x);
                    if (e) cp_async4(dstf + x, e);
                    else dstf[x] = 0.f;
                }
            } else {
                for (int pos = lane; pos < L; pos += 32) {
                    const float* e = pad_src_elem(a.pad, srow, pos);
                    if (e) cp_async4(dstf + 2 * pos, e);
                    else dstf[2 * pos] = 0.f;
                    dstf[2 * pos + 1] = 0.f;
                }
            }
        } else if (LOADER == 0 && !g.odd) {
            const float2* src = reinterpret_cast<const float2*>(a.in_real + grow * g.nx);
            for (int pos = lane; pos < L; pos += 32) cp_async8(dst + pos, src + pos);
        } else if (LOADER == 0) {
            const float* src = a.in_real + grow * g.nx;
            for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(__ldg(src + pos), 0.f);
        } else if (a.tapStart != nullptr) {
            // PSF row from its tap list: clear, then scatter the few taps that land in this row
            for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(0.f, 0.f);
            __syncwarp();
            float* dstf = reinterpret_cast<float*>(dst);
            const int t0 = __ldg(a.tapStart + li), t1 = __ldg(a.tapStart + li + 1);
            for (int q = t0 + lane; q < t1; q += 32) {
                const int xq = __ldg(a.tapX + q);
                const float v = __ldg(a.psf.kernel + __ldg(a.tapIdx + q));
                if (g.odd) dstf[2 * xq] = v;       // (x, 0) per complex slot
                else dstf[xq] = v;                 // two consecutive samples per complex slot
            }
        } else if (!g.odd) {
            const long long f0 = grow * g.nx;
            for (int pos = lane; pos < L; pos += 32)
                dst[pos] = make_float2(psf_tap(a.psf, f0 + 2 * pos), psf_tap(a.psf, f0 + 2 * pos + 1));
        } else {
            const long long f0 = grow * g.nx;
            for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(psf_tap(a.psf, f0 + pos), 0.f);
        }
    }
    cp_async_wait_all();
    if (!__syncthreads_or(tile_has_data)) {
        // every row of the tile is zero: so is its spectrum
        for (int lrow = warp; lrow < NROWS; lrow += nwarps) {
            const long long li = row0 + lrow;
            if (li >= a.nrows) continue;
            float2* dst = a.spec + li * g.xcp;
            for (int k = lane; k < g.xcp; k += 32) dst[k] = make_float2(0.f, 0.f);
        }
        continue;
    }

    float4* cur;
    if constexpr (IsStaticPlan<PL>::value) {
        // first stage straight from the row tile (this is the transposition), the rest on the engine tile
        sstage_rows_first<PL::R0, PL::L, NW>(rowt, P, sm.A, sm.tw, cp, w);
        __syncthreads();
        sstage<PL::R1, PL::L, PL::L / PL::R0, NW, false>(sm.A, sm.tw, cp, w);
        __syncthreads();
        if constexpr (PL::ns >= 3) {
            sstage<PL::R2, PL::L, PL::L / (PL::R0 * PL::R1), NW, false>(sm.A, sm.tw, cp, w);
            __syncthreads();
        }
        if constexpr (PL::ns >= 4) {
            sstage<PL::R3, PL::L, PL::L / (PL::R0 * PL::R1 * PL::R2), NW, false>(sm.A, sm.tw, cp, w);
            __syncthreads();
        }
        cur = sm.A;
    } else {
        for (int pos = w; pos < L; pos += W) {
            const float2 u = rowt[(2 * cp) * P + pos], v = rowt[(2 * cp + 1) * P + pos];
            sm.A[pos * TXP + cp] = make_float4(u.x, v.x, u.y, v.y);
        }
        __syncthreads();
        cur = engine_run<false, IsBigPlan<PL>::value>(a.P, sm.A, sm.B, sm.tw, cp, w, W, TXP, true, sm.rtw);
    }

    // ---- engine tile -> row tile (interleaved spectrum rows, natural kx); even nx: split the packed
    //      half-length transform into the spectrum of the real rows
    {
        float2* ra = rowt + (2 * cp) * P;
        float2* rb = ra + P;
        auto put = [&](int k, p2 re, p2 im) {
            ra[k] = make_float2(re.x, im.x);
            rb[k] = make_float2(re.y, im.y);
        };
        if (g.odd) {
            for (int k = w; k < g.xc; k += W) {
                const float4 v = cur[__ldg(a.P.pos + k) * TXP + cp];
                put(k, make_float2(v.x, v.y), make_float2(v.z, v.w));
            }
        } else {
            const int M = g.M;
            for (int k = w; k <= M / 2; k += W) {
                if (k == 0) {
                    const float4 v = cur[__ldg(a.P.pos) * TXP + cp];
                    const p2 r = make_float2(v.x, v.y), i = make_float2(v.z, v.w), z = make_float2(0.f, 0.f);
                    put(0, padd(r, i), z);
                    put(M, psub(r, i), z);
                } else {
                    const int k2 = M - k;
                    const float4 va = cur[__ldg(a.P.pos + k) * TXP + cp];
                    const float4 vb = cur[__ldg(a.P.pos + k2) * TXP + cp];
                    const float2 tk = __ldg(a.twx + k);  // exp(-2*pi*i*k/nx)
                    const p2 ar = make_float2(va.x, va.y), ai = make_float2(va.z, va.w);
                    const p2 br = make_float2(vb.x, vb.y), bi = make_float2(vb.z, vb.w);
                    const p2 er = pmuls(padd(ar, br), 0.5f), ei = pmuls(psub(ai, bi), 0.5f);
                    const p2 orr = pmuls(padd(ai, bi), 0.5f), oi = pmuls(psub(ar, br), -0.5f);
                    const p2 wr = pfmas(orr, tk.x, pmuls(oi, -tk.y));   // c*or - s*oi
                    const p2 wi = pfmas(oi, tk.x, pmuls(orr, tk.y));    // c*oi + s*or
                    put(k, padd(er, wr), padd(ei, wi));
                    put(k2, psub(er, wr), psub(wi, ei));
                }
            }
        }
        // pad columns [xc, xcp) are kept at zero so that the strided passes never see garbage
        for (int k = g.xc + w; k < g.xcp; k += W) put(k, make_float2(0.f, 0.f), make_float2(0.f, 0.f));
    }
    __syncthreads();

    // ---- row tile -> spectrum rows; interleaved -> pair-planar with one lane exchange:
    //      float2 slot k of a row holds (re_k, re_k+1) for even k and (im_k-1, im_k) for odd k
    for (int lrow = warp; lrow < NROWS; lrow += nwarps) {
        const long long li = row0 + lrow;
        const long long grow = (li < a.nrows) ? (a.rowList ? (long long)a.rowList[li] : li) : -1;
        if (grow < 0) continue;   // warp-uniform
        const float2* src = rowt + lrow * P;
        float2* dst = a.spec + (a.compactOut ? li : grow) * g.xcp;
        for (int k0 = 0; k0 < g.xcp; k0 += 32) {
            const int k = k0 + lane;
            const float2 v = (k < g.xcp) ? src[k] : make_float2(0.f, 0.f);
            const float ox = __shfl_xor_sync(0xffffffffu, v.x, 1);
            const float oy = __shfl_xor_sync(0xffffffffu, v.y, 1);
            if (k < g.xcp) dst[k] = (lane & 1) ? make_float2(oy, v.y) : make_float2(v.x, ox);
        }
    }
    __syncthreads();   // the tiles are free for the next rows
    }
}

// ------------------------------------------------------------------------------------------------
// X inverse: C2R along x for 16 rows per CTA (unnormalised, like cufftExecC2R)
// ------------------------------------------------------------------------------------------------
template <class PL, int THREADS>
__global__ void __launch_bounds__(THREADS) x_inv_kernel(XArgs a)
{
    extern __shared__ float4 smem[];
    const Geometry g = a.g;
    const int L = PlanLen<PL>::get(a.P);
    // rows per CTA = 2 * TXP: 16 for the static plans, fewer for very long rows (run-time plans)
    const int TXP = IsStaticPlan<PL>::value ? 8 : a.txp;
    const int NROWS = 2 * TXP;
    const XSmem sm = x_carve(smem, g, a.P, L, TXP);
    const int P = sm.P;
    float2* rowt = sm.rowt;
    constexpr int NW = THREADS / 8;

    const int t = threadIdx.x;
    const int cp = t % TXP, w = t / TXP, W = blockDim.x / TXP;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;

    pdl_launch_dependents();
    load_twiddles(sm.tw, a.P.tw, L);
    if constexpr (IsBigPlan<PL>::value) load_rader_twiddles(sm.rtw, a.P);
    pdl_wait();

    const long long ntile = (a.nrows + NROWS - 1) / NROWS;
    for (long long tile = blockIdx.x; tile < ntile; tile += gridDim.x) {   // persistent, see x_fwd_kernel
    const long long row0 = tile * NROWS;
    if (a.padOn) {   // fused crop: a tile without a single interior row is not transformed (nothing of it is stored)
        int mine = 0;
        if (t < NROWS && row0 + t < a.nrows) {
            const long long grow = row0 + t;
            const int sz = a.padZ0 + (int)(grow / g.ny) - a.pad.oz, sy = (int)(grow % g.ny) - a.pad.oy;
            mine = sz >= 0 && sz < a.pad.sz && sy >= 0 && sy < a.pad.sy;
        }
        if (!__syncthreads_or(mine)) continue;
    }
    // ---- spectrum rows (pair-planar) -> row tile (interleaved); loads are issued XLB at a time
    for (int lrow = warp; lrow < NROWS; lrow += nwarps) {
        const long long grow = row0 + lrow;
        float2* dst = rowt + lrow * P;
        const bool have = grow < a.nrows;   // warp-uniform
        const float2* src = a.spec + (have ? grow : 0) * g.xcp;
        for (int k0 = 0; k0 < g.xcp; k0 += 32 * XLB) {
            float2 v[XLB];
#pragma unroll
            for (int u = 0; u < XLB; ++u) {
                const int k = k0 + u * 32 + lane;
                v[u] = (have && k < g.xcp) ? src[k] : make_float2(0.f, 0.f);
            }
#pragma unroll
            for (int u = 0; u < XLB; ++u) {
                const int k = k0 + u * 32 + lane;
                const float ox = __shfl_xor_sync(0xffffffffu, v[u].x, 1);
                const float oy = __shfl_xor_sync(0xffffffffu, v[u].y, 1);
                if (k < g.xcp) dst[k] = (lane & 1) ? make_float2(oy, v[u].y) : make_float2(v[u].x, ox);
            }
        }
    }
    __syncthreads();

    // ---- row tile -> engine tile (positions), merging the half spectrum into the packed transform
    {
        const float2* ra = rowt + (2 * cp) * P;
        const float2* rb = ra + P;
        auto get = [&](int k, p2& re, p2& im) {
            const float2 u = ra[k], v = rb[k];
            re = make_float2(u.x, v.x);
            im = make_float2(u.y, v.y);
        };
        if (g.odd) {
            for (int k = w; k < g.xc; k += W) {
                p2 re, im;
                get(k, re, im);
                sm.A[__ldg(a.P.pos + k) * TXP + cp] = make_float4(re.x, re.y, im.x, im.y);
                if (k > 0) sm.A[__ldg(a.P.pos + (g.nx - k)) * TXP + cp] = make_float4(re.x, re.y, -im.x, -im.y);
            }
        } else {
            const int M = g.M;
            for (int k = w; k <= M / 2; k += W) {
                if (k == 0) {
                    p2 x0r, x0i, xmr, xmi;
                    get(0, x0r, x0i);
                    get(M, xmr, xmi);
                    const p2 zr = padd(x0r, xmr), zi = psub(x0r, xmr);
                    sm.A[__ldg(a.P.pos) * TXP + cp] = make_float4(zr.x, zr.y, zi.x, zi.y);
                } else {
                    const int k2 = M - k;
                    p2 ar, ai, br, bi;
                    get(k, ar, ai);
                    get(k2, br, bi);
                    const float2 tk = __ldg(a.twx + k);
                    const p2 sr = padd(ar, br), si = psub(ai, bi);
                    const p2 Dr = psub(ar, br), Di = padd(ai, bi);
                    const p2 dr = pfmas(Dr, tk.x, pmuls(Di, tk.y));    // D * conj(w)
                    const p2 di = pfmas(Di, tk.x, pmuls(Dr, -tk.y));
                    const p2 z1r = psub(sr, di), z1i = padd(si, dr);
                    const p2 z2r = padd(sr, di), z2i = psub(dr, si);
                    sm.A[__ldg(a.P.pos + k) * TXP + cp] = make_float4(z1r.x, z1r.y, z1i.x, z1i.y);
                    if (k2 != k) sm.A[__ldg(a.P.pos + k2) * TXP + cp] = make_float4(z2r.x, z2r.y, z2i.x, z2i.y);
                }
            }
        }
    }
    __syncthreads();

    if constexpr (IsStaticPlan<PL>::value) {
        // inverse stages on the engine tile; the last one writes the row tile (the transposition)
        if constexpr (PL::ns >= 4) {
            sstage<PL::R3, PL::L, PL::R3, NW, true>(sm.A, sm.tw, cp, w);
            __syncthreads();
        }
        if constexpr (PL::ns >= 3) {
            sstage<PL::R2, PL::L, PL::R2 * PL::R3, NW, true>(sm.A, sm.tw, cp, w);
            __syncthreads();
        }
        sstage<PL::R1, PL::L, PL::R1 * PL::R2 * PL::R3, NW, true>(sm.A, sm.tw, cp, w);
        __syncthreads();
        sstage_rows_last_inv<PL::R0, PL::L, NW>(sm.A, rowt, P, sm.tw, cp, w);
    } else {
        float4* cur = engine_run<true, IsBigPlan<PL>::value>(a.P, sm.A, sm.B, sm.tw, cp, w, W, TXP, true, sm.rtw);
        for (int pos = w; pos < L; pos += W) {
            const float4 v = cur[pos * TXP + cp];
            rowt[(2 * cp) * P + pos] = make_float2(v.x, v.z);
            rowt[(2 * cp + 1) * P + pos] = make_float2(v.y, v.w);
        }
    }
    __syncthreads();

    // ---- row tile -> real rows
    for (int lrow = warp; lrow < NROWS; lrow += nwarps) {
        const long long grow = row0 + lrow;
        if (grow >= a.nrows) continue;
        const float2* src = rowt + lrow * P;
        if (a.padOn) {
            // crop: only the interior part of interior rows goes back to the caller's unpadded volume
            const PadGeom& pg = a.pad;
            const int sz = a.padZ0 + (int)(grow / g.ny) - pg.oz, sy = (int)(grow % g.ny) - pg.oy;
            if (sz < 0 || sz >= pg.sz || sy < 0 || sy >= pg.sy) continue;   // warp-uniform
            float* dst = a.out_real + ((size_t)sz * pg.sy + sy) * pg.sx;
            if (g.odd) {
                for (int x = lane; x < pg.sx; x += 32) dst[x] = src[x + pg.ox].x;
            } else {
                const float* srcf = reinterpret_cast<const float*>(src);
                for (int x = lane; x < pg.sx; x += 32) dst[x] = srcf[x + pg.ox];
            }
            continue;
        }
        if (g.odd) {
            float* dst = a.out_real + grow * g.nx;
            for (int pos = lane; pos < L; pos += 32) dst[pos] = src[pos].x;
        } else {
            float2* dst = reinterpret_cast<float2*>(a.out_real + grow * g.nx);
            for (int pos = lane; pos < L; pos += 32) dst[pos] = src[pos];
        }
    }
    __syncthreads();   // the tiles are free for the next rows
    }
}

}  // namespace fcb200
