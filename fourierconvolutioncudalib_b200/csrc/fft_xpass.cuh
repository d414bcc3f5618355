// X-pass kernels (R2C / C2R along the fastest axis), shared by the run-time-radix build
// (fft_kernels.cu) and the compile-time specialised build (fft_static.cu).
#pragma once
#include "fft_static.cuh"
#include "fft_kernels.h"

namespace fcb200 {

struct DynPlan {};

template <class P>
struct PlanLen {
    static __device__ __forceinline__ int get(const AxisPlanDev&) { return P::L; }
};
template <>
struct PlanLen<DynPlan> {
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
//   row tile    [16 rows][P float2], P odd: holds global rows verbatim (real rows as float2 = two
//               consecutive samples; spectrum rows in pair-planar form).  Filled / drained with lanes
//               running along x (coalesced, conflict-free); read / written column-wise with lanes
//               running along row pairs (conflict-free because P is odd);
//   engine tile [L positions][8 row pairs] float4, pair-planar: the layout of fft_engine.cuh.
// Moving data between the two tiles is the transposition; kx comes out in natural order.
// ------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int x_row_pitch(int L, int xcp) { return (L > xcp ? L : xcp) | 1; }

// float index of the real part of column k inside a pair-planar row (imaginary part at +2)
__device__ __forceinline__ int pp(int k) { return ((k >> 1) << 2) | (k & 1); }

struct XSmem {
    float4* A;
    float4* B;
    float4* tw;
    float2* rowt;
    int P;
};

__device__ __forceinline__ XSmem x_carve(float4* smem, const Geometry& g, const AxisPlanDev& pl)
{
    XSmem s;
    s.P = x_row_pitch(pl.L, g.xcp);
    s.A = smem;
    s.B = pl.generic ? (s.A + (size_t)pl.L * 8) : nullptr;
    s.tw = s.A + (size_t)pl.L * 8 * (pl.generic ? 2 : 1);
    s.rowt = reinterpret_cast<float2*>(s.tw + pl.L);
    return s;
}

// P = DynPlan: run-time radices (any length);  P = SPlan<...>: compile-time specialised stages.
template <int LOADER, class PL, int THREADS>  // LOADER 0: dense real rows, 1: PSF gather
__global__ void __launch_bounds__(THREADS) x_fwd_kernel(XArgs a)
{
    extern __shared__ float4 smem[];
    const Geometry g = a.g;
    const int L = PlanLen<PL>::get(a.P);  // complex transform length: nx/2 (even nx) or nx (odd nx)
    const XSmem sm = x_carve(smem, g, a.P);
    const int P = sm.P;
    float2* rowt = sm.rowt;

    const int t = threadIdx.x;
    const int cp = t & 7, w = t >> 3, W = blockDim.x >> 3;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    const long long row0 = (long long)blockIdx.x * 16;

    load_twiddles(sm.tw, a.P.tw, L);

    // ---- global rows -> row tile (one warp per row at a time, lanes along x)
    for (int lrow = warp; lrow < 16; lrow += nwarps) {
        const long long li = row0 + lrow;
        const long long grow = (li < a.nrows) ? (a.rowList ? (long long)a.rowList[li] : li) : -1;
        float2* dst = rowt + lrow * P;
        if (grow < 0) {
            for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(0.f, 0.f);
        } else if (LOADER == 0 && !g.odd) {
            const float2* src = reinterpret_cast<const float2*>(a.in_real + grow * g.nx);
            for (int pos = lane; pos < L; pos += 32) cp_async8(dst + pos, src + pos);
        } else if (LOADER == 0) {
            const float* src = a.in_real + grow * g.nx;
            for (int pos = lane; pos < L; pos += 32) dst[pos] = make_float2(__ldg(src + pos), 0.f);
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
    __syncthreads();

    // ---- row tile -> engine tile (lanes along row pairs: the transposition)
    for (int pos = w; pos < L; pos += W) {
        const float2 u = rowt[(2 * cp) * P + pos], v = rowt[(2 * cp + 1) * P + pos];
        sm.A[pos * 8 + cp] = make_float4(u.x, v.x, u.y, v.y);
    }
    __syncthreads();

    float4* cur = PlanRun<PL, THREADS / 8, false>::run(a.P, sm.A, sm.B, sm.tw, cp, w, W);

    // ---- engine tile -> row tile (pair-planar spectrum rows, natural kx); even nx: split the
    //      packed half-length transform into the spectrum of the real rows
    {
        float* ra = reinterpret_cast<float*>(rowt + (2 * cp) * P);
        float* rb = reinterpret_cast<float*>(rowt + (2 * cp + 1) * P);
        auto put = [&](int k, p2 re, p2 im) {
            const int f = pp(k);
            ra[f] = re.x;
            ra[f + 2] = im.x;
            rb[f] = re.y;
            rb[f + 2] = im.y;
        };
        if (g.odd) {
            for (int k = w; k < g.xc; k += W) {
                const float4 v = cur[__ldg(a.P.pos + k) * 8 + cp];
                put(k, make_float2(v.x, v.y), make_float2(v.z, v.w));
            }
        } else {
            const int M = g.M;
            for (int k = w; k <= M / 2; k += W) {
                if (k == 0) {
                    const float4 v = cur[__ldg(a.P.pos) * 8 + cp];
                    const p2 r = make_float2(v.x, v.y), i = make_float2(v.z, v.w), z = make_float2(0.f, 0.f);
                    put(0, padd(r, i), z);
                    put(M, psub(r, i), z);
                } else {
                    const int k2 = M - k;
                    const float4 va = cur[__ldg(a.P.pos + k) * 8 + cp];
                    const float4 vb = cur[__ldg(a.P.pos + k2) * 8 + cp];
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

    // ---- row tile -> spectrum rows (verbatim)
    for (int lrow = warp; lrow < 16; lrow += nwarps) {
        const long long li = row0 + lrow;
        const long long grow = (li < a.nrows) ? (a.rowList ? (long long)a.rowList[li] : li) : -1;
        if (grow < 0) continue;
        const float2* src = rowt + lrow * P;
        float2* dst = a.spec + grow * g.xcp;
        for (int k = lane; k < g.xcp; k += 32) dst[k] = src[k];
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
    const XSmem sm = x_carve(smem, g, a.P);
    const int P = sm.P;
    float2* rowt = sm.rowt;

    const int t = threadIdx.x;
    const int cp = t & 7, w = t >> 3, W = blockDim.x >> 3;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    const long long row0 = (long long)blockIdx.x * 16;

    load_twiddles(sm.tw, a.P.tw, L);

    // ---- spectrum rows -> row tile (verbatim, asynchronous)
    for (int lrow = warp; lrow < 16; lrow += nwarps) {
        const long long grow = row0 + lrow;
        float2* dst = rowt + lrow * P;
        if (grow < a.nrows) {
            const float2* src = a.spec + grow * g.xcp;
            for (int k = lane; k < g.xcp; k += 32) cp_async8(dst + k, src + k);
        } else {
            for (int k = lane; k < g.xcp; k += 32) dst[k] = make_float2(0.f, 0.f);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    // ---- row tile -> engine tile (positions), merging the half spectrum into the packed transform
    {
        const float* ra = reinterpret_cast<const float*>(rowt + (2 * cp) * P);
        const float* rb = reinterpret_cast<const float*>(rowt + (2 * cp + 1) * P);
        auto get = [&](int k, p2& re, p2& im) {
            const int f = pp(k);
            re = make_float2(ra[f], rb[f]);
            im = make_float2(ra[f + 2], rb[f + 2]);
        };
        if (g.odd) {
            for (int k = w; k < g.xc; k += W) {
                p2 re, im;
                get(k, re, im);
                sm.A[__ldg(a.P.pos + k) * 8 + cp] = make_float4(re.x, re.y, im.x, im.y);
                if (k > 0) sm.A[__ldg(a.P.pos + (g.nx - k)) * 8 + cp] = make_float4(re.x, re.y, -im.x, -im.y);
            }
        } else {
            const int M = g.M;
            for (int k = w; k <= M / 2; k += W) {
                if (k == 0) {
                    p2 x0r, x0i, xmr, xmi;
                    get(0, x0r, x0i);
                    get(M, xmr, xmi);
                    const p2 zr = padd(x0r, xmr), zi = psub(x0r, xmr);
                    sm.A[__ldg(a.P.pos) * 8 + cp] = make_float4(zr.x, zr.y, zi.x, zi.y);
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
                    sm.A[__ldg(a.P.pos + k) * 8 + cp] = make_float4(z1r.x, z1r.y, z1i.x, z1i.y);
                    if (k2 != k) sm.A[__ldg(a.P.pos + k2) * 8 + cp] = make_float4(z2r.x, z2r.y, z2i.x, z2i.y);
                }
            }
        }
    }
    __syncthreads();

    float4* cur = PlanRun<PL, THREADS / 8, true>::run(a.P, sm.A, sm.B, sm.tw, cp, w, W);

    // ---- engine tile -> row tile (real rows, natural order), then row tile -> global
    for (int pos = w; pos < L; pos += W) {
        const float4 v = cur[pos * 8 + cp];
        rowt[(2 * cp) * P + pos] = make_float2(v.x, v.z);
        rowt[(2 * cp + 1) * P + pos] = make_float2(v.y, v.w);
    }
    __syncthreads();
    for (int lrow = warp; lrow < 16; lrow += nwarps) {
        const long long grow = row0 + lrow;
        if (grow >= a.nrows) continue;
        const float2* src = rowt + lrow * P;
        if (g.odd) {
            float* dst = a.out_real + grow * g.nx;
            for (int pos = lane; pos < L; pos += 32) dst[pos] = src[pos].x;
        } else {
            float2* dst = reinterpret_cast<float2*>(a.out_real + grow * g.nx);
            for (int pos = lane; pos < L; pos += 32) dst[pos] = src[pos];
        }
    }
}

}  // namespace fcb200
