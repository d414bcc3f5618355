// General register-resident row-wise x passes: plans (R0, R1) or (R0, R1, R2) of power-of-two radices
// with R0 the largest radix, last radix RL in {8, 16} and every earlier stage stride a multiple of RL
// (e.g. nx = 256 -> (16,8), nx = 1024 -> (8,8,8), nx = 2048 -> (16,8,8)); also (24, 8) for nx = 384 (the
// 384^3 blocks of BASELINE config 4): two-stage plans use no plan table besides the M roots, so the kernel
// does not care which radices the planner chose for the other x kernels.  Same idea as fft_xrow.cuh:
//   * a group of TG = M / R0 threads owns one pair of adjacent rows (packed fp32), lanes run along x,
//     so global loads / stores are coalesced straight from / to registers;
//   * stages are in-place DIF butterflies in registers; between stages the values make one trip through
//     a padded per-group exchange buffer (position p at p + p / RL: every access pattern below is
//     conflict-free);  groups synchronise with __syncwarp (TG <= 32) or a named barrier (TG > 32);
//   * the Hermitian split / merge reads Z[k] and Z[M-k] from a natural-order copy of the spectrum so
//     that consecutive lanes own consecutive bins and the final stores are coalesced.
#pragma once
#include "fft_engine.cuh"
#include "fft_kernels.h"

namespace fcb200 {

template <int R0, int R1, int R2>
struct XRowG {
    static constexpr int M = R0 * R1 * R2;
    static constexpr int NS = (R2 > 1) ? 3 : 2;
    static constexpr int RL = (R2 > 1) ? R2 : R1;
    static constexpr int TG = M / R0;              // threads per row pair
    static constexpr int NV = R0;                  // values held per thread
    static constexpr int PADM = M + M / RL + 16;   // exchange buffer, float4 per row pair
    // twiddle tables, m-major: stage 1 [R0][M/R0], stage 2 (3-stage plans) [R1][M/(R0*R1)]
    static constexpr int TW1 = M;
    static constexpr int TW2 = (NS == 3) ? (M / R0) : 0;
};

template <int TG>
__device__ __forceinline__ void group_sync(int grp)
{
    if (TG <= 32) __syncwarp();
    else asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "n"(TG) : "memory");
}

template <int RL>
__device__ __forceinline__ int gpad(int p) { return p + p / RL; }

// One register stage: for each of the ITER butterflies of this thread, values v[it*R + k] are the inputs
// at positions base + k*S and become the outputs at the same positions.
template <int R, int M, int Li, int TG, bool INV>
struct RStage {
    static constexpr int S = Li / R, NB = M / R, ITER = NB / TG;
    static __device__ __forceinline__ int pos(int tl, int it, int k)
    {
        const int b = tl + it * TG;
        return (b / S) * Li + (b % S) + k * S;
    }
};

template <int R0, int R1, int R2, int THREADS>
__global__ void __launch_bounds__(THREADS) xrowg_fwd_kernel(XArgs a)
{
    typedef XRowG<R0, R1, R2> G;
    constexpr int M = G::M, TG = G::TG, RL = G::RL, NS = G::NS, RP = THREADS / TG;
    extern __shared__ float4 smem[];
    float4* tw1 = smem;                       // [m][j], j < M/R0
    float4* tw2 = tw1 + G::TW1;               // [m][j], j < M/(R0*R1)   (3-stage plans)
    float4* xch = tw2 + G::TW2;

    const Geometry g = a.g;
    const int t = threadIdx.x, lane = t & 31;
    const int grp = t / TG, tl = t % TG;
    pdl_launch_dependents();
    {
        constexpr int S1 = M / R0;
        for (int idx = t; idx < M; idx += THREADS) {
            const int m = idx / S1, j = idx % S1;
            const float2 w = __ldg(a.P.tw + j * m);
            tw1[idx] = make_float4(w.x, w.x, w.y, w.y);
        }
        if constexpr (NS == 3) {
            constexpr int L2 = M / R0, S2 = L2 / R1;
            for (int idx = t; idx < L2; idx += THREADS) {
                const int m = idx / S2, j = idx % S2;
                const float2 w = __ldg(a.P.tw + j * m * R0);   // w_L2^(j m) = w_M^(j m R0)
                tw2[idx] = make_float4(w.x, w.x, w.y, w.y);
            }
        }
    }
    pdl_wait();
    __syncthreads();

    // Persistent: the CTA walks over blocks of RP row pairs, so the twiddle tables above (M + M/R0 roots, as much
    // data as half a row pair for nx = 2048) are built once per CTA instead of once per 2*RP rows.
    float4* x = xch + (size_t)grp * G::PADM;
    const long long nblk = (a.nrows + 2 * RP - 1) / (2 * RP);
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long rowA = (blk * RP + grp) * 2;
    const bool hasA = rowA < a.nrows, hasB = rowA + 1 < a.nrows;
    const float2* srcA = reinterpret_cast<const float2*>(a.in_real + rowA * g.nx);
    const float2* srcB = reinterpret_cast<const float2*>(a.in_real + (rowA + 1) * g.nx);

    p2 r[R0], i[R0];
    // ---- stage 1 (radix R0, stride S1 = TG): thread tl owns butterfly j = tl
    {
        constexpr int S1 = M / R0;
#pragma unroll
        for (int k = 0; k < R0; ++k) {
            const float2 ua = hasA ? __ldg(srcA + tl + S1 * k) : make_float2(0.f, 0.f);
            const float2 ub = hasB ? __ldg(srcB + tl + S1 * k) : make_float2(0.f, 0.f);
            r[k] = make_float2(ua.x, ub.x);
            i[k] = make_float2(ua.y, ub.y);
        }
        Dft<R0>::run(r, i);
#pragma unroll
        for (int m = 1; m < R0; ++m) cmul(r[m], i[m], tw1[m * S1 + tl]);
#pragma unroll
        for (int m = 0; m < R0; ++m) x[gpad<RL>(tl + m * S1)] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
    }
    group_sync<TG>(grp);
    // ---- stage 2 (radix R1 on blocks of L2 = M/R0)
    {
        constexpr int L2 = M / R0;
        typedef RStage<R1, M, L2, TG, false> St;
        static_assert(St::ITER * R1 == R0, "register budget");
#pragma unroll
        for (int it = 0; it < St::ITER; ++it) {
#pragma unroll
            for (int k = 0; k < R1; ++k) {
                const float4 v = x[gpad<RL>(St::pos(tl, it, k))];
                r[it * R1 + k] = make_float2(v.x, v.y);
                i[it * R1 + k] = make_float2(v.z, v.w);
            }
            Dft<R1>::run(r + it * R1, i + it * R1);
            if constexpr (NS == 3) {
                const int j = (tl + it * TG) % St::S;
#pragma unroll
                for (int m = 1; m < R1; ++m) cmul(r[it * R1 + m], i[it * R1 + m], tw2[m * St::S + j]);
            }
        }
        if constexpr (NS == 3) {
            group_sync<TG>(grp);   // all stage-2 inputs are read before anybody overwrites them
#pragma unroll
            for (int it = 0; it < St::ITER; ++it)
#pragma unroll
                for (int m = 0; m < R1; ++m)
                    x[gpad<RL>(St::pos(tl, it, m))] =
                        make_float4(r[it * R1 + m].x, r[it * R1 + m].y, i[it * R1 + m].x, i[it * R1 + m].y);
            group_sync<TG>(grp);
        }
    }
    // ---- stage 3 (3-stage plans: radix R2 on blocks of R2, stride 1)
    if constexpr (NS == 3) {
        typedef RStage<R2, M, R2, TG, false> St;
#pragma unroll
        for (int it = 0; it < St::ITER; ++it) {
#pragma unroll
            for (int k = 0; k < R2; ++k) {
                const float4 v = x[gpad<RL>(St::pos(tl, it, k))];
                r[it * R2 + k] = make_float2(v.x, v.y);
                i[it * R2 + k] = make_float2(v.z, v.w);
            }
            Dft<R2>::run(r + it * R2, i + it * R2);
        }
    }
    // ---- natural-order copy of Z: the value at position p belongs to bin rev[p]
    group_sync<TG>(grp);
    {
        constexpr int ITERL = (M / RL) / TG;
#pragma unroll
        for (int it = 0; it < ITERL; ++it) {
            const int b = tl + it * TG;
            // two stages: position b*RL + m holds bin b + m*R0, whatever radices the plan tables were built for
            const int k0 = (NS == 2) ? b : __ldg(a.P.rev + b * RL);   // bins k0 + m * (M / RL)
#pragma unroll
            for (int m = 0; m < RL; ++m) {
                const int k = k0 + m * (M / RL);
                x[k + k / R0] = make_float4(r[it * RL + m].x, r[it * RL + m].y, i[it * RL + m].x, i[it * RL + m].y);
            }
        }
    }
    group_sync<TG>(grp);

    // ---- split: X[k] = E + w^k O for the bins k = tl + TG*q of this thread; coalesced pair-planar store
    float2* dstA = a.spec + rowA * g.xcp;
    float2* dstB = a.spec + (rowA + 1) * g.xcp;
    const bool odd = lane & 1;
    p2 nyq = make_float2(0.f, 0.f);
#pragma unroll
    for (int q = 0; q < R0; ++q) {
        const int k = tl + TG * q;
        const int k2 = (M & (M - 1)) ? (k ? M - k : 0) : ((M - k) & (M - 1));
        const float4 va = x[k + k / R0], vb = x[k2 + k2 / R0];
        const p2 zr = make_float2(va.x, va.y), zi = make_float2(va.z, va.w);
        const p2 pr = make_float2(vb.x, vb.y), pi = make_float2(vb.z, vb.w);
        if (q == 0) nyq = psub(zr, zi);   // meaningful on tl == 0 only
        const float2 tk = __ldg(a.twx + k);
        const p2 er = pmuls(padd(zr, pr), 0.5f), ei = pmuls(psub(zi, pi), 0.5f);
        const p2 orr = pmuls(padd(zi, pi), 0.5f), oi = pmuls(psub(zr, pr), -0.5f);
        const p2 xr = padd(er, pfmas(orr, tk.x, pmuls(oi, -tk.y)));
        const p2 xi = padd(ei, pfmas(oi, tk.x, pmuls(orr, tk.y)));
        const p2 send = odd ? xr : xi;
        p2 got;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        const float2 oa = odd ? make_float2(got.x, xi.x) : make_float2(xr.x, got.x);
        const float2 ob = odd ? make_float2(got.y, xi.y) : make_float2(xr.y, got.y);
        if (hasA) dstA[k] = oa;
        if (hasB) dstB[k] = ob;
    }
    if (tl == 0) {
        if (hasA) {
            dstA[M] = make_float2(nyq.x, 0.f);
            for (int k = M + 1; k < g.xcp; ++k) dstA[k] = make_float2(0.f, 0.f);
        }
        if (hasB) {
            dstB[M] = make_float2(nyq.y, 0.f);
            for (int k = M + 1; k < g.xcp; ++k) dstB[k] = make_float2(0.f, 0.f);
        }
    }
    group_sync<TG>(grp);   // the exchange buffer is free for the next block of rows
    }
}

template <int R0, int R1, int R2, int THREADS>
__global__ void __launch_bounds__(THREADS) xrowg_inv_kernel(XArgs a)
{
    typedef XRowG<R0, R1, R2> G;
    constexpr int M = G::M, TG = G::TG, RL = G::RL, NS = G::NS, RP = THREADS / TG;
    extern __shared__ float4 smem[];
    float4* tw1 = smem;
    float4* tw2 = tw1 + G::TW1;
    float4* xch = tw2 + G::TW2;

    const Geometry g = a.g;
    const int t = threadIdx.x, lane = t & 31;
    const int grp = t / TG, tl = t % TG;
    pdl_launch_dependents();
    {
        constexpr int S1 = M / R0;
        for (int idx = t; idx < M; idx += THREADS) {
            const int m = idx / S1, j = idx % S1;
            const float2 w = __ldg(a.P.tw + j * m);
            tw1[idx] = make_float4(w.x, w.x, w.y, w.y);
        }
        if constexpr (NS == 3) {
            constexpr int L2 = M / R0, S2 = L2 / R1;
            for (int idx = t; idx < L2; idx += THREADS) {
                const int m = idx / S2, j = idx % S2;
                const float2 w = __ldg(a.P.tw + j * m * R0);
                tw2[idx] = make_float4(w.x, w.x, w.y, w.y);
            }
        }
    }
    pdl_wait();
    __syncthreads();

    const bool odd = lane & 1;
    float4* x = xch + (size_t)grp * G::PADM;
    const long long nblk = (a.nrows + 2 * RP - 1) / (2 * RP);
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {   // persistent, see xrowg_fwd_kernel
    const long long rowA = (blk * RP + grp) * 2;
    const bool hasA = rowA < a.nrows, hasB = rowA + 1 < a.nrows;
    const float2* srcA = a.spec + rowA * g.xcp;
    const float2* srcB = a.spec + (rowA + 1) * g.xcp;

    // ---- pair-planar rows -> natural-order (re, im) copy of X in the exchange buffer
#pragma unroll
    for (int q = 0; q < R0; ++q) {
        const int k = tl + TG * q;
        const float2 va = hasA ? srcA[k] : make_float2(0.f, 0.f);
        const float2 vb = hasB ? srcB[k] : make_float2(0.f, 0.f);
        const p2 send = odd ? make_float2(va.x, vb.x) : make_float2(va.y, vb.y);
        p2 got;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        const p2 re = odd ? got : make_float2(va.x, vb.x);
        const p2 im = odd ? make_float2(va.y, vb.y) : got;
        x[k + k / R0] = make_float4(re.x, re.y, im.x, im.y);
    }
    const p2 xm = make_float2(hasA ? srcA[M].x : 0.f, hasB ? srcB[M].x : 0.f);   // Nyquist (real)
    group_sync<TG>(grp);

    // ---- merge: Z[k] for the bins of this thread's last-stage butterflies, straight into registers
    p2 r[R0], i[R0];
    {
        constexpr int ITERL = (M / RL) / TG;
#pragma unroll
        for (int it = 0; it < ITERL; ++it) {
            const int b = tl + it * TG;
            const int k0 = (NS == 2) ? b : __ldg(a.P.rev + b * RL);
#pragma unroll
            for (int m = 0; m < RL; ++m) {
                const int k = k0 + m * (M / RL);
                const int k2 = (M & (M - 1)) ? (k ? M - k : 0) : ((M - k) & (M - 1));
                const float4 va = x[k + k / R0], vb = x[k2 + k2 / R0];
                const p2 ar = make_float2(va.x, va.y), ai = make_float2(va.z, va.w);
                const p2 br = make_float2(vb.x, vb.y), bi = make_float2(vb.z, vb.w);
                const float2 tk = __ldg(a.twx + k);
                const p2 sr = padd(ar, br), si = psub(ai, bi);
                const p2 Dr = psub(ar, br), Di = padd(ai, bi);
                const p2 dr = pfmas(Dr, tk.x, pmuls(Di, tk.y));
                const p2 di = pfmas(Di, tk.x, pmuls(Dr, -tk.y));
                p2 zr = psub(sr, di), zi = padd(si, dr);
                if (k == 0) {   // pairs with the Nyquist bin, not with itself
                    zr = padd(ar, xm);
                    zi = psub(ar, xm);
                }
                r[it * RL + m] = zr;
                i[it * RL + m] = zi;
            }
            // first inverse stage: radix RL on the block, stride 1, no twiddles
            Dft<RL>::run(i + it * RL, r + it * RL);
        }
        group_sync<TG>(grp);   // the natural-order copy has been consumed
#pragma unroll
        for (int it = 0; it < ITERL; ++it)
#pragma unroll
            for (int m = 0; m < RL; ++m) {
                const int p = (tl + it * TG) * RL + m;
                x[gpad<RL>(p)] = make_float4(r[it * RL + m].x, r[it * RL + m].y, i[it * RL + m].x, i[it * RL + m].y);
            }
    }
    group_sync<TG>(grp);
    // ---- middle inverse stage (3-stage plans): radix R1 on blocks of L2 = M/R0, conj twiddles first
    if constexpr (NS == 3) {
        constexpr int L2 = M / R0;
        typedef RStage<R1, M, L2, TG, true> St;
#pragma unroll
        for (int it = 0; it < St::ITER; ++it) {
            const int j = (tl + it * TG) % St::S;
#pragma unroll
            for (int k = 0; k < R1; ++k) {
                const float4 v = x[gpad<RL>(St::pos(tl, it, k))];
                r[it * R1 + k] = make_float2(v.x, v.y);
                i[it * R1 + k] = make_float2(v.z, v.w);
            }
#pragma unroll
            for (int k = 1; k < R1; ++k) cmulc(r[it * R1 + k], i[it * R1 + k], tw2[k * St::S + j]);
            Dft<R1>::run(i + it * R1, r + it * R1);
        }
        group_sync<TG>(grp);
#pragma unroll
        for (int it = 0; it < St::ITER; ++it)
#pragma unroll
            for (int m = 0; m < R1; ++m)
                x[gpad<RL>(St::pos(tl, it, m))] =
                    make_float4(r[it * R1 + m].x, r[it * R1 + m].y, i[it * R1 + m].x, i[it * R1 + m].y);
        group_sync<TG>(grp);
    }
    // ---- last inverse stage (radix R0, stride S1): outputs z[tl + m*S1] in natural order
    {
        constexpr int S1 = M / R0;
#pragma unroll
        for (int k = 0; k < R0; ++k) {
            const float4 v = x[gpad<RL>(tl + k * S1)];
            r[k] = make_float2(v.x, v.y);
            i[k] = make_float2(v.z, v.w);
        }
#pragma unroll
        for (int k = 1; k < R0; ++k) cmulc(r[k], i[k], tw1[k * S1 + tl]);
        Dft<R0>::run(i, r);
        float2* dstA = reinterpret_cast<float2*>(a.out_real + rowA * g.nx);
        float2* dstB = reinterpret_cast<float2*>(a.out_real + (rowA + 1) * g.nx);
#pragma unroll
        for (int m = 0; m < R0; ++m) {
            if (hasA) dstA[tl + m * S1] = make_float2(r[m].x, i[m].x);
            if (hasB) dstB[tl + m * S1] = make_float2(r[m].y, i[m].y);
        }
    }
    group_sync<TG>(grp);   // the exchange buffer is free for the next block of rows
    }
}

}  // namespace fcb200
