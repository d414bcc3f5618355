// Compile-time specialised FFT stages for sm_100a: transform length, radix sequence and worker
// count are template parameters, so every stride, twiddle index and loop bound folds to an immediate
// and the issued instruction stream is almost only packed fp32 math, 128-bit shared-memory accesses
// and 128-bit global accesses.  The run-time-radix engine (fft_engine.cuh) remains the any-length
// fallback; the host picks a static kernel when one is instantiated for the plan (fft_static.cu).
//
// Same algorithm as fft_engine.cuh: in-place DIF forward (digit-reversed positions), mirrored DIT
// inverse, tile = [L positions][8 column pairs] float4 in pair-planar form.
#pragma once
#include "fft_engine.cuh"

namespace fcb200 {

// Static plan: up to four stages, trailing unused radices are 1.
template <int L_, int R0_, int R1_, int R2_ = 1, int R3_ = 1>
struct SPlan {
    static constexpr int L = L_;
    static constexpr int R0 = R0_, R1 = R1_, R2 = R2_, R3 = R3_;
    static constexpr int ns = (R3_ > 1) ? 4 : ((R2_ > 1) ? 3 : 2);
    static constexpr int RL = (ns == 4) ? R3_ : ((ns == 3) ? R2_ : R1_);   // last radix
    static_assert(R0_ * R1_ * R2_ * R3_ == L_, "radices must multiply to L");
    static_assert(R0_ > 1 && R1_ > 1, "at least two stages");
};

// digit reversal of position p (mixed radix), compile-time radices: frequency held at p
template <class P>
__host__ __device__ constexpr int srev(int p)
{
    int k = 0, mul = 1, rem = p, Li = P::L;
    const int rad[4] = {P::R0, P::R1, P::R2, P::R3};
    for (int s = 0; s < P::ns; ++s) {
        const int S = Li / rad[s];
        const int m = rem / S;
        rem -= m * S;
        k += m * mul;
        mul *= rad[s];
        Li = S;
    }
    return k;
}

// ---- shared -> shared stage ---------------------------------------------------------------------
// SWAPIN: exchange re and im of the inputs (with the same exchange on the outputs of the last stage, a
// forward transform becomes the unnormalised inverse: IDFT(x) = swap(DFT(swap(x))))
template <int R, int L, int Li, int NW, bool INV, int TXP = 8, bool SWAPIN = false>
__device__ __forceinline__ void sstage(float4* __restrict__ buf, const float4* __restrict__ tw, int cp, int w)
{
    constexpr int S = Li / R, nb = L / R, tstep = L / Li;
    constexpr int ITER = (nb + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        const int beta = b / S, j = b % S;     // S is a constant: shifts / multiply-high
        const int idx0 = (beta * Li + j) * TXP + cp;
        p2 r[R], i[R];
        if (SWAPIN) load_pairs<R>(buf, idx0, S * TXP, i, r);
        else load_pairs<R>(buf, idx0, S * TXP, r, i);
        if (INV) {
            if (S > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) cmulc(r[k], i[k], tw[j * (k * tstep)]);
            }
            Dft<R>::run(i, r);
        } else {
            Dft<R>::run(r, i);
            if (S > 1) {
#pragma unroll
                for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * (m * tstep)]);
            }
        }
        store_pairs<R>(buf, idx0, S * TXP, r, i);
    }
}

__device__ __forceinline__ float4 sld(const float2* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void sst(float2* p, p2 r, p2 i)
{
    *reinterpret_cast<float4*>(p) = make_float4(r.x, r.y, i.x, i.y);
}

template <int R>
__device__ __forceinline__ void ssplit(const float4* v, p2* r, p2* i)
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        r[k] = make_float2(v[k].x, v[k].y);
        i[k] = make_float2(v[k].z, v[k].w);
    }
}

// ---- global row addressing of the strided side ----------------------------------------------------
// rows(r0, dr) = pointer to transform row r0 + dr (dr is a compile-time constant after unrolling).
struct RowsLinear {
    float2* base;
    size_t stride;
    __device__ __forceinline__ float2* operator()(int r0, int dr) const
    {
        return base + (size_t)r0 * stride + (size_t)dr * stride;
    }
};
// split layout (ColArgs): block r / splitRows of the axis lives in its own buffer -- a slice of the local
// exchange buffer, or ANOTHER GPU's buffer mapped over NVLink (the fused compute + exchange path)
struct RowsSplit {
    float2* const* peers;
    float2* local;
    long long blockStride;
    long long offset;   // peer offset + group * splitGroup + first column
    int splitRows;
    float invRows;
    size_t stride;
    __device__ __forceinline__ float2* operator()(int r0, int dr) const
    {
        const int r = r0 + dr;
        const int blk = __float2int_rz(((float)r + 0.5f) * invRows);
        float2* b0 = peers ? peers[blk] : local + (size_t)blk * blockStride;
        return b0 + offset + (size_t)(r - blk * splitRows) * stride;
    }
};

// ---- first forward stage: global rows j + k*S -> smem --------------------------------------------
// U butterflies are loaded before any arithmetic (U*R float4 in flight per thread).
template <int R, int L, int NW, int U, bool MASKED, int TXP = 8, bool SWAPIN = false>
__device__ __forceinline__ void sfirst_fwd(const float2* __restrict__ base, size_t stride, float4* __restrict__ sm,
                                           const float4* __restrict__ tw, int cp, int w,
                                           const unsigned char* __restrict__ rowMask)
{
    constexpr int S = L / R;
    constexpr int ITER = (S + NW * U - 1) / (NW * U);
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        float4 v[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = w + (it * U + u) * NW;
            if ((S % (NW * U)) != 0 && j >= S) continue;
            const float2* p = base + (size_t)j * stride;
#pragma unroll
            for (int k = 0; k < R; ++k) {
                if (MASKED) {
                    v[u][k] = make_float4(0.f, 0.f, 0.f, 0.f);
                    if (rowMask[j + k * S]) v[u][k] = sld(p + (size_t)(k * S) * stride);
                } else {
                    v[u][k] = sld(p + (size_t)(k * S) * stride);
                }
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = w + (it * U + u) * NW;
            if ((S % (NW * U)) != 0 && j >= S) continue;
            p2 r[R], i[R];
            if (SWAPIN) ssplit<R>(v[u], i, r);
            else ssplit<R>(v[u], r, i);
            Dft<R>::run(r, i);
#pragma unroll
            for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * m]);
            store_pairs<R>(sm, j * TXP + cp, S * TXP, r, i);
        }
    }
}

// the same with the rows addressed through a functor (split / peer layout on the INPUT side: pull exchange)
template <int R, int L, int NW, int U, int TXP, class ROWS>
__device__ __forceinline__ void sfirst_fwd_rows(const ROWS& rows, float4* __restrict__ sm, const float4* __restrict__ tw,
                                                int cp, int w)
{
    constexpr int S = L / R;
    constexpr int ITER = (S + NW * U - 1) / (NW * U);
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        float4 v[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = w + (it * U + u) * NW;
            if ((S % (NW * U)) != 0 && j >= S) continue;
#pragma unroll
            for (int k = 0; k < R; ++k) v[u][k] = sld(rows(j, k * S));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int j = w + (it * U + u) * NW;
            if ((S % (NW * U)) != 0 && j >= S) continue;
            p2 r[R], i[R];
            ssplit<R>(v[u], r, i);
            Dft<R>::run(r, i);
#pragma unroll
            for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * m]);
            store_pairs<R>(sm, j * TXP + cp, S * TXP, r, i);
        }
    }
}

// ---- first inverse stage: global rows rev(b*R) + k*(L/R) -> smem positions b*R + k ----------------
template <int R, int L, int NW, int U, int TXP = 8, class ROWS>
__device__ __forceinline__ void sfirst_inv(const ROWS& rows, float4* __restrict__ sm, const int* __restrict__ rev, int cp,
                                           int w)
{
    constexpr int nb = L / R, fs = L / R;
    constexpr int ITER = (nb + NW * U - 1) / (NW * U);
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        float4 v[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = w + (it * U + u) * NW;
            if ((nb % (NW * U)) != 0 && b >= nb) continue;
            const int r0 = __ldg(rev + b * R);
#pragma unroll
            for (int k = 0; k < R; ++k) v[u][k] = sld(rows(r0, k * fs));
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = w + (it * U + u) * NW;
            if ((nb % (NW * U)) != 0 && b >= nb) continue;
            p2 r[R], i[R];
            ssplit<R>(v[u], r, i);
            Dft<R>::run(i, r);
            store_pairs<R>(sm, b * R * TXP + cp, TXP, r, i);
        }
    }
}

// ---- last forward stage: smem positions b*R + k -> global rows rev(b*R) + m*(L/R) -----------------
template <int R, int L, int NW, int TXP = 8, bool SWAPOUT = false, class ROWS>
__device__ __forceinline__ void slast_fwd(const ROWS& rows, const float4* __restrict__ sm, const int* __restrict__ rev,
                                          int cp, int w)
{
    constexpr int nb = L / R, fs = L / R;
    constexpr int ITER = (nb + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        p2 r[R], i[R];
        load_pairs<R>(sm, b * R * TXP + cp, TXP, r, i);
        const int r0 = __ldg(rev + b * R);
        Dft<R>::run(r, i);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            if (SWAPOUT) sst(rows(r0, m * fs), i[m], r[m]);
            else sst(rows(r0, m * fs), r[m], i[m]);
        }
    }
}

// ---- last inverse stage: smem positions j + k*S -> global rows j + m*S -------------------------------
template <int R, int L, int NW, int TXP = 8, class ROWS>
__device__ __forceinline__ void slast_inv(const ROWS& rows, const float4* __restrict__ sm, const float4* __restrict__ tw,
                                          int cp, int w)
{
    constexpr int S = L / R;
    constexpr int ITER = (S + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int j = w + it * NW;
        if ((S % NW) != 0 && j >= S) break;
        p2 r[R], i[R];
        load_pairs<R>(sm, j * TXP + cp, S * TXP, r, i);
#pragma unroll
        for (int k = 1; k < R; ++k) cmulc(r[k], i[k], tw[j * k]);
        Dft<R>::run(i, r);
#pragma unroll
        for (int m = 0; m < R; ++m) sst(rows(j, m * S), r[m], i[m]);
    }
}

// (base, stride) forms
template <int R, int L, int NW, int U, int TXP = 8>
__device__ __forceinline__ void sfirst_inv(const float2* __restrict__ base, size_t stride, float4* __restrict__ sm,
                                           const int* __restrict__ rev, int cp, int w)
{
    sfirst_inv<R, L, NW, U, TXP>(RowsLinear{const_cast<float2*>(base), stride}, sm, rev, cp, w);
}
template <int R, int L, int NW, int TXP = 8, bool SWAPOUT = false>
__device__ __forceinline__ void slast_fwd(float2* __restrict__ base, size_t stride, const float4* __restrict__ sm,
                                          const int* __restrict__ rev, int cp, int w)
{
    slast_fwd<R, L, NW, TXP, SWAPOUT>(RowsLinear{base, stride}, sm, rev, cp, w);
}
template <int R, int L, int NW, int TXP = 8>
__device__ __forceinline__ void slast_inv(float2* __restrict__ base, size_t stride, const float4* __restrict__ sm,
                                          const float4* __restrict__ tw, int cp, int w)
{
    slast_inv<R, L, NW, TXP>(RowsLinear{base, stride}, sm, tw, cp, w);
}

// ---- fused middle: last forward stage, x H x c, first inverse stage (in registers) ---------------
template <int R, int L, int NW, int U, int TXP = 8>
__device__ __forceinline__ void smid_fused(const float2* __restrict__ hbase, size_t stride, float4* __restrict__ sm,
                                           const int* __restrict__ rev, int cp, int w, float c)
{
    constexpr int nb = L / R, fs = L / R;
    constexpr int ITER = (nb + NW * U - 1) / (NW * U);
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        float4 h[U][R];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = w + (it * U + u) * NW;
            if ((nb % (NW * U)) != 0 && b >= nb) continue;
            const float2* p = hbase + (size_t)__ldg(rev + b * R) * stride;
#pragma unroll
            for (int k = 0; k < R; ++k) h[u][k] = sld(p + (size_t)(k * fs) * stride);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int b = w + (it * U + u) * NW;
            if ((nb % (NW * U)) != 0 && b >= nb) continue;
            p2 r[R], i[R];
            load_pairs<R>(sm, b * R * TXP + cp, TXP, r, i);
            Dft<R>::run(r, i);
            // Dst = c * (Src * Dst), Src = PSF spectrum (reference mulAndScale, src/convolution3Dfft.cu:41-45)
#pragma unroll
            for (int m = 0; m < R; ++m) {
                const p2 hr = make_float2(h[u][m].x, h[u][m].y), hi = make_float2(h[u][m].z, h[u][m].w);
                const p2 xr = pmuls(pfma(hr, r[m], pneg(pmul(hi, i[m]))), c);
                const p2 xi = pmuls(pfma(hi, r[m], pmul(hr, i[m])), c);
                r[m] = xr;
                i[m] = xi;
            }
            Dft<R>::run(i, r);
            store_pairs<R>(sm, b * R * TXP + cp, TXP, r, i);
        }
    }
}

// ---- whole transforms on a shared-memory tile (used by the x passes) ------------------------------
template <class P, int NW, bool INV>
__device__ __forceinline__ void sengine(float4* sm, const float4* tw, int cp, int w)
{
    constexpr int L = P::L;
    if (!INV) {
        sstage<P::R0, L, L, NW, false>(sm, tw, cp, w);
        __syncthreads();
        sstage<P::R1, L, L / P::R0, NW, false>(sm, tw, cp, w);
        __syncthreads();
        if constexpr (P::ns >= 3) {
            sstage<P::R2, L, L / (P::R0 * P::R1), NW, false>(sm, tw, cp, w);
            __syncthreads();
        }
        if constexpr (P::ns >= 4) {
            sstage<P::R3, L, L / (P::R0 * P::R1 * P::R2), NW, false>(sm, tw, cp, w);
            __syncthreads();
        }
    } else {
        if constexpr (P::ns >= 4) {
            sstage<P::R3, L, P::R3, NW, true>(sm, tw, cp, w);
            __syncthreads();
        }
        if constexpr (P::ns >= 3) {
            sstage<P::R2, L, P::R2 * P::R3, NW, true>(sm, tw, cp, w);
            __syncthreads();
        }
        sstage<P::R1, L, P::R1 * P::R2 * P::R3, NW, true>(sm, tw, cp, w);
        __syncthreads();
        sstage<P::R0, L, L, NW, true>(sm, tw, cp, w);
        __syncthreads();
    }
}

}  // namespace fcb200
