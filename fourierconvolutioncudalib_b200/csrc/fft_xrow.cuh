// Register-resident row-wise x passes for two-stage plans, M = R*R complex points per row
// (nx = 2*R*R real samples: R = 16 -> nx = 512, R = 8 -> nx = 128).
//
// A group of R consecutive lanes owns one PAIR of adjacent rows (packed fp32: lo = row a, hi = row b).
// Lanes run along x, so every global load / store is coalesced straight from / to registers:
//   forward : thread j loads z[j + R*k] (k = 0..R-1), radix-R butterfly, twiddle, ONE exchange through
//             padded shared memory (write j + (R+1)*m, read (R+1)*j + k: both conflict-free), second
//             radix-R butterfly; the thread then holds Z[j + R*m].  The split of the packed half-length
//             transform into the spectrum of the real rows needs Z[M-k] (held by another lane): Z makes
//             one more trip through the exchange buffer.  Pair-planar output rows are assembled with a
//             neighbour-lane exchange and stored 8 bytes per lane, 128 bytes contiguous per group.
//   inverse : the mirror image.
// No CTA barrier after the twiddle table is built (only __syncwarp), ~5x less shared-memory traffic than
// the transposing two-tile kernels (fft_xpass.cuh), which remain the path for every other length.
#pragma once
#include "fft_engine.cuh"
#include "fft_kernels.h"

namespace fcb200 {

template <int R>
struct XRow {
    static constexpr int M = R * R;
    static constexpr int PADM = M + R;   // position p is stored at p + p / R
};

// padded exchange index of bin / position p
template <int R>
__device__ __forceinline__ int xpad(int p) { return p + p / R; }

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS) xrow_fwd_kernel(XArgs a)
{
    constexpr int M = XRow<R>::M, PADM = XRow<R>::PADM, RP = THREADS / R;
    extern __shared__ float4 smem[];
    float4* twt = smem;                    // [m][j] = w_M^(j*m) as (c, c, s, s)
    float4* xch = smem + R * R;            // [RP][PADM]

    const Geometry g = a.g;
    const int t = threadIdx.x, lane = t & 31;
    const int grp = t / R, j = t % R;
    pdl_launch_dependents();
    for (int idx = t; idx < R * R; idx += THREADS) {
        const int m = idx / R, jj = idx % R;
        const float2 w = __ldg(a.P.tw + jj * m);
        twt[idx] = make_float4(w.x, w.x, w.y, w.y);
    }
    pdl_wait();
    __syncthreads();

    // grid-stride over blocks of RP row pairs (persistent launch: the twiddle table is built once per CTA)
    const long long nblk = (a.nrows + 2 * RP - 1) / (2 * RP);
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
    const long long rowA = (blk * RP + grp) * 2;
    const bool hasA = rowA < a.nrows, hasB = rowA + 1 < a.nrows;
    const float2* srcA = reinterpret_cast<const float2*>(a.in_real + rowA * g.nx);
    const float2* srcB = reinterpret_cast<const float2*>(a.in_real + (rowA + 1) * g.nx);

    p2 r[R], i[R];
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float2 ua = hasA ? __ldg(srcA + j + R * k) : make_float2(0.f, 0.f);
        const float2 ub = hasB ? __ldg(srcB + j + R * k) : make_float2(0.f, 0.f);
        r[k] = make_float2(ua.x, ub.x);
        i[k] = make_float2(ua.y, ub.y);
    }
    Dft<R>::run(r, i);
#pragma unroll
    for (int m = 1; m < R; ++m) cmul(r[m], i[m], twt[m * R + j]);

    float4* x = xch + (size_t)grp * PADM;
#pragma unroll
    for (int m = 0; m < R; ++m) x[j + (R + 1) * m] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float4 v = x[(R + 1) * j + k];
        r[k] = make_float2(v.x, v.y);
        i[k] = make_float2(v.z, v.w);
    }
    Dft<R>::run(r, i);   // r[m], i[m] = Z[j + R*m]

    // Nyquist bin from Z[0] (lane 0 of the group), before the arrays are overwritten
    const p2 nyq = psub(r[0], i[0]);

    // X[k] = E + w^k O,  E = (Z[k] + conj Z[M-k]) / 2,  O = -i (Z[k] - conj Z[M-k]) / 2.
    // Z goes through the exchange buffer once more (natural order) so that every lane can read Z[M-k].
    __syncwarp();
#pragma unroll
    for (int m = 0; m < R; ++m) x[j + (R + 1) * m] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);   // bin j + R*m
    __syncwarp();
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int k = j + R * m;
        const float4 pb = x[xpad<R>((M - k) & (M - 1))];   // Z[M] == Z[0]
        const p2 pr = make_float2(pb.x, pb.y), pi = make_float2(pb.z, pb.w);
        const float2 tk = __ldg(a.twx + k);   // exp(-2*pi*i*k/nx)
        const p2 er = pmuls(padd(r[m], pr), 0.5f), ei = pmuls(psub(i[m], pi), 0.5f);
        const p2 orr = pmuls(padd(i[m], pi), 0.5f), oi = pmuls(psub(r[m], pr), -0.5f);
        r[m] = padd(er, pfmas(orr, tk.x, pmuls(oi, -tk.y)));
        i[m] = padd(ei, pfmas(oi, tk.x, pmuls(orr, tk.y)));
    }

    // pair-planar rows: float2 slot k holds (re_k, re_k+1) for even k, (im_k-1, im_k) for odd k
    float2* dstA = a.spec + rowA * g.xcp;
    float2* dstB = a.spec + (rowA + 1) * g.xcp;
    const bool odd = lane & 1;
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const p2 send = odd ? r[m] : i[m];
        p2 got;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        const float2 oa = odd ? make_float2(got.x, i[m].x) : make_float2(r[m].x, got.x);
        const float2 ob = odd ? make_float2(got.y, i[m].y) : make_float2(r[m].y, got.y);
        if (hasA) dstA[j + R * m] = oa;
        if (hasB) dstB[j + R * m] = ob;
    }
    if (j == 0) {   // Nyquist bin (purely real) and the zero pad columns [M+1, xcp)
        if (hasA) {
            dstA[M] = make_float2(nyq.x, 0.f);
            for (int k = M + 1; k < g.xcp; ++k) dstA[k] = make_float2(0.f, 0.f);
        }
        if (hasB) {
            dstB[M] = make_float2(nyq.y, 0.f);
            for (int k = M + 1; k < g.xcp; ++k) dstB[k] = make_float2(0.f, 0.f);
        }
    }
    __syncwarp();   // the exchange buffer is free for the next block of rows
    }
}

template <int R, int THREADS>
__global__ void __launch_bounds__(THREADS) xrow_inv_kernel(XArgs a)
{
    constexpr int M = XRow<R>::M, PADM = XRow<R>::PADM, RP = THREADS / R;
    extern __shared__ float4 smem[];
    float4* twt = smem;
    float4* xch = smem + R * R;

    const Geometry g = a.g;
    const int t = threadIdx.x, lane = t & 31;
    const int grp = t / R, j = t % R;
    pdl_launch_dependents();
    for (int idx = t; idx < R * R; idx += THREADS) {
        const int m = idx / R, jj = idx % R;
        const float2 w = __ldg(a.P.tw + jj * m);
        twt[idx] = make_float4(w.x, w.x, w.y, w.y);
    }
    pdl_wait();
    __syncthreads();

    const bool odd = lane & 1;
    const long long nblk = (a.nrows + 2 * RP - 1) / (2 * RP);
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {   // persistent, see xrow_fwd_kernel
    const long long rowA = (blk * RP + grp) * 2;
    const bool hasA = rowA < a.nrows, hasB = rowA + 1 < a.nrows;
    const float2* srcA = a.spec + rowA * g.xcp;
    const float2* srcB = a.spec + (rowA + 1) * g.xcp;

    // pair-planar rows -> (re, im) of bin k = j + R*m for both rows
    p2 r[R], i[R];
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const float2 va = hasA ? srcA[j + R * m] : make_float2(0.f, 0.f);
        const float2 vb = hasB ? srcB[j + R * m] : make_float2(0.f, 0.f);
        const p2 send = odd ? make_float2(va.x, vb.x) : make_float2(va.y, vb.y);
        p2 got;
        got.x = __shfl_xor_sync(0xffffffffu, send.x, 1);
        got.y = __shfl_xor_sync(0xffffffffu, send.y, 1);
        r[m] = odd ? got : make_float2(va.x, vb.x);
        i[m] = odd ? make_float2(va.y, vb.y) : got;
    }
    // Z[k] = s + i d,  s = X[k] + conj X[M-k],  d = (X[k] - conj X[M-k]) conj(w^k);
    // X goes through the exchange buffer (natural order) so that every lane can read X[M-k]
    float4* x = xch + (size_t)grp * PADM;
#pragma unroll
    for (int m = 0; m < R; ++m) x[j + (R + 1) * m] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
    __syncwarp();
    const p2 x0r = r[0];
#pragma unroll
    for (int m = 0; m < R; ++m) {
        const int k = j + R * m;
        const float4 pb = x[xpad<R>((M - k) & (M - 1))];   // k = 0 is fixed up below (pairs with the Nyquist bin)
        const p2 pr = make_float2(pb.x, pb.y), pi = make_float2(pb.z, pb.w);
        const float2 tk = __ldg(a.twx + k);
        const p2 sr = padd(r[m], pr), si = psub(i[m], pi);
        const p2 Dr = psub(r[m], pr), Di = padd(i[m], pi);
        const p2 dr = pfmas(Dr, tk.x, pmuls(Di, tk.y));
        const p2 di = pfmas(Di, tk.x, pmuls(Dr, -tk.y));
        r[m] = psub(sr, di);
        i[m] = padd(si, dr);
    }
    __syncwarp();   // everybody has read its partners before the buffer is reused
    if (j == 0) {   // k = 0 pairs with the Nyquist bin, not with itself
        const p2 xm = make_float2(hasA ? srcA[M].x : 0.f, hasB ? srcB[M].x : 0.f);
        r[0] = padd(x0r, xm);
        i[0] = psub(x0r, xm);
    }
    // the thread holds Z[j + R*m] = block j of the digit-reversed positions: first inverse stage
    Dft<R>::run(i, r);
#pragma unroll
    for (int k = 0; k < R; ++k) x[(R + 1) * j + k] = make_float4(r[k].x, r[k].y, i[k].x, i[k].y);
    __syncwarp();
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float4 v = x[j + (R + 1) * k];
        r[k] = make_float2(v.x, v.y);
        i[k] = make_float2(v.z, v.w);
    }
#pragma unroll
    for (int k = 1; k < R; ++k) cmulc(r[k], i[k], twt[k * R + j]);
    Dft<R>::run(i, r);   // z[j + R*m] = (x[2p], x[2p+1]), unnormalised

    float2* dstA = reinterpret_cast<float2*>(a.out_real + rowA * g.nx);
    float2* dstB = reinterpret_cast<float2*>(a.out_real + (rowA + 1) * g.nx);
#pragma unroll
    for (int m = 0; m < R; ++m) {
        if (hasA) dstA[j + R * m] = make_float2(r[m].x, i[m].x);
        if (hasB) dstB[j + R * m] = make_float2(r[m].y, i[m].y);
    }
    __syncwarp();   // the exchange buffer is free for the next block of rows
    }
}

}  // namespace fcb200
