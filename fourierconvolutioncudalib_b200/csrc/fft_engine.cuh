// Shared-memory FFT engine for sm_100a: mixed-radix, in-place decimation-in-frequency forward,
// mirrored decimation-in-time inverse, on a tile of pencils held as sm[position][column-pair].
//
// Tile layout: one tile row per transform position, TXP float4 per row; a float4 holds the same
// position of two adjacent pencils in PAIR-PLANAR form (re0, re1, im0, im1), so the real parts and
// the imaginary parts of the two pencils are aligned register pairs and every butterfly runs on
// Blackwell's packed fp32 instructions (fft_butterflies.cuh).  With TXP = 8 a row is 128 bytes = all
// 32 banks, and the 8 lanes that share a worker index read one full row: every shared-memory access
// of the engine is conflict-free without padding.  Thread t works on column pair cp = t % TXP as
// worker w = t / TXP; workers split the butterflies of a stage.
//
// Forward (DIF, in place): after all stages position p holds frequency rev[p] (mixed-radix digit
// reversal).  Inverse (DIT, stages mirrored) takes that order back to natural order.  Callers map
// positions to global rows, so natural order in global memory costs nothing for the strided axes.
//
// Radices 2,3,4,5,6,7,8,9,10,12,15,16 run in registers (6..15: composite, fft_butterflies.cuh).  Any other prime factor p runs as a direct O(p) sum per
// output between two tile buffers (ping-pong), so every length the C ABI can receive is supported
// (the reference's own tests use 79, 109, 173, 37, 23, 53, ... -- SURVEY.md section 4).
//
// Replaces the cuFFT plan/exec calls of /root/reference/src/convolution3Dfft.cu:519-525, :544-547.
#pragma once
#include "fft_butterflies.cuh"
#include "fc_common.h"

namespace fcb200 {

// twiddle table in shared memory: float4 (c, c, s, s) per root, from the global float2 (c, s) table
__device__ __forceinline__ void load_twiddles(float4* tw_s, const float2* tw_g, int L)
{
    for (int i = threadIdx.x; i < L; i += blockDim.x) {
        const float2 t = __ldg(tw_g + i);
        tw_s[i] = make_float4(t.x, t.x, t.y, t.y);
    }
}

template <int R>
__device__ __forceinline__ void load_pairs(const float4* __restrict__ buf, int idx0, int step, p2* r, p2* i)
{
#pragma unroll
    for (int k = 0; k < R; ++k) {
        const float4 v = buf[idx0 + k * step];
        r[k] = make_float2(v.x, v.y);
        i[k] = make_float2(v.z, v.w);
    }
}

template <int R>
__device__ __forceinline__ void store_pairs(float4* __restrict__ buf, int idx0, int step, const p2* r, const p2* i)
{
#pragma unroll
    for (int m = 0; m < R; ++m) buf[idx0 + m * step] = make_float4(r[m].x, r[m].y, i[m].x, i[m].y);
}

// One in-register radix-R stage, shared memory -> shared memory, in place.
//   L   transform length, Li current block length (forward: before the stage; inverse: after it)
template <int R, bool INV>
__device__ __forceinline__ void stage_smem(float4* __restrict__ buf, const float4* __restrict__ tw, int L, int Li,
                                           int cp, int w, int W, int txp)
{
    const int S = Li / R;
    const int nb = L / R;
    const int tstep = L / Li;
    const float invS = 1.0f / (float)S;
    for (int b = w; b < nb; b += W) {
        // b / S without an integer division: exact for b < 2^16 (checked exhaustively on the host)
        const int beta = __float2int_rz(((float)b + 0.5f) * invS);
        const int j = b - beta * S;
        const int idx0 = (beta * Li + j) * txp + cp;
        p2 r[R], i[R];
        load_pairs<R>(buf, idx0, S * txp, r, i);
        if (INV) {
            if (S > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) cmulc(r[k], i[k], tw[j * k * tstep]);
            }
            Dft<R>::run(i, r);
        } else {
            Dft<R>::run(r, i);
            if (S > 1) {
#pragma unroll
                for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * m * tstep]);
            }
        }
        store_pairs<R>(buf, idx0, S * txp, r, i);
    }
}

// Generic prime radix p (run-time), src -> dst (distinct tile buffers).
template <bool INV>
__device__ __forceinline__ void stage_generic(const float4* __restrict__ src, float4* __restrict__ dst,
                                              const float4* __restrict__ tw, int L, int Li, int p, int cp, int w,
                                              int W, int txp)
{
    const int S = Li / p;
    const int tstep = L / Li;
    const int rstep = L / p;
    const float invS = 1.0f / (float)S, invLi = 1.0f / (float)Li;
    for (int o = w; o < L; o += W) {
        const int beta = __float2int_rz(((float)o + 0.5f) * invLi);
        const int rr = o - beta * Li;
        const int m = __float2int_rz(((float)rr + 0.5f) * invS);
        const int j = rr - m * S;
        const int base = beta * Li + j;
        // forward: y_m = w_Li^{j m} * sum_k x_k w_p^{k m}
        // inverse: y_m = sum_k x_k conj(w_Li^{j k} w_p^{k m})      (both roots come from the one table)
        const int inc = INV ? (j * tstep + m * rstep) : (m * rstep);
        int idx = 0;
        p2 ar = make_float2(0.f, 0.f), ai = make_float2(0.f, 0.f);
        for (int k = 0; k < p; ++k) {
            const float4 t = tw[idx];
            const float4 v = src[(base + k * S) * txp + cp];
            const p2 C = make_float2(t.x, t.y), Sn = make_float2(t.z, t.w);
            const p2 xr = make_float2(v.x, v.y), xi = make_float2(v.z, v.w);
            if (INV) {
                ar = pfma(xr, C, pfma(xi, Sn, ar));
                ai = pfma(xi, C, pfma(xr, pneg(Sn), ai));
            } else {
                ar = pfma(xr, C, pfma(xi, pneg(Sn), ar));
                ai = pfma(xi, C, pfma(xr, Sn, ai));
            }
            idx += inc;
            if (idx >= L) idx -= L;
        }
        if (!INV && S > 1) cmul(ar, ai, tw[j * m * tstep]);
        dst[o * txp + cp] = make_float4(ar.x, ar.y, ai.x, ai.y);
    }
}

__device__ __forceinline__ bool is_fast_radix(int R)
{
    return R == 1 || R == 2 || R == 3 || R == 4 || R == 5 || R == 6 || R == 7 || R == 8 || R == 9 || R == 10 || R == 12 ||
           R == 15 || R == 16;
}

template <bool INV>
__device__ __forceinline__ void stage_dispatch(int R, float4*& cur, float4*& oth, const float4* tw, int L, int Li,
                                               int cp, int w, int W, int txp, bool active)
{
    if (active) {
        switch (R) {
            case 1: break;
            case 2: stage_smem<2, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 3: stage_smem<3, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 4: stage_smem<4, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 5: stage_smem<5, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 6: stage_smem<6, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 7: stage_smem<7, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 8: stage_smem<8, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 9: stage_smem<9, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 10: stage_smem<10, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 12: stage_smem<12, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 15: stage_smem<15, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            case 16: stage_smem<16, INV>(cur, tw, L, Li, cp, w, W, txp); break;
            default: stage_generic<INV>(cur, oth, tw, L, Li, R, cp, w, W, txp); break;
        }
    }
    if (!is_fast_radix(R)) {
        float4* t = cur;
        cur = oth;
        oth = t;
    }
}

// Runs all stages of `P` on the tile in `A` (second buffer `B` only needed when P.generic).
// Ends with a __syncthreads(); returns the buffer that holds the result.
//   forward: natural order in, position p holds frequency P.rev[p] out
//   inverse: the mirror image (scaled by L, like cuFFT's unnormalised inverse)
template <bool INV>
__device__ __forceinline__ float4* engine_run(const AxisPlanDev& P, float4* A, float4* B, const float4* tw, int cp,
                                              int w, int W, int txp, bool active)
{
    float4* cur = A;
    float4* oth = B;
    if (!INV) {
        int Li = P.L;
        for (int s = 0; s < P.ns; ++s) {
            const int R = P.radix[s];
            stage_dispatch<false>(R, cur, oth, tw, P.L, Li, cp, w, W, txp, active);
            Li /= R;
            __syncthreads();
        }
    } else {
        int Li = 1;
        for (int s = P.ns - 1; s >= 0; --s) {
            const int R = P.radix[s];
            Li *= R;
            stage_dispatch<true>(R, cur, oth, tw, P.L, Li, cp, w, W, txp, active);
            __syncthreads();
        }
    }
    return cur;
}

}  // namespace fcb200
