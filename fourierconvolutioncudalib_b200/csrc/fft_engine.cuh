// Shared-memory FFT engine for sm_100a: mixed-radix, in-place decimation-in-frequency forward,
// mirrored decimation-in-time inverse, on a tile of pencils held as sm[position][column-pair].
//
// Tile layout: one tile row per transform position, TXP float4 per row; a float4 holds the same
// position of two adjacent pencils (re0, im0, re1, im1).  With TXP = 8 a row is 128 bytes = all 32
// banks, and the 8 lanes that share a worker index read one full row: every shared-memory access
// of the engine is conflict-free without padding.  Thread t works on column pair cp = t % TXP as
// worker w = t / TXP; workers split the butterflies of a stage.
//
// Forward (DIF, in place): after all stages position p holds frequency rev[p] (mixed-radix digit
// reversal).  Inverse (DIT, stages mirrored) takes that order back to natural order.  Callers map
// positions to global rows, so natural order in global memory costs nothing for the strided axes.
//
// Radices 2,3,4,5,7,8 run in registers.  Any other prime factor p runs as a direct O(p) sum per
// output between two tile buffers (ping-pong), so every length the C ABI can receive is supported
// (the reference's own tests use 79, 109, 173, 37, 23, 53, ... -- SURVEY.md section 4).
//
// Replaces the cuFFT plan/exec calls of /root/reference/src/convolution3Dfft.cu:519-525, :544-547.
#pragma once
#include "fft_butterflies.cuh"
#include "fc_common.h"

namespace fcb200 {

// XOR swizzle of the column-pair slot, used by the X pass whose tile is filled by a transposing
// load (lanes run along positions there).  Bijective over 8 consecutive positions, over the 8 even
// and over the 8 odd positions of a 16-aligned group (tests/engine_model.py: swz).
__device__ __forceinline__ int swz8(int pos) { return (pos ^ (pos >> 3)) & 7; }

template <bool SWZ>
__device__ __forceinline__ int tile_idx(int pos, int cp, int txp)
{
    return SWZ ? (pos * 8 + (cp ^ swz8(pos))) : (pos * txp + cp);
}

// One in-register radix-R stage, shared memory -> shared memory, in place.
//   L   transform length, Li current block length (forward: before the stage; inverse: after it)
template <int R, bool INV, bool SWZ>
__device__ __forceinline__ void stage_smem(float4* __restrict__ buf, const float2* __restrict__ tw, int L, int Li,
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
        const int base = beta * Li + j;
        float ar[R], ai[R], br[R], bi[R];
#pragma unroll
        for (int k = 0; k < R; ++k) {
            float4 v = buf[tile_idx<SWZ>(base + k * S, cp, txp)];
            ar[k] = v.x;
            ai[k] = v.y;
            br[k] = v.z;
            bi[k] = v.w;
        }
        if (INV) {
            if (S > 1) {
#pragma unroll
                for (int k = 1; k < R; ++k) {
                    float2 t = tw[j * k * tstep];
                    cmulc(ar[k], ai[k], t.x, t.y);
                    cmulc(br[k], bi[k], t.x, t.y);
                }
            }
            Dft<R>::run(ai, ar);
            Dft<R>::run(bi, br);
        } else {
            Dft<R>::run(ar, ai);
            Dft<R>::run(br, bi);
            if (S > 1) {
#pragma unroll
                for (int m = 1; m < R; ++m) {
                    float2 t = tw[j * m * tstep];
                    cmul(ar[m], ai[m], t.x, t.y);
                    cmul(br[m], bi[m], t.x, t.y);
                }
            }
        }
#pragma unroll
        for (int m = 0; m < R; ++m)
            buf[tile_idx<SWZ>(base + m * S, cp, txp)] = make_float4(ar[m], ai[m], br[m], bi[m]);
    }
}

// Generic prime radix p (run-time), src -> dst (distinct tile buffers).
template <bool INV, bool SWZ>
__device__ __forceinline__ void stage_generic(const float4* __restrict__ src, float4* __restrict__ dst,
                                              const float2* __restrict__ tw, int L, int Li, int p, int cp, int w,
                                              int W, int txp)
{
    const int S = Li / p;
    const int tstep = L / Li;
    const int rstep = L / p;
    const float invS = 1.0f / (float)S, invLi = 1.0f / (float)Li;
    for (int o = w; o < L; o += W) {
        const int beta = __float2int_rz(((float)o + 0.5f) * invLi);
        const int r = o - beta * Li;
        const int m = __float2int_rz(((float)r + 0.5f) * invS);
        const int j = r - m * S;
        const int base = beta * Li + j;
        // forward: y_m = w_Li^{j m} * sum_k x_k w_p^{k m}
        // inverse: y_m = sum_k x_k conj(w_Li^{j k} w_p^{k m})      (both roots come from the one table)
        const int inc = INV ? (j * tstep + m * rstep) : (m * rstep);
        int idx = 0;
        float a0 = 0.f, a1 = 0.f, b0 = 0.f, b1 = 0.f;
        for (int k = 0; k < p; ++k) {
            float2 t = tw[idx];
            float4 v = src[tile_idx<SWZ>(base + k * S, cp, txp)];
            if (INV) {
                a0 = fmaf(v.x, t.x, fmaf(v.y, t.y, a0));
                a1 = fmaf(v.y, t.x, fmaf(-v.x, t.y, a1));
                b0 = fmaf(v.z, t.x, fmaf(v.w, t.y, b0));
                b1 = fmaf(v.w, t.x, fmaf(-v.z, t.y, b1));
            } else {
                a0 = fmaf(v.x, t.x, fmaf(-v.y, t.y, a0));
                a1 = fmaf(v.y, t.x, fmaf(v.x, t.y, a1));
                b0 = fmaf(v.z, t.x, fmaf(-v.w, t.y, b0));
                b1 = fmaf(v.w, t.x, fmaf(v.z, t.y, b1));
            }
            idx += inc;
            if (idx >= L) idx -= L;
        }
        if (!INV && S > 1) {
            float2 t = tw[j * m * tstep];
            cmul(a0, a1, t.x, t.y);
            cmul(b0, b1, t.x, t.y);
        }
        dst[tile_idx<SWZ>(o, cp, txp)] = make_float4(a0, a1, b0, b1);
    }
}

template <bool INV, bool SWZ>
__device__ __forceinline__ void stage_dispatch(int R, float4*& cur, float4*& oth, const float2* tw, int L, int Li,
                                               int cp, int w, int W, int txp, bool active)
{
    bool swap = false;
    if (active) {
        switch (R) {
            case 1: break;
            case 2: stage_smem<2, INV, SWZ>(cur, tw, L, Li, cp, w, W, txp); break;
            case 3: stage_smem<3, INV, SWZ>(cur, tw, L, Li, cp, w, W, txp); break;
            case 4: stage_smem<4, INV, SWZ>(cur, tw, L, Li, cp, w, W, txp); break;
            case 5: stage_smem<5, INV, SWZ>(cur, tw, L, Li, cp, w, W, txp); break;
            case 7: stage_smem<7, INV, SWZ>(cur, tw, L, Li, cp, w, W, txp); break;
            case 8: stage_smem<8, INV, SWZ>(cur, tw, L, Li, cp, w, W, txp); break;
            default: stage_generic<INV, SWZ>(cur, oth, tw, L, Li, R, cp, w, W, txp); break;
        }
    }
    swap = !(R == 1 || R == 2 || R == 3 || R == 4 || R == 5 || R == 7 || R == 8);
    if (swap) {
        float4* t = cur;
        cur = oth;
        oth = t;
    }
}

// Runs all stages of `P` on the tile in `A` (second buffer `B` only needed when P.generic).
// Ends with a __syncthreads(); returns the buffer that holds the result.
//   forward: natural order in, position p holds frequency P.rev[p] out
//   inverse: the mirror image (scaled by L, like cuFFT's unnormalised inverse)
template <bool INV, bool SWZ>
__device__ __forceinline__ float4* engine_run(const AxisPlanDev& P, float4* A, float4* B, const float2* tw, int cp,
                                              int w, int W, int txp, bool active)
{
    float4* cur = A;
    float4* oth = B;
    if (!INV) {
        int Li = P.L;
        for (int s = 0; s < P.ns; ++s) {
            const int R = P.radix[s];
            stage_dispatch<false, SWZ>(R, cur, oth, tw, P.L, Li, cp, w, W, txp, active);
            Li /= R;
            __syncthreads();
        }
    } else {
        int Li = 1;
        for (int s = P.ns - 1; s >= 0; --s) {
            const int R = P.radix[s];
            Li *= R;
            stage_dispatch<true, SWZ>(R, cur, oth, tw, P.L, Li, cp, w, W, txp, active);
            __syncthreads();
        }
    }
    return cur;
}

}  // namespace fcb200
