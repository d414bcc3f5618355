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
// Radices 2,3,4,5,6,7,8,9,10,12,15,16 and the primes 11,13,17,19,23 run in registers (fft_butterflies.cuh).  Any larger prime factor p runs as a direct O(p) sum per
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

// roots of the n-point transform of a Rader stage (after the L main roots in shared memory)
__device__ __forceinline__ void load_rader_twiddles(float4* rtw_s, const AxisPlanDev& P)
{
    if (P.rader_p)
        for (int i = threadIdx.x; i < P.rader_n; i += blockDim.x) {
            const float2 t = __ldg(P.rader_tw + i);
            rtw_s[i] = make_float4(t.x, t.x, t.y, t.y);
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

// Generic prime radix p > 23 (run-time), src -> dst (distinct tile buffers): the symmetric direct DFT.
//   t_k = x_k + x_(p-k),  u_k = x_k - x_(p-k),  k = 1 .. H = (p-1)/2
//   sum_k x_k w_p^(+-km) = (x_0 + sum_k cos(2 pi k m / p) t_k)  -+  i (sum_k sin(2 pi k m / p) u_k)
// One work item = one sub-sequence (beta, j) and TWO output pairs (m, p-m), (m+1, p-m-1): every input pair is loaded
// once for four outputs and every root once for two -- 4 shared-memory loads and 12 packed operations per k for four
// outputs, against 8 loads and 16 operations in the plain one-output-at-a-time sum (158x158x218: 0.95 -> see
// profiles/r02_odd_sizes.jsonl).  The item with m = 1 also produces output 0.
template <bool INV>
__device__ __forceinline__ void stage_generic(const float4* __restrict__ src, float4* __restrict__ dst,
                                              const float4* __restrict__ tw, int L, int Li, int p, int cp, int w,
                                              int W, int txp)
{
    const int S = Li / p;
    const int tstep = L / Li;
    const int rstep = L / p;
    const int H = (p - 1) / 2, Qn = (H + 1) / 2;   // Qn items per sub-sequence
    const int nseq = L / p;                        // sub-sequences (beta, j)
    const float invQ = 1.0f / (float)Qn, invS = 1.0f / (float)S;
    for (int it = w; it < nseq * Qn; it += W) {
        const int sq = __float2int_rz(((float)it + 0.5f) * invQ);   // sub-sequence index = beta * S + j
        const int q = it - sq * Qn;
        const int beta = __float2int_rz(((float)sq + 0.5f) * invS);
        const int j = sq - beta * S;
        const int base = beta * Li + j;
        const int m1 = 1 + 2 * q, m2 = m1 + 1;
        const bool has2 = m2 <= H;
        // x_0 (inverse: inputs are twiddled by conj(w_Li^(j k)) first; k = 0 needs none)
        const float4 v0 = src[base * txp + cp];
        const p2 x0r = make_float2(v0.x, v0.y), x0i = make_float2(v0.z, v0.w);
        p2 a1r = x0r, a1i = x0i, a2r = x0r, a2i = x0i, s0r = x0r, s0i = x0i;
        p2 b1r = make_float2(0.f, 0.f), b1i = b1r, b2r = b1r, b2i = b1r;
        int e1 = 0, e2 = 0, ej = 0, ejn = 0;   // root indices k*m1*rstep, k*m2*rstep, j*k*tstep, j*(p-k)*tstep (mod L)
        const int inc1 = (m1 * rstep) % L, inc2 = (m2 * rstep) % L, incj = (j * tstep) % L;
        ejn = (int)(((long long)j * tstep * p) % L);
        for (int k = 1; k <= H; ++k) {
            e1 += inc1;
            if (e1 >= L) e1 -= L;
            e2 += inc2;
            if (e2 >= L) e2 -= L;
            const float4 va = src[(base + k * S) * txp + cp];
            const float4 vb = src[(base + (p - k) * S) * txp + cp];
            p2 xar = make_float2(va.x, va.y), xai = make_float2(va.z, va.w);
            p2 xbr = make_float2(vb.x, vb.y), xbi = make_float2(vb.z, vb.w);
            if (INV && S > 1) {
                ej += incj;
                if (ej >= L) ej -= L;
                ejn -= incj;
                if (ejn < 0) ejn += L;
                cmulc(xar, xai, tw[ej]);
                cmulc(xbr, xbi, tw[ejn]);
            }
            const p2 tr = padd(xar, xbr), ti = padd(xai, xbi), ur = psub(xar, xbr), ui = psub(xai, xbi);
            const float4 r1 = tw[e1];   // (c, c, -s, -s) of exp(-2 pi i k m1 / p)
            const p2 c1 = make_float2(r1.x, r1.y), n1 = make_float2(r1.z, r1.w);
            a1r = pfma(tr, c1, a1r);
            a1i = pfma(ti, c1, a1i);
            b1r = pfma(ur, n1, b1r);     // b = -sum sin * u
            b1i = pfma(ui, n1, b1i);
            if (has2) {
                const float4 r2 = tw[e2];
                const p2 c2 = make_float2(r2.x, r2.y), n2 = make_float2(r2.z, r2.w);
                a2r = pfma(tr, c2, a2r);
                a2i = pfma(ti, c2, a2i);
                b2r = pfma(ur, n2, b2r);
                b2i = pfma(ui, n2, b2i);
            }
            if (q == 0) {
                s0r = padd(s0r, tr);
                s0i = padd(s0i, ti);
            }
        }
        // with n = -sin:  forward  X_m = a - i*(sum sin u) = a + i*b,  X_(p-m) = a - i*b ;  inverse: the conjugate roots
        auto emit = [&](int m, p2 ar, p2 ai, p2 br, p2 bi) {
            p2 yr, yi, zr, zi;   // y = output m, z = output p - m
            if (INV) {
                yr = padd(ar, bi); yi = psub(ai, br);      // a - i*b
                zr = psub(ar, bi); zi = padd(ai, br);      // a + i*b
            } else {
                yr = psub(ar, bi); yi = padd(ai, br);      // a + i*b
                zr = padd(ar, bi); zi = psub(ai, br);      // a - i*b
                if (S > 1) {
                    cmul(yr, yi, tw[(int)(((long long)j * m * tstep) % L)]);
                    cmul(zr, zi, tw[(int)(((long long)j * (p - m) * tstep) % L)]);
                }
            }
            dst[(base + m * S) * txp + cp] = make_float4(yr.x, yr.y, yi.x, yi.y);
            dst[(base + (p - m) * S) * txp + cp] = make_float4(zr.x, zr.y, zi.x, zi.y);
        };
        emit(m1, a1r, a1i, b1r, b1i);
        if (has2) emit(m2, a2r, a2i, b2r, b2i);
        if (q == 0) dst[base * txp + cp] = make_float4(s0r.x, s0r.y, s0i.x, s0i.y);
    }
}

__device__ __forceinline__ bool is_fast_radix(int R)
{
    return R == 1 || R == 2 || R == 3 || R == 4 || R == 5 || R == 6 || R == 7 || R == 8 || R == 9 || R == 10 || R == 12 ||
           R == 15 || R == 16 || R == 11 || R == 13 || R == 17 || R == 19 || R == 23 || R == 14 || R == 18 || R == 20 ||
           R == 21 || R == 25 || R == 28;
}

// ---- Rader stage: a prime radix p as a cyclic convolution of length n = p - 1 ------------------------------------
// Only for the LAST radix (S = 1: pure DFTs of contiguous blocks of p positions, no stage twiddles), forward or inverse.
//   X[0] = sum_n x[n],   X[g^-q] = x[0] + (a (*) b)[q],   a[m] = x[g^m],   b[t] = w_p^(+-g^-t)
// Per block beta of the tile: gather a into the OTHER tile buffer (positions beta*p+1 .. beta*p+n), n-point DIF there
// (position order), multiply by the precomputed spectrum of b (stored in position order, 1/n folded in; x[0] is added to
// the DC term so that it reaches every output, and x[0] + A[0] is X[0]), n-point DIT back, scatter c[q] to position
// g^-q of the original buffer.  Result in `cur` again: in place as far as the engine is concerned.
// Cost: two n-point FFTs and three passes over the tile instead of (p-1)^2 / 2 multiply-adds per block.
template <bool INV>
__device__ __forceinline__ void rader_substage(int R, float4* buf, const float4* tw, int n, int Li, int cp, int w, int W,
                                               int txp)
{
    switch (R) {
        case 2: stage_smem<2, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 3: stage_smem<3, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 4: stage_smem<4, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 5: stage_smem<5, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 6: stage_smem<6, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 7: stage_smem<7, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 8: stage_smem<8, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 9: stage_smem<9, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 10: stage_smem<10, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 11: stage_smem<11, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 12: stage_smem<12, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 13: stage_smem<13, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 15: stage_smem<15, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        case 16: stage_smem<16, INV>(buf, tw, n, Li, cp, w, W, txp); break;
        default: break;   // the host only builds Rader plans whose n has these radices
    }
}

template <bool INV>
__device__ __forceinline__ void stage_rader(const AxisPlanDev& P, float4* __restrict__ cur, float4* __restrict__ oth,
                                            const float4* __restrict__ rtw, int cp, int w, int W, int txp, bool active)
{
    const int p = P.rader_p, n = P.rader_n, nsub = P.L / p;
    const float invn = 1.0f / (float)n;
    // 1. gather a[m] = x[g^m] into the other buffer
    if (active)
        for (int idx = w; idx < nsub * n; idx += W) {
            const int beta = __float2int_rz(((float)idx + 0.5f) * invn), m = idx - beta * n;
            oth[(beta * p + 1 + m) * txp + cp] = cur[(beta * p + __ldg(P.rader_perm + m)) * txp + cp];
        }
    __syncthreads();
    // 2. n-point forward transforms (DIF, in place, digit-reversed positions)
    {
        int Li = n;
        for (int s = 0; s < P.rader_ns; ++s) {
            const int R = P.rader_radix[s];
            // the butterflies of all blocks are dealt over the workers as ONE sequence (block beta starts where beta-1 ended)
            if (active)
                for (int beta = 0, nb = n / R; beta < nsub; ++beta)
                    rader_substage<false>(R, oth + (size_t)(beta * p + 1) * txp, rtw, n, Li, cp,
                                          ((w - (beta * nb) % W) + W) % W, W, txp);
            Li /= R;
            __syncthreads();
        }
    }
    // 3. multiply by the spectrum of b; DC: X[0] = x[0] + A[0], and x[0] joins the DC term of the product
    if (active) {
        const float2* B = INV ? P.rader_bi : P.rader_bf;
        for (int idx = w; idx < nsub * n; idx += W) {
            const int beta = __float2int_rz(((float)idx + 0.5f) * invn), pp = idx - beta * n;
            float4* slot = oth + (size_t)(beta * p + 1 + pp) * txp + cp;
            const float4 v = *slot;
            const float2 b = __ldg(B + pp);
            float4 o;
            o.x = v.x * b.x - v.z * b.y;
            o.y = v.y * b.x - v.w * b.y;
            o.z = v.x * b.y + v.z * b.x;
            o.w = v.y * b.y + v.w * b.x;
            if (pp == 0) {
                float4* x0p = cur + (size_t)(beta * p) * txp + cp;
                const float4 x0 = *x0p;
                *x0p = make_float4(x0.x + v.x, x0.y + v.y, x0.z + v.z, x0.w + v.w);
                o.x += x0.x;
                o.y += x0.y;
                o.z += x0.z;
                o.w += x0.w;
            }
            *slot = o;
        }
    }
    __syncthreads();
    // 4. n-point inverse transforms (DIT from the digit-reversed positions, unnormalised: 1/n is folded into B)
    {
        int Li = 1;
        for (int s = P.rader_ns - 1; s >= 0; --s) {
            const int R = P.rader_radix[s];
            Li *= R;
            if (active)
                for (int beta = 0, nb = n / R; beta < nsub; ++beta)
                    rader_substage<true>(R, oth + (size_t)(beta * p + 1) * txp, rtw, n, Li, cp,
                                         ((w - (beta * nb) % W) + W) % W, W, txp);
            __syncthreads();
        }
    }
    // 5. scatter X[g^-q] = c[q]
    if (active)
        for (int idx = w; idx < nsub * n; idx += W) {
            const int beta = __float2int_rz(((float)idx + 0.5f) * invn), q = idx - beta * n;
            cur[(beta * p + __ldg(P.rader_iperm + q)) * txp + cp] = oth[(beta * p + 1 + q) * txp + cp];
        }
}

// BIG: the kernel is compiled with the register butterflies of the primes 11..23 (they cost registers, so plans
// without such a radix run kernels compiled without them; AxisPlanDev::big tells the launcher which)
template <bool INV, bool BIG = false>
__device__ __forceinline__ void stage_dispatch(int R, float4*& cur, float4*& oth, const float4* tw, int L, int Li,
                                               int cp, int w, int W, int txp, bool active,
                                               const AxisPlanDev* P = nullptr, const float4* rtw = nullptr)
{
    if constexpr (BIG) {
        if (P != nullptr && P->rader_p == R && Li == R && rtw != nullptr) {   // last radix, S = 1 (uniform branch)
            stage_rader<INV>(*P, cur, oth, rtw, cp, w, W, txp, active);
            return;   // result is in `cur`: no buffer swap
        }
    }
    if (active) {
        if constexpr (BIG) {
            switch (R) {
                case 11: stage_smem<11, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 13: stage_smem<13, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 17: stage_smem<17, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 19: stage_smem<19, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 23: stage_smem<23, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                // fat composite radices of the two-stage plans (fft_butterflies.cuh), for the passes of such a plan that
                // have no compile-time kernel
                case 14: stage_smem<14, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 18: stage_smem<18, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 20: stage_smem<20, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 21: stage_smem<21, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 25: stage_smem<25, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                case 28: stage_smem<28, INV>(cur, tw, L, Li, cp, w, W, txp); return;
                default: break;
            }
        }
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
template <bool INV, bool BIG = false>
__device__ __forceinline__ float4* engine_run(const AxisPlanDev& P, float4* A, float4* B, const float4* tw, int cp,
                                              int w, int W, int txp, bool active, const float4* rtw = nullptr)
{
    float4* cur = A;
    float4* oth = B;
    if (!INV) {
        int Li = P.L;
        for (int s = 0; s < P.ns; ++s) {
            const int R = P.radix[s];
            stage_dispatch<false, BIG>(R, cur, oth, tw, P.L, Li, cp, w, W, txp, active, &P, rtw);
            Li /= R;
            __syncthreads();
        }
    } else {
        int Li = 1;
        for (int s = P.ns - 1; s >= 0; --s) {
            const int R = P.radix[s];
            Li *= R;
            stage_dispatch<true, BIG>(R, cur, oth, tw, P.L, Li, cp, w, W, txp, active, &P, rtw);
            __syncthreads();
        }
    }
    return cur;
}

}  // namespace fcb200
