// Radix butterflies for the sm_100a FFT engine, on PACKED pairs of pencils.
//
// Every value is a float2 holding the same quantity of two adjacent pencils (lo = pencil 0,
// hi = pencil 1); real and imaginary parts live in separate arrays.  All arithmetic uses Blackwell's
// packed fp32 instructions (FADD2 / FMUL2 / FFMA2, PTX add/mul/fma.f32x2, sm_100+), so one issued
// instruction advances two pencils.  The spectrum is therefore stored "pair-planar": 16 bytes =
// (re0, re1, im0, im1) -- see fc_common.h.
//
// All butterflies are forward DFTs (kernel exp(-2*pi*i*k*m/R)).  The inverse transform is obtained
// for free by swapping the two array arguments:  IDFT(x) = swap(DFT(swap(x)))  ->  Dft<R>::run(im, re).
// Part of the replacement for the cuFFT calls of the reference hot path
// (/root/reference/src/convolution3Dfft.cu:519-525, :544-547).
#pragma once
#include <cuda_runtime.h>

namespace fcb200 {

typedef float2 p2;

__device__ __forceinline__ p2 padd(p2 a, p2 b) { return __fadd2_rn(a, b); }
__device__ __forceinline__ p2 pneg(p2 a) { return make_float2(-a.x, -a.y); }      // folds into operand modifiers
__device__ __forceinline__ p2 psub(p2 a, p2 b) { return __fadd2_rn(a, pneg(b)); }
__device__ __forceinline__ p2 pmul(p2 a, p2 b) { return __fmul2_rn(a, b); }
__device__ __forceinline__ p2 pmuls(p2 a, float c) { return __fmul2_rn(a, make_float2(c, c)); }
__device__ __forceinline__ p2 pfma(p2 a, p2 b, p2 c) { return __ffma2_rn(a, b, c); }           // a*b + c
__device__ __forceinline__ p2 pfmas(p2 a, float s, p2 c) { return __ffma2_rn(a, make_float2(s, s), c); }

template <int R>
struct Dft;

template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        p2 t = psub(r[0], r[1]);
        r[0] = padd(r[0], r[1]);
        r[1] = t;
        t = psub(i[0], i[1]);
        i[0] = padd(i[0], i[1]);
        i[1] = t;
    }
};

template <>
struct Dft<3> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        const float S3 = 0.86602540378443864676f;  // sin(2*pi/3)
        p2 tr = padd(r[1], r[2]), ti = padd(i[1], i[2]);
        p2 ur = pmuls(psub(r[1], r[2]), S3), ui = pmuls(psub(i[1], i[2]), S3);
        p2 mr = pfmas(tr, -0.5f, r[0]), mi = pfmas(ti, -0.5f, i[0]);
        r[0] = padd(r[0], tr);
        i[0] = padd(i[0], ti);
        // y1 = m - i*u ; y2 = m + i*u
        r[1] = padd(mr, ui);
        i[1] = psub(mi, ur);
        r[2] = psub(mr, ui);
        i[2] = padd(mi, ur);
    }
};

template <>
struct Dft<4> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        p2 t0r = padd(r[0], r[2]), t0i = padd(i[0], i[2]);
        p2 t1r = psub(r[0], r[2]), t1i = psub(i[0], i[2]);
        p2 t2r = padd(r[1], r[3]), t2i = padd(i[1], i[3]);
        p2 t3r = psub(r[1], r[3]), t3i = psub(i[1], i[3]);
        r[0] = padd(t0r, t2r);
        i[0] = padd(t0i, t2i);
        r[2] = psub(t0r, t2r);
        i[2] = psub(t0i, t2i);
        r[1] = padd(t1r, t3i);  // t1 - i*t3
        i[1] = psub(t1i, t3r);
        r[3] = psub(t1r, t3i);  // t1 + i*t3
        i[3] = padd(t1i, t3r);
    }
};

template <>
struct Dft<5> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        const float C1 = 0.30901699437494742410f;   // cos(2*pi/5)
        const float C2 = -0.80901699437494742410f;  // cos(4*pi/5)
        const float S1 = 0.95105651629515357212f;   // sin(2*pi/5)
        const float S2 = 0.58778525229247312917f;   // sin(4*pi/5)
        p2 t1r = padd(r[1], r[4]), t1i = padd(i[1], i[4]);
        p2 t2r = padd(r[2], r[3]), t2i = padd(i[2], i[3]);
        p2 t3r = psub(r[1], r[4]), t3i = psub(i[1], i[4]);
        p2 t4r = psub(r[2], r[3]), t4i = psub(i[2], i[3]);
        p2 a1r = pfmas(t2r, C2, pfmas(t1r, C1, r[0])), a1i = pfmas(t2i, C2, pfmas(t1i, C1, i[0]));
        p2 a2r = pfmas(t2r, C1, pfmas(t1r, C2, r[0])), a2i = pfmas(t2i, C1, pfmas(t1i, C2, i[0]));
        p2 b1r = pfmas(t4r, S2, pmuls(t3r, S1)), b1i = pfmas(t4i, S2, pmuls(t3i, S1));
        p2 b2r = pfmas(t4r, -S1, pmuls(t3r, S2)), b2i = pfmas(t4i, -S1, pmuls(t3i, S2));
        r[0] = padd(padd(r[0], t1r), t2r);
        i[0] = padd(padd(i[0], t1i), t2i);
        // y1 = a1 - i*b1 ; y4 = a1 + i*b1 ; y2 = a2 - i*b2 ; y3 = a2 + i*b2
        r[1] = padd(a1r, b1i);
        i[1] = psub(a1i, b1r);
        r[4] = psub(a1r, b1i);
        i[4] = padd(a1i, b1r);
        r[2] = padd(a2r, b2i);
        i[2] = psub(a2i, b2r);
        r[3] = psub(a2r, b2i);
        i[3] = padd(a2i, b2r);
    }
};

template <>
struct Dft<7> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        const float C1 = 0.62348980185873353053f;   // cos(2*pi/7)
        const float C2 = -0.22252093395631440429f;  // cos(4*pi/7)
        const float C3 = -0.90096886790241912624f;  // cos(6*pi/7)
        const float S1 = 0.78183148246802980871f;   // sin(2*pi/7)
        const float S2 = 0.97492791218182360702f;   // sin(4*pi/7)
        const float S3 = 0.43388373911755812048f;   // sin(6*pi/7)
        p2 t1r = padd(r[1], r[6]), t1i = padd(i[1], i[6]), u1r = psub(r[1], r[6]), u1i = psub(i[1], i[6]);
        p2 t2r = padd(r[2], r[5]), t2i = padd(i[2], i[5]), u2r = psub(r[2], r[5]), u2i = psub(i[2], i[5]);
        p2 t3r = padd(r[3], r[4]), t3i = padd(i[3], i[4]), u3r = psub(r[3], r[4]), u3i = psub(i[3], i[4]);
        // cos index (m*k mod 7 folded to 1..3), sin index with sign: m=1: (1,2,3); m=2: (2,-3,-1); m=3: (3,-1,2)
        p2 a1r = pfmas(t3r, C3, pfmas(t2r, C2, pfmas(t1r, C1, r[0])));
        p2 a1i = pfmas(t3i, C3, pfmas(t2i, C2, pfmas(t1i, C1, i[0])));
        p2 a2r = pfmas(t3r, C1, pfmas(t2r, C3, pfmas(t1r, C2, r[0])));
        p2 a2i = pfmas(t3i, C1, pfmas(t2i, C3, pfmas(t1i, C2, i[0])));
        p2 a3r = pfmas(t3r, C2, pfmas(t2r, C1, pfmas(t1r, C3, r[0])));
        p2 a3i = pfmas(t3i, C2, pfmas(t2i, C1, pfmas(t1i, C3, i[0])));
        p2 b1r = pfmas(u3r, S3, pfmas(u2r, S2, pmuls(u1r, S1))), b1i = pfmas(u3i, S3, pfmas(u2i, S2, pmuls(u1i, S1)));
        p2 b2r = pfmas(u3r, -S1, pfmas(u2r, -S3, pmuls(u1r, S2))), b2i = pfmas(u3i, -S1, pfmas(u2i, -S3, pmuls(u1i, S2)));
        p2 b3r = pfmas(u3r, S2, pfmas(u2r, -S1, pmuls(u1r, S3))), b3i = pfmas(u3i, S2, pfmas(u2i, -S1, pmuls(u1i, S3)));
        r[0] = padd(padd(r[0], t1r), padd(t2r, t3r));
        i[0] = padd(padd(i[0], t1i), padd(t2i, t3i));
        r[1] = padd(a1r, b1i);
        i[1] = psub(a1i, b1r);
        r[6] = psub(a1r, b1i);
        i[6] = padd(a1i, b1r);
        r[2] = padd(a2r, b2i);
        i[2] = psub(a2i, b2r);
        r[5] = psub(a2r, b2i);
        i[5] = padd(a2i, b2r);
        r[3] = padd(a3r, b3i);
        i[3] = psub(a3i, b3r);
        r[4] = psub(a3r, b3i);
        i[4] = padd(a3i, b3r);
    }
};

template <>
struct Dft<8> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        const float C = 0.70710678118654752440f;
        p2 er[4] = {r[0], r[2], r[4], r[6]}, ei[4] = {i[0], i[2], i[4], i[6]};
        p2 qr[4] = {r[1], r[3], r[5], r[7]}, qi[4] = {i[1], i[3], i[5], i[7]};
        Dft<4>::run(er, ei);
        Dft<4>::run(qr, qi);
        // odd outputs times w8^k
        p2 o1r = pmuls(padd(qr[1], qi[1]), C), o1i = pmuls(psub(qi[1], qr[1]), C);
        p2 o2r = qi[2], o2i = pneg(qr[2]);
        p2 o3r = pmuls(psub(qi[3], qr[3]), C), o3i = pmuls(padd(qr[3], qi[3]), -C);
        r[0] = padd(er[0], qr[0]);
        i[0] = padd(ei[0], qi[0]);
        r[4] = psub(er[0], qr[0]);
        i[4] = psub(ei[0], qi[0]);
        r[1] = padd(er[1], o1r);
        i[1] = padd(ei[1], o1i);
        r[5] = psub(er[1], o1r);
        i[5] = psub(ei[1], o1i);
        r[2] = padd(er[2], o2r);
        i[2] = padd(ei[2], o2i);
        r[6] = psub(er[2], o2r);
        i[6] = psub(ei[2], o2i);
        r[3] = padd(er[3], o3r);
        i[3] = padd(ei[3], o3i);
        r[7] = psub(er[3], o3r);
        i[7] = psub(ei[3], o3i);
    }
};


// Radix 16 = 4 x 4 with the internal twiddles w16^(b*c) folded in as constants.
//   n = 4a + b, k = c + 4d:  X[c + 4d] = sum_b w4^(b d) [ w16^(b c) sum_a x[4a + b] w4^(a c) ]
template <>
struct Dft<16> {
    // (xr + i xi) *= (c - i s)   with c, s > 0 constants (forward roots lie in the lower half plane)
    static __device__ __forceinline__ void rot(p2& xr, p2& xi, float c, float s)
    {
        const p2 nr = pfmas(xi, s, pmuls(xr, c));
        xi = pfmas(xr, -s, pmuls(xi, c));
        xr = nr;
    }
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        const float C1 = 0.92387953251128675613f;  // cos(pi/8)
        const float S1 = 0.38268343236508977173f;  // sin(pi/8)
        const float C2 = 0.70710678118654752440f;  // cos(pi/4)
        p2 tr[16], ti[16];
#pragma unroll
        for (int b = 0; b < 4; ++b) {
            p2 ar[4] = {r[b], r[4 + b], r[8 + b], r[12 + b]};
            p2 ai[4] = {i[b], i[4 + b], i[8 + b], i[12 + b]};
            Dft<4>::run(ar, ai);
#pragma unroll
            for (int c = 0; c < 4; ++c) {
                tr[4 * c + b] = ar[c];
                ti[4 * c + b] = ai[c];
            }
        }
        // t_b[c] *= w16^(b c), index 4c + b
        rot(tr[4 * 1 + 1], ti[4 * 1 + 1], C1, S1);                 // w16^1
        rot(tr[4 * 1 + 2], ti[4 * 1 + 2], C2, C2);                 // w16^2
        rot(tr[4 * 1 + 3], ti[4 * 1 + 3], S1, C1);                 // w16^3
        rot(tr[4 * 2 + 1], ti[4 * 2 + 1], C2, C2);                 // w16^2
        {                                                          // w16^4 = -i
            const p2 t = tr[4 * 2 + 2];
            tr[4 * 2 + 2] = ti[4 * 2 + 2];
            ti[4 * 2 + 2] = pneg(t);
        }
        rot(tr[4 * 2 + 3], ti[4 * 2 + 3], -C2, C2);                // w16^6 = (-C2, -C2)
        rot(tr[4 * 3 + 1], ti[4 * 3 + 1], S1, C1);                 // w16^3
        rot(tr[4 * 3 + 2], ti[4 * 3 + 2], -C2, C2);                // w16^6
        rot(tr[4 * 3 + 3], ti[4 * 3 + 3], -C1, -S1);               // w16^9 = (-C1, +S1)
#pragma unroll
        for (int c = 0; c < 4; ++c) {
            p2 ar[4] = {tr[4 * c], tr[4 * c + 1], tr[4 * c + 2], tr[4 * c + 3]};
            p2 ai[4] = {ti[4 * c], ti[4 * c + 1], ti[4 * c + 2], ti[4 * c + 3]};
            Dft<4>::run(ar, ai);
#pragma unroll
            for (int d = 0; d < 4; ++d) {
                r[c + 4 * d] = ar[d];
                i[c + 4 * d] = ai[d];
            }
        }
    }
};

// Radix 24 = 8 x 3 (x rows of 384 voxels: M = 192 = 24 * 8), internal twiddles w24^(b*c) as constants.
//   n = 3a + b, k = c + 8d:  X[c + 8d] = sum_b w3^(b d) [ w24^(b c) sum_a x[3a + b] w8^(a c) ]
template <>
struct Dft<24> {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        const float C1 = 0.96592582628906828675f;  // cos(pi/12)
        const float S1 = 0.25881904510252076235f;  // sin(pi/12)
        const float C2 = 0.86602540378443864676f;  // cos(pi/6)
        const float C3 = 0.70710678118654752440f;  // cos(pi/4)
        p2 tr[24], ti[24];   // t_b[c] at index 3c + b
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            p2 ar[8], ai[8];
#pragma unroll
            for (int a = 0; a < 8; ++a) {
                ar[a] = r[3 * a + b];
                ai[a] = i[3 * a + b];
            }
            Dft<8>::run(ar, ai);
#pragma unroll
            for (int c = 0; c < 8; ++c) {
                tr[3 * c + b] = ar[c];
                ti[3 * c + b] = ai[c];
            }
        }
        // b = 1: w24^c, c = 1..7
        Dft<16>::rot(tr[3 * 1 + 1], ti[3 * 1 + 1], C1, S1);        // w24^1
        Dft<16>::rot(tr[3 * 2 + 1], ti[3 * 2 + 1], C2, 0.5f);      // w24^2
        Dft<16>::rot(tr[3 * 3 + 1], ti[3 * 3 + 1], C3, C3);        // w24^3
        Dft<16>::rot(tr[3 * 4 + 1], ti[3 * 4 + 1], 0.5f, C2);      // w24^4
        Dft<16>::rot(tr[3 * 5 + 1], ti[3 * 5 + 1], S1, C1);        // w24^5
        {                                                          // w24^6 = -i
            const p2 t = tr[3 * 6 + 1];
            tr[3 * 6 + 1] = ti[3 * 6 + 1];
            ti[3 * 6 + 1] = pneg(t);
        }
        Dft<16>::rot(tr[3 * 7 + 1], ti[3 * 7 + 1], -S1, C1);       // w24^7
        // b = 2: w24^(2c), c = 1..7
        Dft<16>::rot(tr[3 * 1 + 2], ti[3 * 1 + 2], C2, 0.5f);      // w24^2
        Dft<16>::rot(tr[3 * 2 + 2], ti[3 * 2 + 2], 0.5f, C2);      // w24^4
        {                                                          // w24^6 = -i
            const p2 t = tr[3 * 3 + 2];
            tr[3 * 3 + 2] = ti[3 * 3 + 2];
            ti[3 * 3 + 2] = pneg(t);
        }
        Dft<16>::rot(tr[3 * 4 + 2], ti[3 * 4 + 2], -0.5f, C2);     // w24^8
        Dft<16>::rot(tr[3 * 5 + 2], ti[3 * 5 + 2], -C2, 0.5f);     // w24^10
        tr[3 * 6 + 2] = pneg(tr[3 * 6 + 2]);                       // w24^12 = -1
        ti[3 * 6 + 2] = pneg(ti[3 * 6 + 2]);
        Dft<16>::rot(tr[3 * 7 + 2], ti[3 * 7 + 2], -C2, -0.5f);    // w24^14
#pragma unroll
        for (int c = 0; c < 8; ++c) {
            Dft<3>::run(tr + 3 * c, ti + 3 * c);
#pragma unroll
            for (int d = 0; d < 3; ++d) {
                r[c + 8 * d] = tr[3 * c + d];
                i[c + 8 * d] = ti[3 * c + d];
            }
        }
    }
};

// ---- composite radices 6, 9, 10, 12, 15 ------------------------------------------------------------------
// The 7-smooth extents callers pad to (270, 300, 420, 1080, 1125, 2160 ...) have many small odd factors; one
// register stage per prime would mean four or five shared-memory round trips.  A composite radix A*B runs as
// Cooley-Tukey inside the registers of one thread: A-point DFTs, constant twiddles w_N^(b c), B-point DFTs.
//   n = B a + b, k = c + A d:  X[c + A d] = sum_b w_B^(b d) [ w_N^(b c) sum_a x[B a + b] w_A^(a c) ]
// The roots are compile-time constants (Taylor series evaluated in double by the compiler).
namespace ctrig {
constexpr double kPi = 3.14159265358979323846264338327950288;
constexpr double ccos(double x)
{
    double t = 1.0, s = 1.0;
    for (int k = 1; k <= 16; ++k) {
        t *= -x * x / ((2.0 * k - 1.0) * (2.0 * k));
        s += t;
    }
    return s;
}
constexpr double csin(double x)
{
    double t = x, s = x;
    for (int k = 1; k <= 16; ++k) {
        t *= -x * x / ((2.0 * k) * (2.0 * k + 1.0));
        s += t;
    }
    return s;
}
// c[m] - i s[m] = exp(-2 pi i m / N), angle reduced to (-pi, pi]
template <int N>
struct Roots {
    float c[N], s[N];
    constexpr Roots() : c{}, s{}
    {
        for (int m = 0; m < N; ++m) {
            const int mm = (2 * m > N) ? m - N : m;
            const double a = 2.0 * kPi * mm / N;
            c[m] = (float)ccos(a);
            s[m] = (float)csin(a);
        }
    }
};
}  // namespace ctrig

template <int A, int B>
struct DftCT {
    static constexpr int N = A * B;
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        constexpr ctrig::Roots<N> W{};
        p2 tr[N], ti[N];   // t_b[c] at index B c + b
#pragma unroll
        for (int b = 0; b < B; ++b) {
            p2 ar[A], ai[A];
#pragma unroll
            for (int a = 0; a < A; ++a) {
                ar[a] = r[B * a + b];
                ai[a] = i[B * a + b];
            }
            Dft<A>::run(ar, ai);
#pragma unroll
            for (int c = 0; c < A; ++c) {
                if (b * c != 0) {   // (t_r + i t_i) *= (C - i S)
                    const float C = W.c[(b * c) % N], S = W.s[(b * c) % N];
                    const p2 nr = pfmas(ai[c], S, pmuls(ar[c], C));
                    ai[c] = pfmas(ar[c], -S, pmuls(ai[c], C));
                    ar[c] = nr;
                }
                tr[B * c + b] = ar[c];
                ti[B * c + b] = ai[c];
            }
        }
#pragma unroll
        for (int c = 0; c < A; ++c) {
            Dft<B>::run(tr + B * c, ti + B * c);
#pragma unroll
            for (int d = 0; d < B; ++d) {
                r[c + A * d] = tr[B * c + d];
                i[c + A * d] = ti[B * c + d];
            }
        }
    }
};
template <> struct Dft<6> : DftCT<2, 3> {};
template <> struct Dft<9> : DftCT<3, 3> {};
template <> struct Dft<10> : DftCT<2, 5> {};
template <> struct Dft<12> : DftCT<4, 3> {};
template <> struct Dft<15> : DftCT<3, 5> {};
// fat radices of the two-stage plans (300 = 20 * 15, 420 = 20 * 21, 280 = 20 * 14 ...): one shared-memory round trip
// instead of three for lengths with four small prime factors
template <> struct Dft<14> : DftCT<2, 7> {};
template <> struct Dft<18> : DftCT<2, 9> {};
template <> struct Dft<20> : DftCT<4, 5> {};
template <> struct Dft<21> : DftCT<3, 7> {};
template <> struct Dft<25> : DftCT<5, 5> {};
template <> struct Dft<28> : DftCT<4, 7> {};

// ---- odd primes 11, 13, 17, 19, 23 in registers ----------------------------------------------------------------
// The extents reference callers pad to by image + kernel - 1 are rarely 7-smooth (its own tests use 130 = 2*5*13,
// 132 = 4*3*11, 46 = 2*23, 66 = 2*3*11 ...: /root/reference/tests/test_gpu_numerical_stability.cpp).  A prime radix R runs
// as the symmetric direct DFT in the registers of one thread, (R-1)^2 / 2 packed multiply-adds on compile-time roots:
//   t_j = x_j + x_(R-j),  u_j = x_j - x_(R-j),  j = 1 .. (R-1)/2
//   X_m, X_(R-m) = (x_0 + sum_j cos(2 pi j m / R) t_j)  -/+  i (sum_j sin(2 pi j m / R) u_j)
// instead of an O(R) sum per OUTPUT through shared memory (the generic stage, still used for larger primes).
template <int R>
struct DftOddPrime {
    static __device__ __forceinline__ void run(p2* r, p2* i)
    {
        constexpr ctrig::Roots<R> W{};
        constexpr int H = (R - 1) / 2;
        p2 tr[H], ti[H], ur[H], ui[H];
#pragma unroll
        for (int j = 1; j <= H; ++j) {
            tr[j - 1] = padd(r[j], r[R - j]);
            ti[j - 1] = padd(i[j], i[R - j]);
            ur[j - 1] = psub(r[j], r[R - j]);
            ui[j - 1] = psub(i[j], i[R - j]);
        }
        const p2 x0r = r[0], x0i = i[0];
        p2 sr = x0r, si = x0i;
#pragma unroll
        for (int j = 0; j < H; ++j) {
            sr = padd(sr, tr[j]);
            si = padd(si, ti[j]);
        }
        r[0] = sr;
        i[0] = si;
#pragma unroll
        for (int m = 1; m <= H; ++m) {
            p2 ar = x0r, ai = x0i;
            p2 br = make_float2(0.f, 0.f), bi = make_float2(0.f, 0.f);
#pragma unroll
            for (int j = 1; j <= H; ++j) {
                const float C = W.c[(j * m) % R], S = W.s[(j * m) % R];
                ar = pfmas(tr[j - 1], C, ar);
                ai = pfmas(ti[j - 1], C, ai);
                br = pfmas(ur[j - 1], S, br);
                bi = pfmas(ui[j - 1], S, bi);
            }
            // X_m = a - i b,  X_(R-m) = a + i b
            r[m] = padd(ar, bi);
            i[m] = psub(ai, br);
            r[R - m] = psub(ar, bi);
            i[R - m] = padd(ai, br);
        }
    }
};
template <> struct Dft<11> : DftOddPrime<11> {};
template <> struct Dft<13> : DftOddPrime<13> {};
template <> struct Dft<17> : DftOddPrime<17> {};
template <> struct Dft<19> : DftOddPrime<19> {};
template <> struct Dft<23> : DftOddPrime<23> {};

// Twiddles are kept in shared memory as float4 (c, c, s, s): both packed operands come out of one
// 128-bit load as aligned register pairs.
// (xr + i*xi) *= (c + i*s)
__device__ __forceinline__ void cmul(p2& xr, p2& xi, const float4& t)
{
    const p2 C = make_float2(t.x, t.y), S = make_float2(t.z, t.w);
    const p2 nr = pfma(xr, C, pneg(pmul(xi, S)));
    xi = pfma(xr, S, pmul(xi, C));
    xr = nr;
}
// (xr + i*xi) *= conj(c + i*s)
__device__ __forceinline__ void cmulc(p2& xr, p2& xi, const float4& t)
{
    const p2 C = make_float2(t.x, t.y), S = make_float2(t.z, t.w);
    const p2 nr = pfma(xi, S, pmul(xr, C));
    xi = pfma(xi, C, pneg(pmul(xr, S)));
    xr = nr;
}

}  // namespace fcb200
