// Radix butterflies for the sm_100a FFT engine.
//
// All butterflies are forward DFTs (kernel exp(-2*pi*i*k*m/R)) on split real/imag register arrays.
// The inverse transform is obtained for free by swapping the two array arguments:
//   IDFT(x) = swap(DFT(swap(x)))   ->   Dft<R>::run(im, re)
// Part of the replacement for the cuFFT calls of the reference hot path
// (/root/reference/src/convolution3Dfft.cu:519-525, :544-547).
#pragma once
#include <cuda_runtime.h>

namespace fcb200 {

template <int R>
struct Dft;

template <>
struct Dft<2> {
    static __device__ __forceinline__ void run(float* r, float* i)
    {
        float t = r[0] - r[1];
        r[0] = r[0] + r[1];
        r[1] = t;
        t = i[0] - i[1];
        i[0] = i[0] + i[1];
        i[1] = t;
    }
};

template <>
struct Dft<3> {
    static __device__ __forceinline__ void run(float* r, float* i)
    {
        const float S3 = 0.86602540378443864676f;  // sin(2*pi/3)
        float tr = r[1] + r[2], ti = i[1] + i[2];
        float ur = (r[1] - r[2]) * S3, ui = (i[1] - i[2]) * S3;
        float mr = fmaf(-0.5f, tr, r[0]), mi = fmaf(-0.5f, ti, i[0]);
        r[0] += tr;
        i[0] += ti;
        // y1 = m - i*u ; y2 = m + i*u
        r[1] = mr + ui;
        i[1] = mi - ur;
        r[2] = mr - ui;
        i[2] = mi + ur;
    }
};

template <>
struct Dft<4> {
    static __device__ __forceinline__ void run(float* r, float* i)
    {
        float t0r = r[0] + r[2], t0i = i[0] + i[2];
        float t1r = r[0] - r[2], t1i = i[0] - i[2];
        float t2r = r[1] + r[3], t2i = i[1] + i[3];
        float t3r = r[1] - r[3], t3i = i[1] - i[3];
        r[0] = t0r + t2r;
        i[0] = t0i + t2i;
        r[2] = t0r - t2r;
        i[2] = t0i - t2i;
        r[1] = t1r + t3i;  // t1 - i*t3
        i[1] = t1i - t3r;
        r[3] = t1r - t3i;  // t1 + i*t3
        i[3] = t1i + t3r;
    }
};

template <>
struct Dft<5> {
    static __device__ __forceinline__ void run(float* r, float* i)
    {
        const float C1 = 0.30901699437494742410f;   // cos(2*pi/5)
        const float C2 = -0.80901699437494742410f;  // cos(4*pi/5)
        const float S1 = 0.95105651629515357212f;   // sin(2*pi/5)
        const float S2 = 0.58778525229247312917f;   // sin(4*pi/5)
        float t1r = r[1] + r[4], t1i = i[1] + i[4];
        float t2r = r[2] + r[3], t2i = i[2] + i[3];
        float t3r = r[1] - r[4], t3i = i[1] - i[4];
        float t4r = r[2] - r[3], t4i = i[2] - i[3];
        float a1r = fmaf(C2, t2r, fmaf(C1, t1r, r[0])), a1i = fmaf(C2, t2i, fmaf(C1, t1i, i[0]));
        float a2r = fmaf(C1, t2r, fmaf(C2, t1r, r[0])), a2i = fmaf(C1, t2i, fmaf(C2, t1i, i[0]));
        float b1r = fmaf(S2, t4r, S1 * t3r), b1i = fmaf(S2, t4i, S1 * t3i);
        float b2r = fmaf(-S1, t4r, S2 * t3r), b2i = fmaf(-S1, t4i, S2 * t3i);
        r[0] = r[0] + t1r + t2r;
        i[0] = i[0] + t1i + t2i;
        // y1 = a1 - i*b1 ; y4 = a1 + i*b1 ; y2 = a2 - i*b2 ; y3 = a2 + i*b2
        r[1] = a1r + b1i;
        i[1] = a1i - b1r;
        r[4] = a1r - b1i;
        i[4] = a1i + b1r;
        r[2] = a2r + b2i;
        i[2] = a2i - b2r;
        r[3] = a2r - b2i;
        i[3] = a2i + b2r;
    }
};

template <>
struct Dft<7> {
    static __device__ __forceinline__ void run(float* r, float* i)
    {
        const float C1 = 0.62348980185873353053f;   // cos(2*pi/7)
        const float C2 = -0.22252093395631440429f;  // cos(4*pi/7)
        const float C3 = -0.90096886790241912624f;  // cos(6*pi/7)
        const float S1 = 0.78183148246802980871f;   // sin(2*pi/7)
        const float S2 = 0.97492791218182360702f;   // sin(4*pi/7)
        const float S3 = 0.43388373911755812048f;   // sin(6*pi/7)
        float t1r = r[1] + r[6], t1i = i[1] + i[6], u1r = r[1] - r[6], u1i = i[1] - i[6];
        float t2r = r[2] + r[5], t2i = i[2] + i[5], u2r = r[2] - r[5], u2i = i[2] - i[5];
        float t3r = r[3] + r[4], t3i = i[3] + i[4], u3r = r[3] - r[4], u3i = i[3] - i[4];
        // m=1: cos idx (1,2,3) sin idx (1,2,3); m=2: (2,4->3,6->1) sin(2, 4->-3, 6->-1); m=3: (3,6->1,9->2) sin(3,-1,2)
        float a1r = fmaf(C3, t3r, fmaf(C2, t2r, fmaf(C1, t1r, r[0])));
        float a1i = fmaf(C3, t3i, fmaf(C2, t2i, fmaf(C1, t1i, i[0])));
        float a2r = fmaf(C1, t3r, fmaf(C3, t2r, fmaf(C2, t1r, r[0])));
        float a2i = fmaf(C1, t3i, fmaf(C3, t2i, fmaf(C2, t1i, i[0])));
        float a3r = fmaf(C2, t3r, fmaf(C1, t2r, fmaf(C3, t1r, r[0])));
        float a3i = fmaf(C2, t3i, fmaf(C1, t2i, fmaf(C3, t1i, i[0])));
        float b1r = fmaf(S3, u3r, fmaf(S2, u2r, S1 * u1r)), b1i = fmaf(S3, u3i, fmaf(S2, u2i, S1 * u1i));
        float b2r = fmaf(-S1, u3r, fmaf(-S3, u2r, S2 * u1r)), b2i = fmaf(-S1, u3i, fmaf(-S3, u2i, S2 * u1i));
        float b3r = fmaf(S2, u3r, fmaf(-S1, u2r, S3 * u1r)), b3i = fmaf(S2, u3i, fmaf(-S1, u2i, S3 * u1i));
        r[0] = r[0] + t1r + t2r + t3r;
        i[0] = i[0] + t1i + t2i + t3i;
        r[1] = a1r + b1i;
        i[1] = a1i - b1r;
        r[6] = a1r - b1i;
        i[6] = a1i + b1r;
        r[2] = a2r + b2i;
        i[2] = a2i - b2r;
        r[5] = a2r - b2i;
        i[5] = a2i + b2r;
        r[3] = a3r + b3i;
        i[3] = a3i - b3r;
        r[4] = a3r - b3i;
        i[4] = a3i + b3r;
    }
};

template <>
struct Dft<8> {
    static __device__ __forceinline__ void run(float* r, float* i)
    {
        const float C = 0.70710678118654752440f;
        float er[4] = {r[0], r[2], r[4], r[6]}, ei[4] = {i[0], i[2], i[4], i[6]};
        float qr[4] = {r[1], r[3], r[5], r[7]}, qi[4] = {i[1], i[3], i[5], i[7]};
        Dft<4>::run(er, ei);
        Dft<4>::run(qr, qi);
        // odd outputs times w8^k
        float o1r = C * (qr[1] + qi[1]), o1i = C * (qi[1] - qr[1]);
        float o2r = qi[2], o2i = -qr[2];
        float o3r = C * (qi[3] - qr[3]), o3i = -C * (qr[3] + qi[3]);
        r[0] = er[0] + qr[0];
        i[0] = ei[0] + qi[0];
        r[4] = er[0] - qr[0];
        i[4] = ei[0] - qi[0];
        r[1] = er[1] + o1r;
        i[1] = ei[1] + o1i;
        r[5] = er[1] - o1r;
        i[5] = ei[1] - o1i;
        r[2] = er[2] + o2r;
        i[2] = ei[2] + o2i;
        r[6] = er[2] - o2r;
        i[6] = ei[2] - o2i;
        r[3] = er[3] + o3r;
        i[3] = ei[3] + o3i;
        r[7] = er[3] - o3r;
        i[7] = ei[3] - o3i;
    }
};

// (xr + i*xi) *= (c + i*s)
__device__ __forceinline__ void cmul(float& xr, float& xi, float c, float s)
{
    float t = xr * c - xi * s;
    xi = fmaf(xr, s, xi * c);
    xr = t;
}
// (xr + i*xi) *= conj(c + i*s)
__device__ __forceinline__ void cmulc(float& xr, float& xi, float c, float s)
{
    float t = fmaf(xi, s, xr * c);
    xi = xi * c - xr * s;
    xr = t;
}

}  // namespace fcb200
