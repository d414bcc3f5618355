// sm_100a kernels of the 3D FFT convolution hot path and their launchers.
//
//   x_fwd_kernel   real rows -> half spectrum rows (R2C along x); image loader or PSF gather loader
//                  (PSF zero-pad + circular shift of /root/reference/src/convolution3Dfft.cu:128-166
//                  fused into the load, never materialising the padded PSF)
//   col_kernel     strided-axis complex passes (y and z): forward, inverse, or fused
//                  forward -> x H * 1/N -> inverse (modulateAndNormalize_kernel of
//                  /root/reference/src/convolution3Dfft.cu:41-62 fused between the two z transforms)
//   x_inv_kernel   half spectrum rows -> real rows (C2R along x)
//
// Together they replace cufftExecR2C x2 + modulateAndNormalize_kernel + cufftExecC2R
// (/root/reference/src/convolution3Dfft.cu:519-547) and the row re-layout loops (:474-486, :495-510,
// :561-575), which do not exist here: the X passes read and write dense rows directly.
#include "fft_engine.cuh"
#include "fft_kernels.h"

namespace fcb200 {

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

// float2 slot of (local row, position) in the swizzled X-pass tile (16 float2 per tile row)
__device__ __forceinline__ int xslot(int lrow, int pos)
{
    return pos * 16 + ((((lrow >> 1) ^ swz8(pos)) << 1) | (lrow & 1));
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

__device__ __forceinline__ void load_twiddles(float2* tw_s, const float2* tw_g, int L)
{
    for (int i = threadIdx.x; i < L; i += blockDim.x) tw_s[i] = __ldg(tw_g + i);
}

// ------------------------------------------------------------------------------------------------
// X forward: R2C along x for 16 rows per CTA
// ------------------------------------------------------------------------------------------------
template <int LOADER>  // 0: dense real rows, 1: PSF gather
__global__ void __launch_bounds__(kColThreads) x_fwd_kernel(XArgs a)
{
    extern __shared__ float4 smem[];
    const Geometry g = a.g;
    const int L = a.P.L;
    const int tile_rows = g.odd ? L : (g.M + 1);
    float4* A = smem;
    float4* B = a.P.generic ? (A + (size_t)tile_rows * 8) : nullptr;
    float2* tw_s = reinterpret_cast<float2*>(A + (size_t)tile_rows * 8 * (a.P.generic ? 2 : 1));

    const int t = threadIdx.x;
    const int cp = t & 7, w = t >> 3, W = blockDim.x >> 3;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    const int pl = lane & 7, rr = (lane >> 3) & 1, ph = lane >> 4;
    const long long row0 = (long long)blockIdx.x * 16;

    load_twiddles(tw_s, a.P.tw, L);

    // ---- transposing load: global rows -> tile[position][row]
    {
        float2* A2 = reinterpret_cast<float2*>(A);
        const int npos = L;
        const int nchunks = (npos + 15) >> 4;
        for (int u = warp; u < 8 * nchunks; u += nwarps) {
            const int rp = u / nchunks, c = u - rp * nchunks;
            const int pos = c * 16 + ph * 8 + pl;
            const int lrow = 2 * rp + rr;
            const long long li = row0 + lrow;
            const long long grow = (li < a.nrows) ? (a.rowList ? (long long)a.rowList[li] : li) : -1;
            float2 v = make_float2(0.f, 0.f);
            if (LOADER == 0 && !g.odd && pos < npos && grow >= 0) {
                // asynchronous 8-byte copies: every thread keeps all its loads in flight, no registers
                cp_async8(&A2[xslot(lrow, pos)], reinterpret_cast<const float2*>(a.in_real + grow * g.nx) + pos);
                continue;
            }
            if (pos < npos && grow >= 0) {
                if (LOADER == 0) {
                    v.x = __ldg(a.in_real + grow * g.nx + pos);
                } else {
                    if (g.odd) v.x = psf_tap(a.psf, grow * g.nx + pos);
                    else {
                        v.x = psf_tap(a.psf, grow * g.nx + 2 * pos);
                        v.y = psf_tap(a.psf, grow * g.nx + 2 * pos + 1);
                    }
                }
            }
            if (pos < npos) A2[xslot(lrow, pos)] = v;
        }
    }
    cp_async_wait_all();
    __syncthreads();

    float4* cur = engine_run<false, true>(a.P, A, B, tw_s, cp, w, W, 8, true);

    // ---- split the packed transform into the spectrum of the real rows (even nx)
    if (!g.odd) {
        const int M = g.M;
        for (int k = w; k <= M / 2; k += W) {
            if (k == 0) {
                const int p0 = __ldg(a.P.pos);
                float4 v = cur[tile_idx<true>(p0, cp, 8)];
                cur[tile_idx<true>(p0, cp, 8)] = make_float4(v.x + v.y, 0.f, v.z + v.w, 0.f);
                cur[tile_idx<true>(M, cp, 8)] = make_float4(v.x - v.y, 0.f, v.z - v.w, 0.f);
            } else {
                const int k2 = M - k;
                const int pk = __ldg(a.P.pos + k), pk2 = __ldg(a.P.pos + k2);
                const float4 va = cur[tile_idx<true>(pk, cp, 8)];
                const float4 vb = cur[tile_idx<true>(pk2, cp, 8)];
                const float2 tk = __ldg(a.twx + k);  // exp(-2*pi*i*k/nx)
                float4 o1, o2;
                {
                    float er = 0.5f * (va.x + vb.x), ei = 0.5f * (va.y - vb.y);
                    float orr = 0.5f * (va.y + vb.y), oi = -0.5f * (va.x - vb.x);
                    float wr = tk.x * orr - tk.y * oi, wi = tk.x * oi + tk.y * orr;
                    o1.x = er + wr;
                    o1.y = ei + wi;
                    o2.x = er - wr;
                    o2.y = -(ei - wi);
                }
                {
                    float er = 0.5f * (va.z + vb.z), ei = 0.5f * (va.w - vb.w);
                    float orr = 0.5f * (va.w + vb.w), oi = -0.5f * (va.z - vb.z);
                    float wr = tk.x * orr - tk.y * oi, wi = tk.x * oi + tk.y * orr;
                    o1.z = er + wr;
                    o1.w = ei + wi;
                    o2.z = er - wr;
                    o2.w = -(ei - wi);
                }
                cur[tile_idx<true>(pk, cp, 8)] = o1;
                cur[tile_idx<true>(pk2, cp, 8)] = o2;
            }
        }
        __syncthreads();
    }

    // ---- transposing store: tile -> spectrum rows (pad columns are written as zeros)
    {
        const float2* C2 = reinterpret_cast<const float2*>(cur);
        const int nchunks = (g.xcp + 15) >> 4;
        for (int u = warp; u < 8 * nchunks; u += nwarps) {
            const int rp = u / nchunks, c = u - rp * nchunks;
            const int pos = c * 16 + ph * 8 + pl;
            const int lrow = 2 * rp + rr;
            const long long li = row0 + lrow;
            const long long grow = (li < a.nrows) ? (a.rowList ? (long long)a.rowList[li] : li) : -1;
            if (pos < g.xcp && grow >= 0) {
                float2 v = make_float2(0.f, 0.f);
                if (pos < g.xc) {
                    const int p = g.odd ? __ldg(a.P.pos + pos) : pos;
                    v = C2[xslot(lrow, p)];
                }
                a.spec[grow * g.xcp + pos] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// X inverse: C2R along x for 16 rows per CTA (unnormalised, like cufftExecC2R)
// ------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kColThreads) x_inv_kernel(XArgs a)
{
    extern __shared__ float4 smem[];
    const Geometry g = a.g;
    const int L = a.P.L;
    const int tile_rows = g.odd ? L : (g.M + 1);
    float4* A = smem;
    float4* B = a.P.generic ? (A + (size_t)tile_rows * 8) : nullptr;
    float2* tw_s = reinterpret_cast<float2*>(A + (size_t)tile_rows * 8 * (a.P.generic ? 2 : 1));

    const int t = threadIdx.x;
    const int cp = t & 7, w = t >> 3, W = blockDim.x >> 3;
    const int lane = t & 31, warp = t >> 5, nwarps = blockDim.x >> 5;
    const int pl = lane & 7, rr = (lane >> 3) & 1, ph = lane >> 4;
    const long long row0 = (long long)blockIdx.x * 16;

    load_twiddles(tw_s, a.P.tw, L);

    {
        float2* A2 = reinterpret_cast<float2*>(A);
        const int nchunks = (g.xc + 15) >> 4;
        for (int u = warp; u < 8 * nchunks; u += nwarps) {
            const int rp = u / nchunks, c = u - rp * nchunks;
            const int pos = c * 16 + ph * 8 + pl;
            const int lrow = 2 * rp + rr;
            const long long grow = row0 + lrow;
            if (pos < g.xc) {
                if (!g.odd && grow < a.nrows) {
                    cp_async8(&A2[xslot(lrow, pos)], a.spec + grow * g.xcp + pos);
                    continue;
                }
                float2 v = make_float2(0.f, 0.f);
                if (grow < a.nrows) v = __ldg(a.spec + grow * g.xcp + pos);
                if (!g.odd) {
                    A2[xslot(lrow, pos)] = v;
                } else {
                    A2[xslot(lrow, __ldg(a.P.pos + pos))] = v;
                    if (pos > 0) A2[xslot(lrow, __ldg(a.P.pos + (g.nx - pos)))] = make_float2(v.x, -v.y);
                }
            }
        }
    }
    cp_async_wait_all();
    __syncthreads();

    if (!g.odd) {
        const int M = g.M;
        for (int k = w; k <= M / 2; k += W) {
            if (k == 0) {
                const int p0 = __ldg(a.P.pos);
                const float4 x0 = A[tile_idx<true>(p0, cp, 8)];
                const float4 xm = A[tile_idx<true>(M, cp, 8)];
                A[tile_idx<true>(p0, cp, 8)] = make_float4(x0.x + xm.x, x0.x - xm.x, x0.z + xm.z, x0.z - xm.z);
            } else {
                const int k2 = M - k;
                const int pk = __ldg(a.P.pos + k), pk2 = __ldg(a.P.pos + k2);
                const float4 va = A[tile_idx<true>(pk, cp, 8)];
                const float4 vb = A[tile_idx<true>(pk2, cp, 8)];
                const float2 tk = __ldg(a.twx + k);
                float4 o1, o2;
                {
                    float sr = va.x + vb.x, si = va.y - vb.y;
                    float Dr = va.x - vb.x, Di = va.y + vb.y;
                    // d = D * conj(w)
                    float dr = Dr * tk.x + Di * tk.y, di = Di * tk.x - Dr * tk.y;
                    o1.x = sr - di;
                    o1.y = si + dr;
                    o2.x = sr + di;
                    o2.y = dr - si;
                }
                {
                    float sr = va.z + vb.z, si = va.w - vb.w;
                    float Dr = va.z - vb.z, Di = va.w + vb.w;
                    float dr = Dr * tk.x + Di * tk.y, di = Di * tk.x - Dr * tk.y;
                    o1.z = sr - di;
                    o1.w = si + dr;
                    o2.z = sr + di;
                    o2.w = dr - si;
                }
                A[tile_idx<true>(pk, cp, 8)] = o1;
                if (k2 != k) A[tile_idx<true>(pk2, cp, 8)] = o2;
            }
        }
        __syncthreads();
    }

    float4* cur = engine_run<true, true>(a.P, A, B, tw_s, cp, w, W, 8, true);

    {
        const float2* C2 = reinterpret_cast<const float2*>(cur);
        const int npos = L;
        const int nchunks = (npos + 15) >> 4;
        for (int u = warp; u < 8 * nchunks; u += nwarps) {
            const int rp = u / nchunks, c = u - rp * nchunks;
            const int pos = c * 16 + ph * 8 + pl;
            const int lrow = 2 * rp + rr;
            const long long grow = row0 + lrow;
            if (pos < npos && grow < a.nrows) {
                const float2 v = C2[xslot(lrow, pos)];
                if (g.odd) a.out_real[grow * g.nx + pos] = v.x;
                else reinterpret_cast<float2*>(a.out_real + grow * g.nx)[pos] = v;
            }
        }
    }
}

// ------------------------------------------------------------------------------------------------
// Strided-axis passes (y, z): generic all-shared-memory version (any length, any tile width)
// ------------------------------------------------------------------------------------------------
template <int MODE>  // 0 forward, 1 inverse, 2 forward * H * scale, inverse
__global__ void __launch_bounds__(kColThreads) col_kernel(ColArgs a)
{
    extern __shared__ float4 smem[];
    const int L = a.P.L;
    const int txp = a.txp;
    float4* A = smem;
    float4* B = a.P.generic ? (A + (size_t)L * txp) : nullptr;
    float2* tw_s = reinterpret_cast<float2*>(A + (size_t)L * txp * (a.P.generic ? 2 : 1));

    const int t = threadIdx.x;
    const int cp = t % txp, w = t / txp, W = blockDim.x / txp;
    const int gi = blockIdx.x / a.tilesPerGroup;
    const int tt = blockIdx.x - gi * a.tilesPerGroup;
    const long long group = a.groupList ? (long long)a.groupList[gi] : (long long)gi;
    const int col0 = tt * 2 * txp;
    const int npairs = min(txp, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    const size_t off = (size_t)group * a.groupStride + col0 + 2 * cp;
    float2* base = a.data + off;

    load_twiddles(tw_s, a.P.tw, L);

    if (active) {
        for (int r = w; r < L; r += W) {
            const int p = (MODE == 1) ? __ldg(a.P.pos + r) : r;
            if (a.rowMask != nullptr && a.rowMask[r] == 0) A[p * txp + cp] = make_float4(0.f, 0.f, 0.f, 0.f);
            else cp_async16(&A[p * txp + cp], base + (size_t)r * a.stride);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    float4* cur;
    if (MODE == 1) {
        cur = engine_run<true, false>(a.P, A, B, tw_s, cp, w, W, txp, active);
    } else {
        cur = engine_run<false, false>(a.P, A, B, tw_s, cp, w, W, txp, active);
    }

    if (MODE == 2) {
        // Dst = c * (Src * Dst) with Src = PSF spectrum, Dst = image spectrum (reference :41-45, :54-58)
        const float2* hb = a.H + off;
        const float c = a.scale;
        if (active) {
            for (int p = w; p < L; p += W) {
                const int k = __ldg(a.P.rev + p);
                const float4 h = __ldg(reinterpret_cast<const float4*>(hb + (size_t)k * a.stride));
                const float4 v = cur[p * txp + cp];
                float4 o;
                o.x = c * (h.x * v.x - h.y * v.y);
                o.y = c * (h.y * v.x + h.x * v.y);
                o.z = c * (h.z * v.z - h.w * v.w);
                o.w = c * (h.w * v.z + h.z * v.w);
                cur[p * txp + cp] = o;
            }
        }
        __syncthreads();
        float4* oth = (cur == A) ? B : A;
        cur = engine_run<true, false>(a.P, cur, oth, tw_s, cp, w, W, txp, active);
    }

    if (active) {
        for (int p = w; p < L; p += W) {
            const int row = (MODE == 0) ? __ldg(a.P.rev + p) : p;
            *reinterpret_cast<float4*>(base + (size_t)row * a.stride) = cur[p * txp + cp];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
static size_t x_smem_bytes(const Geometry& g, const AxisPlanDev& P)
{
    const size_t tile_rows = g.odd ? (size_t)P.L : (size_t)g.M + 1;
    return tile_rows * 8 * sizeof(float4) * (P.generic ? 2 : 1) + (size_t)P.L * sizeof(float2);
}

bool x_pass_supported(const Geometry& g, const AxisPlanDev& P) { return x_smem_bytes(g, P) <= (size_t)kMaxDynSmem; }

int col_pick_txp(const AxisPlanDev& P)
{
    for (int txp = 8; txp >= 1; txp >>= 1) {
        size_t need = (size_t)P.L * txp * sizeof(float4) * (P.generic ? 2 : 1) + (size_t)P.L * sizeof(float2);
        if (need <= (size_t)kMaxDynSmem) return txp;
    }
    return 0;
}

template <typename K>
static void set_smem(K kernel, size_t bytes)
{
    if (bytes > 48 * 1024)
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes));
}

void launch_x_fwd(const XArgs& a, bool psf, cudaStream_t st)
{
    const size_t smem = x_smem_bytes(a.g, a.P);
    const long long tiles = (a.nrows + 15) / 16;
    if (tiles == 0) return;
    if (psf) {
        set_smem(x_fwd_kernel<1>, smem);
        x_fwd_kernel<1><<<(unsigned)tiles, kColThreads, smem, st>>>(a);
    } else {
        set_smem(x_fwd_kernel<0>, smem);
        x_fwd_kernel<0><<<(unsigned)tiles, kColThreads, smem, st>>>(a);
    }
    FC_CUDA_KERNEL();
}

void launch_x_inv(const XArgs& a, cudaStream_t st)
{
    const size_t smem = x_smem_bytes(a.g, a.P);
    const long long tiles = (a.nrows + 15) / 16;
    if (tiles == 0) return;
    set_smem(x_inv_kernel, smem);
    x_inv_kernel<<<(unsigned)tiles, kColThreads, smem, st>>>(a);
    FC_CUDA_KERNEL();
}

void launch_col(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const size_t smem = (size_t)a.P.L * a.txp * sizeof(float4) * (a.P.generic ? 2 : 1) + (size_t)a.P.L * sizeof(float2);
    const long long grid = ngroups * a.tilesPerGroup;
    if (grid == 0) return;
    if (grid > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    switch (mode) {
        case 0:
            set_smem(col_kernel<0>, smem);
            col_kernel<0><<<(unsigned)grid, kColThreads, smem, st>>>(a);
            break;
        case 1:
            set_smem(col_kernel<1>, smem);
            col_kernel<1><<<(unsigned)grid, kColThreads, smem, st>>>(a);
            break;
        default:
            set_smem(col_kernel<2>, smem);
            col_kernel<2><<<(unsigned)grid, kColThreads, smem, st>>>(a);
            break;
    }
    FC_CUDA_KERNEL();
}

}  // namespace fcb200
