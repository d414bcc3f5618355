// sm_100a kernels of the 3D FFT convolution hot path and their launchers.
//
//   x_fwd_kernel   real rows -> half spectrum rows (R2C along x); image loader or PSF gather loader
//                  (PSF zero-pad + circular shift of /root/reference/src/convolution3Dfft.cu:128-166
//                  fused into the load, never materialising the padded PSF)
//   col_kernel     strided-axis complex passes (y and z), generic version: forward, inverse, or fused
//                  forward -> x H * 1/N -> inverse (modulateAndNormalize_kernel of
//                  /root/reference/src/convolution3Dfft.cu:41-62 fused between the two z transforms);
//                  the fast version lives in fft_col_fast.cu
//   x_inv_kernel   half spectrum rows -> real rows (C2R along x)
//
// Together they replace cufftExecR2C x2 + modulateAndNormalize_kernel + cufftExecC2R
// (/root/reference/src/convolution3Dfft.cu:519-547) and the row re-layout loops (:474-486, :495-510,
// :561-575), which do not exist here: the X passes read and write dense rows directly.
#include "fft_xpass.cuh"

#include <algorithm>
#include <cstdlib>

namespace fcb200 {

// ------------------------------------------------------------------------------------------------
// Strided-axis passes (y, z): generic all-shared-memory version (any length, any tile width)
// ------------------------------------------------------------------------------------------------
template <int MODE, bool BIG>  // 0 forward, 1 inverse, 2 forward * H * scale, inverse; BIG: with the prime radices 11..23
__global__ void __launch_bounds__(kColThreads) col_kernel(ColArgs a)
{
    extern __shared__ float4 smem[];
    const int L = a.P.L;
    const int txp = a.txp;
    float4* A = smem;
    float4* B = a.P.generic ? (A + (size_t)L * txp) : nullptr;
    float4* tw_s = A + (size_t)L * txp * (a.P.generic ? 2 : 1);
    float4* rtw_s = tw_s + L;   // roots of the Rader n-point transform (BIG builds, plans with a Rader stage)

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
    // split layout (see ColArgs): address of transform row r on the split side (local buffer or peer GPUs)
    const bool has_split = a.split != nullptr || a.splitPeers != nullptr;
    float2* sbase = has_split ? reinterpret_cast<float2*>(1) : nullptr;   // only a flag below
    auto split_row = [&](int r) {
        const int blk = r / a.splitRows;
        float2* b0 = a.splitPeers ? a.splitPeers[blk] + a.splitPeerOffset : a.split + (size_t)blk * a.splitBlock;
        return b0 + (size_t)group * a.splitGroup + col0 + 2 * cp + (size_t)(r - blk * a.splitRows) * a.stride;
    };

    load_twiddles(tw_s, a.P.tw, L);
    if constexpr (BIG) load_rader_twiddles(rtw_s, a.P);

    if (active) {
        for (int r = w; r < L; r += W) {
            const int p = (MODE == 1) ? __ldg(a.P.pos + r) : r;
            if (a.rowMask != nullptr && a.rowMask[r] == 0) A[p * txp + cp] = make_float4(0.f, 0.f, 0.f, 0.f);
            else if (MODE == 2 && a.splitInPeers != nullptr) {
                const int blk = r / a.splitRows;
                cp_async16(&A[p * txp + cp], a.splitInPeers[blk] + (size_t)group * a.splitGroup + col0 + 2 * cp +
                                                 (size_t)(r - blk * a.splitRows) * a.stride);
            } else cp_async16(&A[p * txp + cp], (MODE == 1 && sbase) ? split_row(r) : base + (size_t)r * a.stride);
        }
    }
    cp_async_wait_all();
    __syncthreads();

    float4* cur;
    if (MODE == 1) {
        cur = engine_run<true, BIG>(a.P, A, B, tw_s, cp, w, W, txp, active, rtw_s);
    } else {
        cur = engine_run<false, BIG>(a.P, A, B, tw_s, cp, w, W, txp, active, rtw_s);
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
                // pair-planar: (x, y) = real parts, (z, w) = imaginary parts of the two pencils
                float4 o;
                o.x = c * (h.x * v.x - h.z * v.z);
                o.y = c * (h.y * v.y - h.w * v.w);
                o.z = c * (h.z * v.x + h.x * v.z);
                o.w = c * (h.w * v.y + h.y * v.w);
                cur[p * txp + cp] = o;
            }
        }
        __syncthreads();
        float4* oth = (cur == A) ? B : A;
        cur = engine_run<true, BIG>(a.P, cur, oth, tw_s, cp, w, W, txp, active, rtw_s);
    }

    if (active) {
        for (int p = w; p < L; p += W) {
            const int row = (MODE == 0) ? __ldg(a.P.rev + p) : p;
            float2* dst = ((MODE == 0 || MODE == 2) && sbase && (MODE == 0 || a.splitPeers)) ? split_row(row)
                                                                                            : base + (size_t)row * a.stride;
            *reinterpret_cast<float4*>(dst) = cur[p * txp + cp];
        }
    }
}

// ------------------------------------------------------------------------------------------------
// launchers
// ------------------------------------------------------------------------------------------------
size_t x_smem_bytes(const Geometry& g, const AxisPlanDev& P, int txp)
{
    const size_t rowt = 2 * (size_t)txp * (size_t)x_row_pitch(P.L, g.xcp) * sizeof(float2);
    return (size_t)P.L * txp * sizeof(float4) * (P.generic ? 2 : 1) + (size_t)(P.L + P.rader_n) * sizeof(float4) + rowt;
}

int x_pick_txp(const Geometry& g, const AxisPlanDev& P)
{
    for (int txp = 8; txp >= 1; txp >>= 1)
        if (x_smem_bytes(g, P, txp) <= (size_t)kMaxDynSmem) return txp;
    return 0;
}

bool x_pass_supported(const Geometry& g, const AxisPlanDev& P) { return x_pick_txp(g, P) > 0; }

int col_pick_txp(const AxisPlanDev& P)
{
    for (int txp = 8; txp >= 1; txp >>= 1) {
        size_t need = (size_t)P.L * txp * sizeof(float4) * (P.generic ? 2 : 1) + (size_t)(P.L + P.rader_n) * sizeof(float4);
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

static int x_threads()
{
    static const int t = [] {
        const char* e = std::getenv("FCB200_X_THREADS");
        int v = e ? std::atoi(e) : kColThreads;
        return std::max(64, std::min(v, kColThreads)) & ~31;
    }();
    return t;
}

void launch_x_fwd(const XArgs& a, bool psf, cudaStream_t st)
{
    if (launch_x_fwd_static(a, psf, st)) return;
    XArgs b = a;
    b.txp = x_pick_txp(a.g, a.P);
    const XArgs& a2 = b;
    const size_t smem = x_smem_bytes(a.g, a.P, b.txp);
    const long long tiles = (a.nrows + 2 * b.txp - 1) / (2 * b.txp);
    if (tiles == 0) return;
    auto go = [&](auto kernel) {
        set_smem(kernel, smem);
        kernel<<<(unsigned)tiles, x_threads(), smem, st>>>(a2);
    };
    if (a.P.big) {
        if (psf) go(x_fwd_kernel<1, DynPlanBig, kColThreads>);
        else go(x_fwd_kernel<0, DynPlanBig, kColThreads>);
    } else {
        if (psf) go(x_fwd_kernel<1, DynPlan, kColThreads>);
        else go(x_fwd_kernel<0, DynPlan, kColThreads>);
    }
    FC_CUDA_KERNEL();
}

void launch_x_inv(const XArgs& a, cudaStream_t st)
{
    if (launch_x_inv_static(a, st)) return;
    XArgs b = a;
    b.txp = x_pick_txp(a.g, a.P);
    const XArgs& a2 = b;
    const size_t smem = x_smem_bytes(a.g, a.P, b.txp);
    const long long tiles = (a.nrows + 2 * b.txp - 1) / (2 * b.txp);
    if (tiles == 0) return;
    if (a.P.big) {
        set_smem(x_inv_kernel<DynPlanBig, kColThreads>, smem);
        x_inv_kernel<DynPlanBig, kColThreads><<<(unsigned)tiles, x_threads(), smem, st>>>(a2);
    } else {
        set_smem(x_inv_kernel<DynPlan, kColThreads>, smem);
        x_inv_kernel<DynPlan, kColThreads><<<(unsigned)tiles, x_threads(), smem, st>>>(a2);
    }
    FC_CUDA_KERNEL();
}

void launch_col(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const size_t smem = (size_t)a.P.L * a.txp * sizeof(float4) * (a.P.generic ? 2 : 1) + (size_t)(a.P.L + a.P.rader_n) * sizeof(float4);
    const long long grid = ngroups * a.tilesPerGroup;
    if (grid == 0) return;
    if (grid > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    auto go = [&](auto kernel) {
        set_smem(kernel, smem);
        kernel<<<(unsigned)grid, kColThreads, smem, st>>>(a);
    };
    if (a.P.big) {
        if (mode == 0) go(col_kernel<0, true>);
        else if (mode == 1) go(col_kernel<1, true>);
        else go(col_kernel<2, true>);
    } else {
        if (mode == 0) go(col_kernel<0, false>);
        else if (mode == 1) go(col_kernel<1, false>);
        else go(col_kernel<2, false>);
    }
    FC_CUDA_KERNEL();
}

}  // namespace fcb200
