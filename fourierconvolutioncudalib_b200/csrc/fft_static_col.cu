// Compile-time specialised strided-axis (y / z) kernels for the hot transform lengths, and the dispatch
// that picks them when the run-time plan has exactly the same radix sequence.  Anything else falls back
// to the run-time-radix kernels (fft_col_fast.cu / fft_kernels.cu).
#include "fft_static_plans.h"

namespace fcb200 {

namespace {


// MODE 0 forward, 1 inverse, 2 fused forward x H x scale inverse (see fft_col_fast.cu)
// SPLIT: the strided side of the pass (output of modes 0 and 2, input of mode 1) uses the split / peer layout
template <int MODE, class P, int THREADS, int U, bool MASKED, int TXP, bool SPLIT = false>
__global__ void __launch_bounds__(THREADS) col_static_kernel(ColArgs a, int tilesPerGroup)
{
    constexpr int L = P::L, NW = THREADS / TXP;
    extern __shared__ float4 smem[];
    float4* sm = smem;
    float4* tw = sm + (size_t)L * TXP;

    const int t = threadIdx.x;
    const int cp = t % TXP, w = t / TXP;
    const int gi = blockIdx.x / tilesPerGroup;
    const int tt = blockIdx.x - gi * tilesPerGroup;
    const long long group = a.groupList ? (long long)a.groupList[gi] : (long long)gi;
    const int col0 = tt * 2 * TXP;
    const int npairs = min(TXP, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    const size_t off = (size_t)group * a.groupStride + col0 + 2 * cp;
    float2* base = a.data + off;
    const size_t stride = (size_t)a.stride;
    const RowsSplit srows{a.splitPeers, a.split, a.splitBlock,
                          (a.splitPeers ? a.splitPeerOffset : 0) + group * a.splitGroup + col0 + 2 * cp,
                          a.splitRows, SPLIT ? 1.0f / (float)a.splitRows : 0.f, stride};

    pdl_launch_dependents();
    load_twiddles(tw, a.P.tw, L);
    pdl_wait();
    __syncthreads();

    // A plain inverse on the linear layout runs the FORWARD stage sequence with re/im exchanged on the way in
    // and out (IDFT(x) = swap(DFT(swap(x)))): first stage straight from global, last stage straight to global,
    // exactly the cost of the forward pass.  The mirrored decimation-in-time sequence below serves the fused
    // pass (which continues from digit-reversed positions) and the split / peer input layout.
    constexpr bool SWAP = (MODE == 1 && !SPLIT);
    if (MODE == 0 || MODE == 2 || SWAP) {
        bool pulled = false;
        if constexpr (SPLIT && MODE == 2) {
            if (a.splitInPeers != nullptr) {   // pull exchange: the planes are read out of the GPUs that produced them
                const RowsSplit irows{a.splitInPeers, nullptr, 0, (long long)group * a.splitGroup + col0 + 2 * cp, a.splitRows,
                                      1.0f / (float)a.splitRows, stride};
                if (active) sfirst_fwd_rows<P::R0, L, NW, U, TXP>(irows, sm, tw, cp, w);
                pulled = true;
            }
        }
        if (!pulled && active) sfirst_fwd<P::R0, L, NW, U, MASKED, TXP, SWAP>(base, stride, sm, tw, cp, w, a.rowMask);
        __syncthreads();
        if constexpr (P::ns >= 3) {
            if (active) sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
            __syncthreads();
        }
        if constexpr (P::ns >= 4) {
            if (active) sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, TXP>(sm, tw, cp, w);
            __syncthreads();
        }
        if (MODE == 0 || SWAP) {
            if (!active) return;
            if constexpr (SPLIT) slast_fwd<P::RL, L, NW, TXP, false>(srows, sm, a.P.rev, cp, w);
            else slast_fwd<P::RL, L, NW, TXP, SWAP>(base, stride, sm, a.P.rev, cp, w);
            return;
        }
        if (active) smid_fused<P::RL, L, NW, U, TXP>(a.H + off, stride, sm, a.P.rev, cp, w, a.scale);
        __syncthreads();
    } else {
        if (active) {
            if constexpr (SPLIT) sfirst_inv<P::RL, L, NW, U, TXP>(srows, sm, a.P.rev, cp, w);
            else sfirst_inv<P::RL, L, NW, U, TXP>(base, stride, sm, a.P.rev, cp, w);
        }
        __syncthreads();
    }
    if constexpr (P::ns >= 4) {
        if (active) sstage<P::R2, L, P::R2 * P::R3, NW, true, TXP>(sm, tw, cp, w);
        __syncthreads();
    }
    if constexpr (P::ns >= 3) {
        if (active) sstage<P::R1, L, P::R1 * P::R2 * P::R3, NW, true, TXP>(sm, tw, cp, w);
        __syncthreads();
    }
    if (!active) return;
    if constexpr (SPLIT && MODE == 2) slast_inv<P::R0, L, NW, TXP>(srows, sm, tw, cp, w);
    else slast_inv<P::R0, L, NW, TXP>(base, stride, sm, tw, cp, w);
}


// ------------------------------------------------------------------------------------------------
// Persistent, software-pipelined variant: one CTA per SM loops over tiles; the next NBUF-1 tiles are
// in flight (cp.async, 16 bytes per copy, straight into shared memory, no registers) while the current
// one is transformed.  HBM latency is hidden by the prefetch depth instead of by occupancy.
//   MODE 0 forward, 1 inverse, 2 fused (the PSF-spectrum tile is prefetched next to the data tile).
// ------------------------------------------------------------------------------------------------
__device__ __forceinline__ void cpa16(void* smem_dst, const void* gmem_src)
{
    const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cpa_commit() { asm volatile("cp.async.commit_group;\n" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cpa_wait() { asm volatile("cp.async.wait_group %0;\n" ::"n"(N) : "memory"); }

// last forward stage, x H x c, first inverse stage, with the PSF-spectrum tile in shared memory
template <int R, int L, int NW>
__device__ __forceinline__ void smid_fused_sm(const float4* __restrict__ hs, float4* __restrict__ sm, int cp, int w, float c)
{
    constexpr int nb = L / R;
    constexpr int ITER = (nb + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        p2 r[R], i[R];
        load_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
        Dft<R>::run(r, i);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const float4 h = hs[(b * R + m) * 8 + cp];
            const p2 hr = make_float2(h.x, h.y), hi = make_float2(h.z, h.w);
            const p2 xr = pmuls(pfma(hr, r[m], pneg(pmul(hi, i[m]))), c);
            const p2 xi = pmuls(pfma(hi, r[m], pmul(hr, i[m])), c);
            r[m] = xr;
            i[m] = xi;
        }
        Dft<R>::run(i, r);
        store_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
    }
}

// TXP: column pairs per tile row (8 = 128-byte rows; 4 = 64-byte rows, half the tile size: more CTAs per SM)
template <int MODE, class P, int THREADS, int NBUF, bool MASKED, int TXP = 8>
__global__ void __launch_bounds__(THREADS, 1) col_pipe_kernel(ColArgs a, int tilesPerGroup, int totalTiles)
{
    constexpr int L = P::L, NW = THREADS / TXP, TILE = L * TXP;
    static_assert(MODE != 2 || TXP == 8, "the fused mode is written for 128-byte tile rows");
    extern __shared__ float4 smem[];
    float4* tw = smem;
    float4* dbuf = tw + L;
    float4* hbuf = dbuf + (size_t)NBUF * TILE;   // MODE 2 only
    int* pos_s = reinterpret_cast<int*>(hbuf + (MODE == 2 ? (size_t)NBUF * TILE : 0));   // [L] digit-reversal table

    const int t = threadIdx.x;
    const int cp = t % TXP, w = t / TXP;
    const size_t stride = (size_t)a.stride;

    pdl_launch_dependents();
    load_twiddles(tw, a.P.tw, L);
    if (MODE == 2) {
        for (int r = t; r < L; r += THREADS) pos_s[r] = __ldg(a.P.pos + r);
        __syncthreads();
    }
    pdl_wait();

    auto decode = [&](int tile, size_t& off) -> bool {
        const int gi = tile / tilesPerGroup;
        const int tt = tile - gi * tilesPerGroup;
        const long long group = a.groupList ? (long long)a.groupList[gi] : (long long)gi;
        const int col0 = tt * 2 * TXP;
        const int npairs = min(TXP, (a.rowLen - col0) >> 1);
        off = (size_t)group * a.groupStride + col0 + 2 * cp;
        return cp < npairs;
    };
    auto issue = [&](int tile, int slot) {
        size_t off;
        if (tile < totalTiles && decode(tile, off)) {
            float4* d = dbuf + (size_t)slot * TILE + cp;
            const float2* src = a.data + off;
#pragma unroll 4
            for (int r = w; r < L; r += NW) {
                const int p = (MODE == 2) ? pos_s[r] : r;   // H: row k is needed at position pos[k]
                const int pd = r;
                if (MASKED && a.rowMask[r] == 0) d[pd * TXP] = make_float4(0.f, 0.f, 0.f, 0.f);
                else cpa16(d + pd * TXP, src + (size_t)r * stride);
                if (MODE == 2) cpa16(hbuf + (size_t)slot * TILE + p * 8 + cp, a.H + off + (size_t)r * stride);
            }
        }
        cpa_commit();
    };

    for (int k = 0; k < NBUF - 1; ++k) issue(blockIdx.x + k * gridDim.x, k);

    int it = 0;
    for (int tile = blockIdx.x; tile < totalTiles; tile += gridDim.x, ++it) {
        const int slot = it % NBUF;
        cpa_wait<NBUF - 2>();
        __syncthreads();   // tile `it` has landed for every thread; everybody is done with tile it-1
        issue(tile + (NBUF - 1) * (int)gridDim.x, (it + NBUF - 1) % NBUF);
        size_t off;
        const bool active = decode(tile, off);
        float4* sm = dbuf + (size_t)slot * TILE;
        float2* base = a.data + off;
        // MODE 1 (plain inverse) runs the FORWARD stage sequence with re/im exchanged on the way in and out
        {
            if (active) sstage<P::R0, L, L, NW, false, TXP, MODE == 1>(sm, tw, cp, w);
            __syncthreads();
            if constexpr (P::ns >= 3) {
                if (active) sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
                __syncthreads();
            }
            if constexpr (P::ns >= 4) {
                if (active) sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, TXP>(sm, tw, cp, w);
                __syncthreads();
            }
            if (MODE != 2) {
                if (active) slast_fwd<P::RL, L, NW, TXP, MODE == 1>(RowsLinear{base, stride}, sm, a.P.rev, cp, w);
                continue;
            }
            if constexpr (MODE == 2) {
                if (active) smid_fused_sm<P::RL, L, NW>(hbuf + (size_t)slot * TILE, sm, cp, w, a.scale);
                __syncthreads();
            }
        }
        if constexpr (MODE == 2) {
            if constexpr (P::ns >= 4) {
                if (active) sstage<P::R2, L, P::R2 * P::R3, NW, true>(sm, tw, cp, w);
                __syncthreads();
            }
            if constexpr (P::ns >= 3) {
                if (active) sstage<P::R1, L, P::R1 * P::R2 * P::R3, NW, true>(sm, tw, cp, w);
                __syncthreads();
            }
            if (active) slast_inv<P::R0, L, NW>(base, stride, sm, tw, cp, w);
        }
    }
    cpa_wait<0>();
}

// SM count of the CURRENT device (one process may drive several GPUs: fc_multi.cu)
static int sm_count()
{
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    if (dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = v;
    }
    return cache[dev];
}

template <class P, int THREADS, int NBUF>
bool run_col_pipe(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const int tpg = (a.rowLen + 15) / 16;
    const long long total = ngroups * tpg;
    if (total == 0) return true;
    if ((size_t)P::L * (sizeof(float4) + sizeof(int)) + (size_t)NBUF * P::L * 8 * sizeof(float4) * (mode == 2 ? 2 : 1) >
        (size_t)kMaxDynSmem)
        return false;
    if (total > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    const size_t tile = (size_t)P::L * 8 * sizeof(float4);
    const size_t smem = (size_t)P::L * (sizeof(float4) + sizeof(int)) + (size_t)NBUF * tile * (mode == 2 ? 2 : 1);
    auto go = [&](auto kernel) {
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
        const int grid = (int)std::min<long long>(total, (long long)sm_count() * std::max(1, per_sm));
        launch_pdl(a.pdl != 0, kernel, dim3(grid), dim3(THREADS), smem, st, a, tpg, (int)total);
        FC_CUDA_KERNEL();
    };
    if (mode == 0 && a.rowMask) go(col_pipe_kernel<0, P, THREADS, NBUF, true>);
    else if (mode == 0) go(col_pipe_kernel<0, P, THREADS, NBUF, false>);
    else if (mode == 1) go(col_pipe_kernel<1, P, THREADS, NBUF, false>);
    else go(col_pipe_kernel<2, P, THREADS, NBUF, false>);
    return true;
}

// modes 0 / 1 without row mask only (what try_pipe_plain needs): two kernels per configuration
template <class P, int THREADS, int NBUF, int TXP = 8>
bool run_col_pipe_plain(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const int tpg = (a.rowLen + 2 * TXP - 1) / (2 * TXP);
    const long long total = ngroups * tpg;
    if (total == 0) return true;
    if (mode > 1 || a.rowMask) return false;
    const size_t tile = (size_t)P::L * TXP * sizeof(float4);
    const size_t smem = (size_t)P::L * (sizeof(float4) + sizeof(int)) + (size_t)NBUF * tile;
    if (smem > (size_t)kMaxDynSmem) return false;
    if (total > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    auto go = [&](auto kernel) {
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS, smem));
        const int grid = (int)std::min<long long>(total, (long long)sm_count() * std::max(1, per_sm));
        launch_pdl(a.pdl != 0, kernel, dim3(grid), dim3(THREADS), smem, st, a, tpg, (int)total);
        FC_CUDA_KERNEL();
    };
    if (mode == 0) go(col_pipe_kernel<0, P, THREADS, NBUF, false, TXP>);
    else go(col_pipe_kernel<1, P, THREADS, NBUF, false, TXP>);
    return true;
}

template <typename K>
void launch(K kernel, long long grid, int threads, size_t smem, cudaStream_t st, const ColArgs& a, int tpg)
{
    if (smem > 48 * 1024) FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(a.pdl != 0, kernel, dim3((unsigned)grid), dim3(threads), smem, st, a, tpg);
    FC_CUDA_KERNEL();
}

// split / peer layout variants are compiled for the power-of-two lengths only (slab mode targets those);
// other lengths fall back to the generic kernel
template <class P>
constexpr bool split_capable()
{
    return (P::L & (P::L - 1)) == 0 && P::L >= 64;
}

// One static configuration: plan P, CTA size, load batch U, tile width TXP (column pairs per row).
// MODES: which modes this configuration is instantiated for (bit 0: forward, bit 1: inverse, bit 2: fused) --
// a configuration used for one mode only does not compile the other kernels
template <class P, int THREADS, int U, int TXP, int MODES = 7>
void run_col(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const int tpg = (a.rowLen + 2 * TXP - 1) / (2 * TXP);
    const long long grid = ngroups * tpg;
    if (grid == 0) return;
    if (grid > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    if (!((MODES >> mode) & 1)) throw std::runtime_error("fcb200: internal error, column kernel mode not compiled");
    const size_t smem = (size_t)P::L * TXP * sizeof(float4) + (size_t)P::L * sizeof(float4);
    if constexpr (split_capable<P>()) {
        if (a.split || a.splitPeers) {
            if constexpr ((MODES & 1) != 0)
                if (mode == 0) launch(col_static_kernel<0, P, THREADS, U, false, TXP, true>, grid, THREADS, smem, st, a, tpg);
            if constexpr ((MODES & 2) != 0)
                if (mode == 1) launch(col_static_kernel<1, P, THREADS, U, false, TXP, true>, grid, THREADS, smem, st, a, tpg);
            if constexpr ((MODES & 4) != 0)
                if (mode == 2) launch(col_static_kernel<2, P, THREADS, U, false, TXP, true>, grid, THREADS, smem, st, a, tpg);
            return;
        }
    }
    if constexpr ((MODES & 1) != 0) {
        if (mode == 0 && a.rowMask) launch(col_static_kernel<0, P, THREADS, U, true, TXP>, grid, THREADS, smem, st, a, tpg);
        else if (mode == 0) launch(col_static_kernel<0, P, THREADS, U, false, TXP>, grid, THREADS, smem, st, a, tpg);
    }
    if constexpr ((MODES & 2) != 0)
        if (mode == 1) launch(col_static_kernel<1, P, THREADS, U, false, TXP>, grid, THREADS, smem, st, a, tpg);
    if constexpr ((MODES & 4) != 0)
        if (mode == 2) launch(col_static_kernel<2, P, THREADS, U, false, TXP>, grid, THREADS, smem, st, a, tpg);
}


}  // namespace

// ------------------------------------------------------------------------------------------------
// Fused z pass with the PSF spectrum computed ON THE FLY (SaveMemory path): when the placed PSF spans
// at most 16 consecutive z planes (mod L), every CTA derives the H tile of its 16 pencils from the 16
// window rows of the (x,y)-transformed PSF -- one input-pruned radix-16 butterfly per output residue
// (see psf_z_pruned_kernel) -- into shared memory, then runs forward z, x H x 1/N, inverse z as usual.
// No image-sized PSF spectrum is ever written or read: -1/3 of the fused pass's HBM traffic and no PSF z
// kernel at all.  a.H points at the COMPACT buffer [<=16 planes][ny][xcp] of (x,y)-transformed PSF planes;
// a.winSlot[n] is the compact plane of window position n (z = (z0 + n) mod L), -1 when it holds no tap.
// ------------------------------------------------------------------------------------------------
template <int R, int L, int NW>
__device__ __forceinline__ void smid_fused_nat(const float4* __restrict__ hs, float4* __restrict__ sm,
                                               const int* __restrict__ rev, int cp, int w, float c)
{
    constexpr int nb = L / R, fs = L / R;
    constexpr int ITER = (nb + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        p2 r[R], i[R];
        load_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
        const int k0 = __ldg(rev + b * R);
        Dft<R>::run(r, i);
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const float4 h = hs[(k0 + m * fs) * 8 + cp];   // H tile is in natural frequency order
            const p2 hr = make_float2(h.x, h.y), hi = make_float2(h.z, h.w);
            const p2 xr = pmuls(pfma(hr, r[m], pneg(pmul(hi, i[m]))), c);
            const p2 xi = pmuls(pfma(hi, r[m], pmul(hr, i[m])), c);
            r[m] = xr;
            i[m] = xi;
        }
        Dft<R>::run(i, r);
        store_pairs<R>(sm, b * R * 8 + cp, 8, r, i);
    }
}

template <class P, int THREADS, int U>
__global__ void __launch_bounds__(THREADS) col_otf_kernel(ColArgs a, int tilesPerGroup, int z0)
{
    constexpr int L = P::L, NW = THREADS / 8, Q = L / 16;
    static_assert(L % 16 == 0, "on-the-fly PSF spectrum needs L % 16 == 0");
    extern __shared__ float4 smem[];
    float4* sm = smem;
    float4* tw = sm + (size_t)L * 8;
    float4* ht = tw + L;                 // H tile, natural frequency order
    float4* win = ht + (size_t)L * 8;    // 16 window rows

    const int t = threadIdx.x;
    const int cp = t & 7, w = t >> 3;
    const int gi = blockIdx.x / tilesPerGroup;
    const int tt = blockIdx.x - gi * tilesPerGroup;
    const long long group = a.groupList ? (long long)a.groupList[gi] : (long long)gi;
    const int col0 = tt * 16;
    const int npairs = min(8, (a.rowLen - col0) >> 1);
    const bool active = cp < npairs;
    const size_t off = (size_t)group * a.groupStride + col0 + 2 * cp;
    float2* base = a.data + off;
    const size_t stride = (size_t)a.stride;

    pdl_launch_dependents();
    load_twiddles(tw, a.P.tw, L);
    pdl_wait();
    // window rows: issued now, consumed after the first stage so that their latency overlaps with the
    // first-stage loads (q & 7 == cp because THREADS % 8 == 0)
    constexpr int NWIN = (128 + THREADS - 1) / THREADS;
    float4 wv[NWIN];
#pragma unroll
    for (int u = 0; u < NWIN; ++u) {
        const int q = t + u * THREADS;
        wv[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (q < 128 && active) {
            const int slot = __ldg(a.winSlot + (q >> 3));   // compact plane that holds window position n
            if (slot >= 0) wv[u] = __ldg(reinterpret_cast<const float4*>(a.H + off + (size_t)slot * stride));
        }
    }
    __syncthreads();   // twiddles ready

    if (active) sfirst_fwd<P::R0, L, NW, U, false, 8>(base, stride, sm, tw, cp, w, nullptr);
#pragma unroll
    for (int u = 0; u < NWIN; ++u) {
        const int q = t + u * THREADS;
        if (q < 128) win[q] = wv[u];
    }
    __syncthreads();   // first-stage outputs and window rows ready

    if (active) {
        // H[k1*Q + k2] = w16^(z0 k1) sum_n [ win[n] w_L^((z0+n) k2) ] w16^(n k1)
        const int s16 = z0 & 15;
        for (int k2 = w; k2 < Q; k2 += NW) {
            p2 r[16], i[16];
            int e = (z0 * k2) % L;
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                const float4 v = win[n * 8 + cp];
                r[n] = make_float2(v.x, v.y);
                i[n] = make_float2(v.z, v.w);
                cmul(r[n], i[n], tw[e]);
                e += k2;
                if (e >= L) e -= L;
            }
            Dft<16>::run(r, i);
#pragma unroll
            for (int k1 = 0; k1 < 16; ++k1) {
                if (s16 != 0 && k1 != 0) cmul(r[k1], i[k1], tw[((s16 * k1) & 15) * Q]);
                ht[(k1 * Q + k2) * 8 + cp] = make_float4(r[k1].x, r[k1].y, i[k1].x, i[k1].y);
            }
        }
    }
    // the H tile is consumed by the mid stage only: the next stage barrier covers it (ns >= 3)
    if constexpr (P::ns == 2) __syncthreads();
    if constexpr (P::ns >= 3) {
        if (active) sstage<P::R1, L, L / P::R0, NW, false, 8>(sm, tw, cp, w);
        __syncthreads();
    }
    if constexpr (P::ns >= 4) {
        if (active) sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, 8>(sm, tw, cp, w);
        __syncthreads();
    }
    if (active) smid_fused_nat<P::RL, L, NW>(ht, sm, a.P.rev, cp, w, a.scale);
    __syncthreads();
    if constexpr (P::ns >= 4) {
        if (active) sstage<P::R2, L, P::R2 * P::R3, NW, true, 8>(sm, tw, cp, w);
        __syncthreads();
    }
    if constexpr (P::ns >= 3) {
        if (active) sstage<P::R1, L, P::R1 * P::R2 * P::R3, NW, true, 8>(sm, tw, cp, w);
        __syncthreads();
    }
    if (active) slast_inv<P::R0, L, NW, 8>(base, stride, sm, tw, cp, w);
}

template <class P, int THREADS, int U>
bool try_col_otf(const ColArgs& a, long long ngroups, int z0, cudaStream_t st, bool probe)
{
    if (!plan_matches<P>(a.P)) return false;
    const size_t smem = ((size_t)P::L * 8 * 2 + P::L + 128) * sizeof(float4);
    if (2 * smem > (size_t)kMaxDynSmem) return false;   // keep at least two CTAs per SM
    if (probe) return true;
    const int tpg = (a.rowLen + 15) / 16;
    const long long grid = ngroups * tpg;
    if (grid == 0) return true;
    if (grid > 0x7fffffffLL) throw std::runtime_error("fcb200: volume too large for one launch");
    auto kernel = col_otf_kernel<P, THREADS, U>;
    if (smem > 48 * 1024) FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    launch_pdl(a.pdl != 0, kernel, dim3((unsigned)grid), dim3(THREADS), smem, st, a, tpg, z0);
    FC_CUDA_KERNEL();
    return true;
}

static int pipe_mode()
{
    static const int v = env_int("FCB200_PIPE", 0);
    return v;
}

// Plain (non-fused) y / z passes of the long pencils run on the persistent multi-buffered kernel by default:
// measured 0.117 ms against 0.127 ms per pass on C3 (profiles/r01_notes.md).  FCB200_PIPE_Y=0 turns it off.
static bool try_pipe_plain(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    static const int on = env_int("FCB200_PIPE_Y", 1);
    if (!on || mode == 2 || a.split || a.splitPeers || a.rowMask || a.groupList) return false;
    const long long total = ngroups * ((a.rowLen + 15) / 16);
    if (total < 4LL * sm_count()) return false;   // too few tiles to fill the pipeline
    static const int p512 = env_int("FCB200_PIPE512", 0);   // experiments: 64-byte tile rows, more CTAs per SM
    if (plan_matches<P512>(a.P)) {
        if (p512 == 1) return run_col_pipe_plain<P512, 256, 3, 4>(a, mode, ngroups, st);
        if (p512 == 2) return run_col_pipe_plain<P512, 256, 2, 4>(a, mode, ngroups, st);
        if (p512 == 3) return run_col_pipe_plain<P512, 512, 3, 4>(a, mode, ngroups, st);
    }
    if (plan_matches<P512>(a.P)) return run_col_pipe<P512, 512, 3>(a, mode, ngroups, st);
    // shorter pencils: TWO tile buffers and two (or three) CTAs per SM beat three buffers in one CTA
    // (384^3 y passes 0.110 -> 0.097 ms, profiles/r01_sweep_pipe384.jsonl); FCB200_PIPE_SHORT=0 turns these off
    static const int pshort = env_int("FCB200_PIPE_SHORT", 1);
    if (pshort == 0) return false;
    if (plan_matches<P384>(a.P)) return run_col_pipe_plain<P384, 192, 2>(a, mode, ngroups, st);
    // measured on the caller-padded grids (profiles/r01_sweep_pipe_short.jsonl): 560: 0.181 -> 0.164 ms, 300: 0.068 -> 0.056 ms;
    // 420 (0.164 -> 0.183) and 448 / 256 (unchanged) stay on the one-tile-per-CTA kernels
    if (plan_matches<P560>(a.P)) return run_col_pipe_plain<P560, 320, 3>(a, mode, ngroups, st);
    // L = 256 plain passes (256^3: 0.0307 / 0.0293 -> 0.0271 / 0.0255 ms, profiles/r01_sweep256.jsonl); three buffers or
    // 256-thread CTAs are slower, and L = 270 does not move
    if (plan_matches<P256>(a.P)) return run_col_pipe_plain<P256, 128, 2>(a, mode, ngroups, st);
    return false;
}


// true when a compile-time column kernel exists for the plan (the run-time-radix TMA pipeline leaves those alone)
bool col_static_has_plan(const AxisPlanDev& P)
{
    return plan_matches<P64>(P) || plan_matches<P128>(P) || plan_matches<P256>(P) || plan_matches<P256b>(P) ||
           plan_matches<P384>(P) || plan_matches<P512>(P) || plan_matches<P1024>(P) || plan_matches<P2048>(P) ||
           plan_matches<P560>(P) || plan_matches<P560z>(P) || plan_matches<P448>(P) || plan_matches<P448y>(P) ||
           plan_matches<P420>(P) || plan_matches<P300>(P) || plan_matches<P270>(P);
}

bool launch_col_static(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    if (!static_enabled()) return false;
    if (a.txp == 4 && plan_matches<P2048>(a.P) && !((a.split || a.splitPeers) && a.rowMask)) {
        run_col<P2048, 512, 1, 4>(a, mode, ngroups, st);
        return true;
    }
    if (a.txp != 8) return false;
    if (pipe_mode() == 0 && try_pipe_plain(a, mode, ngroups, st)) return true;
    if ((a.split || a.splitPeers) && ((a.P.L & (a.P.L - 1)) != 0 || a.P.L < 64 || a.rowMask || pipe_mode() > 0)) return false;
    if (plan_matches<P64>(a.P)) run_col<P64, 64, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P128>(a.P)) run_col<P128, 64, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P256>(a.P)) run_col<P256, 128, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P384>(a.P)) {
        // fused pass: three PSF-spectrum batches in flight (384^3: 0.214 -> 0.169 ms)
        if (mode == 2) run_col<P384, 192, 3, 8, 4>(a, mode, ngroups, st);
        else run_col<P384, 192, 1, 8, 3>(a, mode, ngroups, st);
    }
    else if (plan_matches<P512>(a.P) && pipe_mode() > 0 &&
             (pipe_mode() == 2   ? run_col_pipe<P512, 512, 2>(a, mode, ngroups, st)
              : pipe_mode() == 3 ? run_col_pipe<P512, 1024, 3>(a, mode, ngroups, st)
                                 : run_col_pipe<P512, 512, 3>(a, mode, ngroups, st))) {
    } else if (plan_matches<P256b>(a.P) && pipe_mode() > 0 &&
               (pipe_mode() == 2   ? run_col_pipe<P256b, 256, 2>(a, mode, ngroups, st)
                : pipe_mode() == 3 ? run_col_pipe<P256b, 512, 3>(a, mode, ngroups, st)
                                   : run_col_pipe<P256b, 256, 3>(a, mode, ngroups, st))) {
    } else if (plan_matches<P512>(a.P)) {
        // measured (profiles/r01_notes.md, r01_sweep_col512_variants.jsonl): plain passes like 512-thread CTAs (more
        // warps), the fused pass is register-bound there and prefers 256 threads with two butterflies in flight
        if (mode == 2) run_col<P512, 256, 2, 8, 4>(a, mode, ngroups, st);
        else run_col<P512, 512, 1, 8, 3>(a, mode, ngroups, st);
    } else if (plan_matches<P256b>(a.P)) {
        // fused pass: all four PSF-spectrum batches of a thread in flight at once (0.1765 -> 0.1625 ms on C3)
        if (mode == 2) run_col<P256b, 128, 4, 8, 4>(a, mode, ngroups, st);
        else run_col<P256b, 128, 2, 8, 3>(a, mode, ngroups, st);
    }
    else if (plan_matches<P560>(a.P) && mode != 2) run_col<P560, 320, 1, 8, 3>(a, mode, ngroups, st);   // y axis only
    else if (plan_matches<P448>(a.P)) {
        // fused pass: 256-thread CTAs, two PSF-spectrum batches in flight (420x420x448: 0.365 -> 0.219 ms,
        // profiles/r01_sweep_z448.jsonl); the plain passes keep 512 threads
        if (mode == 2) run_col<P448, 256, 2, 8, 4>(a, mode, ngroups, st);
        else run_col<P448, 512, 1, 8, 3>(a, mode, ngroups, st);
    }
    // two-stage plans with fat radices: one worker per butterfly of the larger stage (L / 15 = 20, L / 20 = 21, ...)
    else if (plan_matches<P270>(a.P)) run_col<P270, 160, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P420>(a.P)) run_col<P420, 192, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P300>(a.P)) run_col<P300, 160, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P448y>(a.P) && mode != 2) run_col<P448y, 224, 1, 8, 3>(a, mode, ngroups, st);   // y axis only
    else if (plan_matches<P560z>(a.P)) run_col<P560z, 224, 1, 8>(a, mode, ngroups, st);
    else if (plan_matches<P1024>(a.P)) {
        // plain passes: 64-byte row segments, two 256-thread CTAs per SM (1024x1024x256: 0.568 -> 0.529 ms)
        if (mode == 2) run_col<P1024, 512, 1, 8, 4>(a, mode, ngroups, st);
        else run_col<P1024, 256, 2, 4, 3>(a, mode, ngroups, st);
    }
    else return false;
    return true;
}

bool launch_col_otf(const ColArgs& a, long long ngroups, int z0, cudaStream_t st, bool probe)
{
    static const bool on = env_int("FCB200_OTF", 1) != 0;
    if (!on || !static_enabled() || a.txp != 8) return false;
    static const int t256 = env_int("FCB200_OTF_T", 256);
    if (t256 == 128 && try_col_otf<P256b, 128, 2>(a, ngroups, z0, st, probe)) return true;
    return try_col_otf<P256b, 256, 1>(a, ngroups, z0, st, probe) || try_col_otf<P128, 64, 1>(a, ngroups, z0, st, probe) ||
           try_col_otf<P64, 64, 1>(a, ngroups, z0, st, probe) || try_col_otf<P384, 192, 1>(a, ngroups, z0, st, probe);
}


}  // namespace fcb200
