// TMA-staged strided-axis (y / z) passes for sm_100a.
//
// A persistent CTA per SM walks over tiles of [L transform positions][TXP column pairs] (128-byte rows for TXP = 8).
// Every tile is brought in by the Tensor Memory Accelerator (cp.async.bulk.tensor, completion on an mbarrier) into a
// ring of NBUF shared-memory buffers, transformed IN PLACE by the compile-time specialised stages of fft_static.cuh
// (shared memory only: the consumer warps issue no global load, no global store and no address arithmetic for them),
// and written back by a TMA tensor store.  A producer warp owns all TMA traffic: the loads of the next NBUF-1 tiles
// and the store of the previous tile are in flight while the consumer warps transform the current tile, and a
// slot is refilled as soon as its store has been read out of shared memory.
//
// Digit reversal is done by the TMA unit itself.  The in-place decimation-in-frequency stages leave frequency
// rev[p] at position p (mixed-radix digit reversal).  Instead of permuting in the kernel, the OUTPUT tensor map
// describes the transform axis as a ns-dimensional tensor, one dimension per stage digit, ordered so that walking
// the shared-memory tile linearly (position p) walks the global rows in digit-reversed order:
//   p = ((m0 * R1 + m1) * R2 + m2)   holds   k = m0 + R0 * (m1 + R1 * m2)
//   => dimensions (innermost first): [columns] [m2: R2 entries, stride R0*R1 rows] [m1: R1, stride R0 rows]
//      [m0: R0, stride 1 row] [group]
// and ONE tensor store of the whole tile lands every row in natural frequency order.  The fused z pass loads its
// PSF-spectrum tile through the same kind of map, so that H arrives in shared memory already in digit-reversed
// position order and the multiply needs no index table.
//
//   MODE 0 forward, 1 inverse (forward stage sequence on re/im-exchanged data: IDFT(x) = swap(DFT(swap(x)))),
//   MODE 2 fused: forward . x H x 1/N . inverse (modulateAndNormalize_kernel of
//          /root/reference/src/convolution3Dfft.cu:41-62 between the two z transforms), natural order in and out.
//
// Replaces (with the x passes) the cuFFT executions of /root/reference/src/convolution3Dfft.cu:519-547.
#include <cuda.h>

#include <mutex>

#include "fft_static_plans.h"

namespace fcb200 {

namespace {

// ---- host: tensor maps -------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn()
{
    static EncodeTiledFn fn = [] {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) != cudaSuccess ||
            q != cudaDriverEntryPointSuccess)
            p = nullptr;
        cudaGetLastError();
        return reinterpret_cast<EncodeTiledFn>(p);
    }();
    return fn;
}

struct MapSpec {
    int rank = 0;
    cuuint64_t dims[5];
    cuuint64_t strides[4];   // bytes, dims 1..rank-1
    cuuint32_t box[5];
};

bool encode(CUtensorMap* m, void* base, const MapSpec& s)
{
    EncodeTiledFn fn = encode_fn();
    if (!fn) return false;
    const cuuint32_t ones[5] = {1, 1, 1, 1, 1};
    return fn(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)s.rank, base, s.dims, s.strides, s.box, ones,
              CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// natural order: [cols][L rows][groups]; rows are fetched in boxes of `boxRows` rows
MapSpec natural_spec(const ColArgs& a, long long ngroups, int L, int boxRows, int boxFloats)
{
    MapSpec s;
    s.rank = 3;
    s.dims[0] = (cuuint64_t)a.rowLen * 2;
    s.dims[1] = (cuuint64_t)L;
    s.dims[2] = (cuuint64_t)std::max<long long>(1, ngroups);
    s.strides[0] = (cuuint64_t)a.stride * sizeof(float2);
    s.strides[1] = ngroups > 1 ? (cuuint64_t)a.groupStride * sizeof(float2) : s.strides[0] * (cuuint64_t)L;
    s.box[0] = (cuuint32_t)boxFloats;
    s.box[1] = (cuuint32_t)boxRows;
    s.box[2] = 1;
    return s;
}

// digit-reversing order (see the header comment): one dimension per stage, last stage innermost
template <class P>
bool permuted_spec(const ColArgs& a, long long ngroups, int boxFloats, MapSpec& s)
{
    const int rad[4] = {P::R0, P::R1, P::R2, P::R3};
    const bool with_groups = ngroups > 1;
    s.rank = 1 + P::ns + (with_groups ? 1 : 0);
    if (s.rank > 5) return false;
    s.dims[0] = (cuuint64_t)a.rowLen * 2;
    s.box[0] = (cuuint32_t)boxFloats;
    const cuuint64_t row = (cuuint64_t)a.stride * sizeof(float2);
    // k-stride (in rows) of the digit of stage i = R0 * ... * R(i-1)
    cuuint64_t kstride[4], acc = 1;
    for (int i = 0; i < P::ns; ++i) {
        kstride[i] = acc;
        acc *= (cuuint64_t)rad[i];
    }
    for (int d = 0; d < P::ns; ++d) {           // dimension 1 + d  <->  stage ns-1-d
        const int st = P::ns - 1 - d;
        s.dims[1 + d] = (cuuint64_t)rad[st];
        s.strides[d] = kstride[st] * row;
        s.box[1 + d] = (cuuint32_t)rad[st];
    }
    if (with_groups) {
        s.dims[1 + P::ns] = (cuuint64_t)ngroups;
        s.strides[P::ns] = (cuuint64_t)a.groupStride * sizeof(float2);
        s.box[1 + P::ns] = 1;
    }
    return true;
}

// ---- device: mbarrier / TMA primitives ---------------------------------------------------------------------
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(unsigned long long* bar, int count)
{
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long* bar, unsigned bytes)
{
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long* bar, unsigned parity)
{
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "WAIT_%=:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n"
        "@p bra DONE_%=;\n"
        "bra WAIT_%=;\n"
        "DONE_%=:\n"
        "}\n" ::"r"(smem_u32(bar)),
        "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3,
                                            unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5}], [%6];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_load_5d(void* dst, const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4,
                                            unsigned long long* bar)
{
    asm volatile(
        "cp.async.bulk.tensor.5d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4, %5, %6}], [%7];" ::"r"(
            smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void tma_store_3d(const CUtensorMap* map, int c0, int c1, int c2, const void* src)
{
    asm volatile("cp.async.bulk.tensor.3d.global.shared::cta.bulk_group [%0, {%1, %2, %3}], [%4];" ::"l"(map), "r"(c0), "r"(c1),
                 "r"(c2), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_4d(const CUtensorMap* map, int c0, int c1, int c2, int c3, const void* src)
{
    asm volatile("cp.async.bulk.tensor.4d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4}], [%5];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_store_5d(const CUtensorMap* map, int c0, int c1, int c2, int c3, int c4, const void* src)
{
    asm volatile("cp.async.bulk.tensor.5d.global.shared::cta.bulk_group [%0, {%1, %2, %3, %4, %5}], [%6];" ::"l"(map), "r"(c0),
                 "r"(c1), "r"(c2), "r"(c3), "r"(c4), "r"(smem_u32(src))
                 : "memory");
}
__device__ __forceinline__ void tma_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void tma_wait_read()
{
    asm volatile("cp.async.bulk.wait_group.read %0;" ::"n"(N) : "memory");
}
// generic-proxy writes to shared memory -> visible to the async proxy (the TMA store that follows the barrier)
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// whole tile through the digit-reversing map: coordinates (column, 0, ..., 0, group)
template <int NS, bool GROUPS>
__device__ __forceinline__ void perm_store(const CUtensorMap* map, int c0, int group, const void* src)
{
    if constexpr (NS == 2 && !GROUPS) tma_store_3d(map, c0, 0, 0, src);
    else if constexpr (NS == 2 && GROUPS) tma_store_4d(map, c0, 0, 0, group, src);
    else if constexpr (NS == 3 && !GROUPS) tma_store_4d(map, c0, 0, 0, 0, src);
    else if constexpr (NS == 3 && GROUPS) tma_store_5d(map, c0, 0, 0, 0, group, src);
    else tma_store_5d(map, c0, 0, 0, 0, 0, src);   // NS == 4, no groups
}
template <int NS, bool GROUPS>
__device__ __forceinline__ void perm_load(void* dst, const CUtensorMap* map, int c0, int group, unsigned long long* bar)
{
    if constexpr (NS == 2 && !GROUPS) tma_load_3d(dst, map, c0, 0, 0, bar);
    else if constexpr (NS == 2 && GROUPS) tma_load_4d(dst, map, c0, 0, 0, group, bar);
    else if constexpr (NS == 3 && !GROUPS) tma_load_4d(dst, map, c0, 0, 0, 0, bar);
    else if constexpr (NS == 3 && GROUPS) tma_load_5d(dst, map, c0, 0, 0, 0, group, bar);
    else tma_load_5d(dst, map, c0, 0, 0, 0, 0, bar);
}

// in-place stage with re/im exchanged on the way out (last stage of a MODE 1 pass)
template <int R, int L, int Li, int NW, int TXP>
__device__ __forceinline__ void sstage_swapout(float4* __restrict__ buf, const float4* __restrict__ tw, int cp, int w)
{
    constexpr int S = Li / R, nb = L / R, tstep = L / Li;
    constexpr int ITER = (nb + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        const int beta = b / S, j = b % S;
        const int idx0 = (beta * Li + j) * TXP + cp;
        p2 r[R], i[R];
        load_pairs<R>(buf, idx0, S * TXP, r, i);
        Dft<R>::run(r, i);
        if (S > 1) {
#pragma unroll
            for (int m = 1; m < R; ++m) cmul(r[m], i[m], tw[j * (m * tstep)]);
        }
        store_pairs<R>(buf, idx0, S * TXP, i, r);
    }
}

// In-place stage whose twiddles live in REGISTERS.  In a persistent kernel the twiddles of the second stage are the same
// for every butterfly a thread ever runs (j = b mod S with NW a multiple of S), and in shared memory they are the
// accesses that conflict: consecutive workers read roots m*tstep*16 bytes apart, a multiple of 128 bytes once
// tstep >= 8, i.e. a 4-way bank conflict on every twiddle load (ncu: 15 % of the shared-memory wavefronts of the
// y pass).  tc / ts: cos / -sin of w_L^(j m tstep), m = 1 .. R-1.
template <int R, int L, int Li, int NW, bool INV, int TXP>
__device__ __forceinline__ void sstage_regtw(float4* __restrict__ buf, int cp, int w, const float* tc, const float* ts)
{
    constexpr int S = Li / R, nb = L / R;
    constexpr int ITER = (nb + NW - 1) / NW;
    static_assert(S > 1 && NW % S == 0, "register twiddles need a tile-invariant j");
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        const int beta = b / S, j = b % S;
        const int idx0 = (beta * Li + j) * TXP + cp;
        p2 r[R], i[R];
        load_pairs<R>(buf, idx0, S * TXP, r, i);
        if (INV) {
#pragma unroll
            for (int k = 1; k < R; ++k) cmulc(r[k], i[k], make_float4(tc[k - 1], tc[k - 1], ts[k - 1], ts[k - 1]));
            Dft<R>::run(i, r);
        } else {
            Dft<R>::run(r, i);
#pragma unroll
            for (int m = 1; m < R; ++m) cmul(r[m], i[m], make_float4(tc[m - 1], tc[m - 1], ts[m - 1], ts[m - 1]));
        }
        store_pairs<R>(buf, idx0, S * TXP, r, i);
    }
}

// last forward stage, x H x c, first inverse stage; H tile in shared memory in the SAME (digit-reversed) order
template <int R, int L, int NW, int TXP>
__device__ __forceinline__ void smid_fused_tma(const float4* __restrict__ hs, float4* __restrict__ sm, int cp, int w, float c)
{
    constexpr int nb = L / R;
    constexpr int ITER = (nb + NW - 1) / NW;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int b = w + it * NW;
        if ((nb % NW) != 0 && b >= nb) break;
        p2 r[R], i[R];
        load_pairs<R>(sm, b * R * TXP + cp, TXP, r, i);
        Dft<R>::run(r, i);
        // Dst = c * (Src * Dst), Src = PSF spectrum (reference mulAndScale, src/convolution3Dfft.cu:41-45)
#pragma unroll
        for (int m = 0; m < R; ++m) {
            const float4 h = hs[(b * R + m) * TXP + cp];
            const p2 hr = make_float2(h.x, h.y), hi = make_float2(h.z, h.w);
            const p2 xr = pmuls(pfma(hr, r[m], pneg(pmul(hi, i[m]))), c);
            const p2 xi = pmuls(pfma(hi, r[m], pmul(hr, i[m])), c);
            r[m] = xr;
            i[m] = xi;
        }
        Dft<R>::run(i, r);
        store_pairs<R>(sm, b * R * TXP + cp, TXP, r, i);
    }
}

struct TmaArgs {
    const float2* tw;     // L roots
    int tilesPerGroup;
    int totalTiles;
    int boxRows;          // rows per natural-order box (L % boxRows == 0)
    float scale;
    int reverse;          // walk the tiles from the last to the first
    const int* pos;       // MODE 3: pos[k] = position that holds frequency k
    int z0;               // MODE 3: first plane of the PSF window
};

// PSF-spectrum tile of 2*TXP pencils from the (x,y)-transformed PSF planes of the window [z0, z0 + 16*NH) mod L:
//   H[k1*Q + k2] = w16^(z0 k1) * sum_{n<16} [ sum_h win[16h + n] w_L^((z0 + 16h + n) k2) ] w16^(n k1),  Q = L / 16
// i.e. one radix-16 butterfly per output residue k2 on twiddled inputs (an input-pruned FFT: the other L - 16*NH
// planes of the padded PSF are zero).  Written in digit-reversed POSITION order, like the data tile at the multiply.
template <int L, int NH, int NW, int TXP>
__device__ __forceinline__ void otf_h_tile(const float4* __restrict__ win, float4* __restrict__ hb,
                                           const float4* __restrict__ tw, const int* __restrict__ pos_s, int z0, int cp, int w)
{
    constexpr int Q = L / 16;
    const int s16 = z0 & 15;
    for (int k2 = w; k2 < Q; k2 += NW) {
        p2 r[16], i[16];
        int e = (z0 * k2) % L;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float4 v = win[n * TXP + cp];
            r[n] = make_float2(v.x, v.y);
            i[n] = make_float2(v.z, v.w);
            cmul(r[n], i[n], tw[e]);
            e += k2;
            if (e >= L) e -= L;
        }
#pragma unroll
        for (int h = 1; h < NH; ++h) {
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                const float4 v = win[(16 * h + n) * TXP + cp];
                p2 tr = make_float2(v.x, v.y), ti = make_float2(v.z, v.w);
                cmul(tr, ti, tw[e]);
                r[n] = padd(r[n], tr);
                i[n] = padd(i[n], ti);
                e += k2;
                if (e >= L) e -= L;
            }
        }
        Dft<16>::run(r, i);
#pragma unroll
        for (int k1 = 0; k1 < 16; ++k1) {
            if (s16 != 0 && k1 != 0) cmul(r[k1], i[k1], tw[((s16 * k1) & 15) * Q]);
            hb[pos_s[k1 * Q + k2] * TXP + cp] = make_float4(r[k1].x, r[k1].y, i[k1].x, i[k1].y);
        }
    }
}

// Two-stage plans whose last radix is 16: the frequencies of last-stage butterfly b are k2 + m * (L / 16) with k2 = b, i.e.
// exactly the 16 outputs k1 = m of ONE butterfly of the H derivation above.  The thread that runs the fused middle stage on
// butterfly b therefore derives its 16 H values in registers -- no H tile in shared memory, no position table, one barrier
// less -- then: last forward stage, x H x c, first inverse stage, as smid_fused_tma.
template <int L, int NH, int NW, int TXP>
__device__ __forceinline__ void smid_fused_otf16(const float4* __restrict__ win, float4* __restrict__ sm,
                                                 const float4* __restrict__ tw, int z0, int cp, int w, float c)
{
    constexpr int Q = L / 16, ITER = (Q + NW - 1) / NW;
    const int s16 = z0 & 15;
#pragma unroll
    for (int it = 0; it < ITER; ++it) {
        const int k2 = w + it * NW;
        if ((Q % NW) != 0 && k2 >= Q) break;
        p2 hr[16], hi[16];
        int e = (z0 * k2) % L;
#pragma unroll
        for (int n = 0; n < 16; ++n) {
            const float4 v = win[n * TXP + cp];
            hr[n] = make_float2(v.x, v.y);
            hi[n] = make_float2(v.z, v.w);
            cmul(hr[n], hi[n], tw[e]);
            e += k2;
            if (e >= L) e -= L;
        }
#pragma unroll
        for (int h = 1; h < NH; ++h) {
#pragma unroll
            for (int n = 0; n < 16; ++n) {
                const float4 v = win[(16 * h + n) * TXP + cp];
                p2 tr = make_float2(v.x, v.y), ti = make_float2(v.z, v.w);
                cmul(tr, ti, tw[e]);
                hr[n] = padd(hr[n], tr);
                hi[n] = padd(hi[n], ti);
                e += k2;
                if (e >= L) e -= L;
            }
        }
        Dft<16>::run(hr, hi);
        if (s16 != 0) {
#pragma unroll
            for (int k1 = 1; k1 < 16; ++k1) cmul(hr[k1], hi[k1], tw[((s16 * k1) & 15) * Q]);
        }
        p2 r[16], i[16];
        load_pairs<16>(sm, k2 * 16 * TXP + cp, TXP, r, i);
        Dft<16>::run(r, i);
#pragma unroll
        for (int m = 0; m < 16; ++m) {
            const p2 xr = pmuls(pfma(hr[m], r[m], pneg(pmul(hi[m], i[m]))), c);
            const p2 xi = pmuls(pfma(hi[m], r[m], pmul(hr[m], i[m])), c);
            r[m] = xr;
            i[m] = xi;
        }
        Dft<16>::run(i, r);
        store_pairs<16>(sm, k2 * 16 * TXP + cp, TXP, r, i);
    }
}

__device__ __forceinline__ void mbar_arrive(unsigned long long* bar)
{
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// barrier among the THREADS consumer threads only (the producer warp never joins it)
template <int THREADS>
__device__ __forceinline__ void consumer_sync()
{
    asm volatile("bar.sync 1, %0;" ::"n"(THREADS) : "memory");
}

// Warp-specialised pipeline.  THREADS consumer threads transform tiles in shared memory; one extra warp is the
// PRODUCER: its elected lane issues every TMA load and store and recycles a slot the moment its store has been read.
//   full[s]  : TMA load of slot s has landed          (producer arms it with expect_tx, the TMA unit completes it)
//   done[s]  : every consumer has finished slot s     (THREADS arrivals, after a fence.proxy.async each)
// shared memory: [NBUF][slot] | twiddles [L] float4 | full[NBUF] | done[NBUF]
//   slot = data tile (+ H tile for MODE 2), each L * TXP float4
//   MODE 3: MODE 2 with the PSF-spectrum tile derived ON THE FLY from the 16*NH window planes (h_map = window buffer):
//   slot = data tile + window rows; one shared H tile; no image-sized PSF spectrum is read (or exists)
template <int MODE, class P, int THREADS, int NBUF, int TXP, bool GROUPS, int NH = 0>
__global__ void __launch_bounds__(THREADS + 32, 1)
    col_tma_kernel(const __grid_constant__ CUtensorMap nat_map, const __grid_constant__ CUtensorMap perm_map,
                   const __grid_constant__ CUtensorMap h_map, TmaArgs a)
{
    constexpr int L = P::L, NW = THREADS / TXP, TILE = L * TXP;
    constexpr int WIN = 16 * NH * TXP;
    constexpr int SLOT = MODE == 2 ? 2 * TILE : (MODE == 3 ? TILE + WIN : TILE);
    constexpr unsigned SLOT_BYTES = (unsigned)SLOT * sizeof(float4);
    extern __shared__ __align__(128) float4 smem[];
    float4* bufs = smem;
    // MODE 3 with a two-stage plan whose last radix is 16: H stays in registers (smid_fused_otf16), no H tile, no table.
    // Measured: L = 384 0.156 -> 0.148 ms; L = 256, which runs two CTAs per SM either way, 0.132 -> 0.150 ms (196 registers,
    // the H derivation and the data butterflies no longer overlap): only the longer length uses it.
    constexpr bool HREG = MODE == 3 && P::ns == 2 && P::RL == 16 && P::L >= 384;
    float4* hbuf = bufs + (size_t)NBUF * SLOT;                  // MODE 3 only
    float4* tw = hbuf + ((MODE == 3 && !HREG) ? TILE : 0);
    int* pos_s = reinterpret_cast<int*>(tw + L);                // MODE 3 only
    unsigned long long* full = reinterpret_cast<unsigned long long*>(pos_s + ((MODE == 3 && !HREG) ? L : 0));
    unsigned long long* done = full + NBUF;

    const int t = threadIdx.x;
    const int stride = (int)gridDim.x;

    if (t == 0) {
#pragma unroll
        for (int i = 0; i < NBUF; ++i) {
            mbar_init(full + i, 1);
            mbar_init(done + i, THREADS);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_launch_dependents();
    load_twiddles(tw, a.tw, L);
    if constexpr (MODE == 3 && !HREG)
        for (int k = t; k < L; k += THREADS + 32) {   // position of frequency k for THIS kernel's radix sequence
            const int m0 = k % P::R0, q = k / P::R0;
            int pk;
            if (P::ns == 2) pk = m0 * P::R1 + q;
            else if (P::ns == 3) pk = (m0 * P::R1 + q % P::R1) * P::R2 + q / P::R1;
            else pk = ((m0 * P::R1 + q % P::R1) * P::R2 + (q / P::R1) % P::R2) * P::R3 + q / (P::R1 * P::R2);
            pos_s[k] = pk;
        }
    __syncthreads();
    pdl_wait();   // everything below reads what the previous pass wrote

    if (t >= THREADS) {
        // ---------------- producer warp ----------------
        if (t != THREADS) return;
        auto issue_load = [&](int tile, int slot) {
            if (a.reverse) tile = a.totalTiles - 1 - tile;
            const int gi = tile / a.tilesPerGroup;
            const int tt = tile - gi * a.tilesPerGroup;
            float4* dst = bufs + (size_t)slot * SLOT;
            mbar_expect_tx(full + slot, SLOT_BYTES);
            for (int r0 = 0; r0 < L; r0 += a.boxRows)
                tma_load_3d(dst + (size_t)r0 * TXP, &nat_map, tt * 4 * TXP, r0, gi, full + slot);
            if (MODE == 2) perm_load<P::ns, GROUPS>(dst + TILE, &h_map, tt * 4 * TXP, gi, full + slot);
            if (MODE == 3) tma_load_3d(dst + TILE, &h_map, tt * 4 * TXP, 0, 0, full + slot);
        };
#pragma unroll
        for (int k = 0; k < NBUF; ++k) {
            const int tile = (int)blockIdx.x + k * stride;
            if (tile < a.totalTiles) issue_load(tile, k);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < a.totalTiles; tile += stride, ++it) {
            const int slot = it % NBUF;
            float4* sm = bufs + (size_t)slot * SLOT;
            mbar_wait(done + slot, (unsigned)((it / NBUF) & 1));
            const int tq = a.reverse ? a.totalTiles - 1 - tile : tile;
            const int gi = tq / a.tilesPerGroup;
            const int tt = tq - gi * a.tilesPerGroup;
            if constexpr (MODE >= 2) {
                for (int r0 = 0; r0 < L; r0 += a.boxRows)
                    tma_store_3d(&nat_map, tt * 4 * TXP, r0, gi, sm + (size_t)r0 * TXP);
            } else {
                perm_store<P::ns, GROUPS>(&perm_map, tt * 4 * TXP, gi, sm);
            }
            tma_commit();
            const int next = tile + NBUF * stride;
            if (next < a.totalTiles) {
                tma_wait_read<0>();   // the store has read the slot: refill it
                issue_load(next, slot);
            }
        }
        tma_wait_read<0>();
        return;
    }

    // ---------------- consumers ----------------
    const int cp = t % TXP, w = t / TXP;
    // second-stage twiddles in registers (see sstage_regtw): plans with >= 3 stages whose stage-2 j is tile-invariant
    constexpr int S2 = (L / P::R0) / P::R1;
    constexpr bool REG2 = (P::ns >= 3) && (S2 > 1) && (NW % (S2 > 1 ? S2 : 1) == 0) && (P::R1 <= 16);
    float tc2[REG2 ? P::R1 - 1 : 1], ts2[REG2 ? P::R1 - 1 : 1];
    if constexpr (REG2) {
        const int j2 = w % S2;
#pragma unroll
        for (int m = 1; m < P::R1; ++m) {
            const float4 tv = tw[j2 * m * P::R0];   // w_L^(j m tstep), tstep = L / Li = R0
            tc2[m - 1] = tv.x;
            ts2[m - 1] = tv.z;
        }
    }
    int it = 0;
    for (int tile = blockIdx.x; tile < a.totalTiles; tile += stride, ++it) {
        const int slot = it % NBUF;
        float4* sm = bufs + (size_t)slot * SLOT;
        mbar_wait(full + slot, (unsigned)((it / NBUF) & 1));

        if constexpr (MODE < 2) {
            // forward stage sequence, in place; MODE 1 exchanges re/im on the way in and on the way out
            if constexpr (P::ns == 2) {
                sstage<P::R0, L, L, NW, false, TXP, MODE == 1>(sm, tw, cp, w);
                consumer_sync<THREADS>();
                if constexpr (MODE == 1) sstage_swapout<P::R1, L, L / P::R0, NW, TXP>(sm, tw, cp, w);
                else sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
            } else if constexpr (P::ns == 3) {
                sstage<P::R0, L, L, NW, false, TXP, MODE == 1>(sm, tw, cp, w);
                consumer_sync<THREADS>();
                if constexpr (REG2) sstage_regtw<P::R1, L, L / P::R0, NW, false, TXP>(sm, cp, w, tc2, ts2);
                else sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
                if constexpr (MODE == 1) sstage_swapout<P::R2, L, L / (P::R0 * P::R1), NW, TXP>(sm, tw, cp, w);
                else sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, TXP>(sm, tw, cp, w);
            } else {
                sstage<P::R0, L, L, NW, false, TXP, MODE == 1>(sm, tw, cp, w);
                consumer_sync<THREADS>();
                if constexpr (REG2) sstage_regtw<P::R1, L, L / P::R0, NW, false, TXP>(sm, cp, w, tc2, ts2);
                else sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
                sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
                if constexpr (MODE == 1) sstage_swapout<P::R3, L, L / (P::R0 * P::R1 * P::R2), NW, TXP>(sm, tw, cp, w);
                else sstage<P::R3, L, L / (P::R0 * P::R1 * P::R2), NW, false, TXP>(sm, tw, cp, w);
            }
        } else {
            sstage<P::R0, L, L, NW, false, TXP>(sm, tw, cp, w);
            if constexpr (MODE == 3 && !HREG) otf_h_tile<L, NH, NW, TXP>(sm + TILE, hbuf, tw, pos_s, a.z0, cp, w);
            consumer_sync<THREADS>();
            if constexpr (P::ns >= 3) {
                if constexpr (REG2) sstage_regtw<P::R1, L, L / P::R0, NW, false, TXP>(sm, cp, w, tc2, ts2);
                else sstage<P::R1, L, L / P::R0, NW, false, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
            }
            if constexpr (P::ns >= 4) {
                sstage<P::R2, L, L / (P::R0 * P::R1), NW, false, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
            }
            if constexpr (HREG) smid_fused_otf16<L, NH, NW, TXP>(sm + TILE, sm, tw, a.z0, cp, w, a.scale);
            else smid_fused_tma<P::RL, L, NW, TXP>(MODE == 3 ? hbuf : sm + TILE, sm, cp, w, a.scale);
            consumer_sync<THREADS>();
            if constexpr (P::ns >= 4) {
                sstage<P::R2, L, P::R2 * P::R3, NW, true, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
            }
            if constexpr (P::ns >= 3) {
                if constexpr (REG2) sstage_regtw<P::R1, L, P::R1 * P::R2 * P::R3, NW, true, TXP>(sm, cp, w, tc2, ts2);
                else sstage<P::R1, L, P::R1 * P::R2 * P::R3, NW, true, TXP>(sm, tw, cp, w);
                consumer_sync<THREADS>();
            }
            sstage<P::R0, L, L, NW, true, TXP>(sm, tw, cp, w);   // natural order again
        }
        fence_async_smem();        // my writes to the tile -> visible to the TMA store
        mbar_arrive(done + slot);
    }
}

int sm_count_of_current_device()
{
    static std::mutex mu;
    static int cache[64] = {0};
    int dev = 0;
    cudaGetDevice(&dev);
    std::lock_guard<std::mutex> lock(mu);
    if (dev < 0 || dev >= 64) return 148;
    if (cache[dev] == 0) {
        int v = 148;
        cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev);
        cache[dev] = v;
    }
    return cache[dev];
}

// MODES: which kernels get compiled for this configuration (1 = plain passes, 2 = fused z pass, 3 = both)
template <class P, int THREADS, int NBUF, int TXP, int MODES = 3>
bool run_col_tma(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    if (mode == 2 ? !(MODES & 2) : !(MODES & 1)) return false;
    constexpr int L = P::L;
    if (a.split || a.splitPeers || a.rowMask || a.groupList || a.winSlot) return false;
    if (!encode_fn()) return false;
    const int boxFloats = 4 * TXP;
    const int tpg = (a.rowLen * 2 + boxFloats - 1) / boxFloats;
    const long long total = ngroups * tpg;
    if (total == 0) return true;
    if (total > 0x7fffffffLL) return false;
    if (ngroups > 1 && (a.groupStride % 2) != 0) return false;            // 16-byte global strides
    if ((a.stride % 2) != 0 || (reinterpret_cast<uintptr_t>(a.data) & 15) != 0) return false;
    int boxRows = std::min(L, 256);
    while (L % boxRows) --boxRows;
    const size_t tile = (size_t)L * TXP * sizeof(float4);
    const size_t smem = (size_t)NBUF * tile * (mode == 2 ? 2 : 1) + (size_t)L * sizeof(float4) + 2 * NBUF * sizeof(unsigned long long);
    if (smem > (size_t)kMaxDynSmem) return false;

    CUtensorMap nat, perm, hmap;
    if (!encode(&nat, a.data, natural_spec(a, ngroups, L, boxRows, boxFloats))) return false;
    MapSpec ps;
    if (!permuted_spec<P>(a, ngroups, boxFloats, ps)) return false;
    if (!encode(&perm, a.data, ps)) return false;
    hmap = perm;
    if (mode == 2) {
        if ((reinterpret_cast<uintptr_t>(a.H) & 15) != 0) return false;
        if (!encode(&hmap, const_cast<float2*>(a.H), ps)) return false;
    }
    // plain y passes walk the planes from the last to the first: the x pass before a forward y pass has just written
    // the last planes (still in the 126 MB L2), and the x pass after an inverse y pass starts with the first ones
    static const int rev_env = env_int("FCB200_TMA_REVERSE", 1);
    const int reverse = (rev_env == 1 && ngroups > 1 && mode != 2) ? 1 : (rev_env == 2 ? 1 : 0);
    TmaArgs ta{a.P.tw, tpg, (int)total, boxRows, a.scale, reverse, nullptr, 0};
    auto go = [&](auto kernel) {
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;
        FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS + 32, smem));
        const int grid = (int)std::min<long long>(total, (long long)sm_count_of_current_device() * std::max(1, per_sm));
        launch_pdl(a.pdl != 0, kernel, dim3(grid), dim3(THREADS + 32), smem, st, nat, perm, hmap, ta);
        FC_CUDA_KERNEL();
    };
    if (mode == 2) {
        if (ngroups > 1) return false;   // the fused pass runs along z: one group
        if constexpr ((MODES & 2) != 0) go(col_tma_kernel<2, P, THREADS, NBUF, TXP, false>);
    } else if constexpr ((MODES & 1) != 0) {
        if (ngroups > 1) {
            if (mode == 0) go(col_tma_kernel<0, P, THREADS, NBUF, TXP, true>);
            else go(col_tma_kernel<1, P, THREADS, NBUF, TXP, true>);
        } else {
            if (mode == 0) go(col_tma_kernel<0, P, THREADS, NBUF, TXP, false>);
            else go(col_tma_kernel<1, P, THREADS, NBUF, TXP, false>);
        }
    }
    return true;
}

// ------------------------------------------------------------------------------------------------------------------
// The same pipeline for the PLAIN passes of long lengths WITHOUT a compile-time plan (7-smooth lengths of two or three
// register stages: 400, 480, 600, 640 ...): the producer warp is identical -- the tensor maps are built on the host from
// the plan's run-time radices -- and the consumers run the run-time-radix stages of fft_engine.cuh (one switch per stage)
// in place.  The one-tile-per-CTA kernels these lengths ran on are latency-bound (ncu, 400^3 y pass: 16 warps per SM, 33 %
// SM throughput, 3.1 TB/s).  Measured (profiles/r02_tma_dyn_ab.jsonl): y passes 400: 0.163 -> 0.150 ms, 480: 0.274 -> 0.238,
// 600: 0.183 -> 0.138, 640: 0.172 -> 0.128; shorter lengths (160, 200) lose and a fused variant (forward stages, multiply,
// inverse stages: 2 ns + 1 round trips with two slots) lost on every length, so those stay where they were.
//   MODE 0 forward: natural-order load, digit-reversing store.   MODE 1 inverse (mirrored stages): digit-reversing LOAD,
//   natural-order store.
constexpr int kDynMaxThreads = 352;
constexpr int kDynMaxBuf = 4;

template <int MODE, bool GROUPS>
__global__ void __launch_bounds__(kDynMaxThreads + 32, 1)
    col_tma_dyn_kernel(const __grid_constant__ CUtensorMap nat_map, const __grid_constant__ CUtensorMap perm_map, TmaArgs a,
                       AxisPlanDev P, int nbuf)
{
    constexpr int TXP = 8;
    const int L = P.L, ns = P.ns, TILE = L * TXP;
    const int SLOT = TILE;
    const unsigned slot_bytes = (unsigned)SLOT * sizeof(float4);
    extern __shared__ __align__(128) float4 smem[];
    float4* bufs = smem;
    float4* tw = bufs + (size_t)nbuf * SLOT;
    unsigned long long* full = reinterpret_cast<unsigned long long*>(tw + L);
    unsigned long long* done = full + kDynMaxBuf;

    const int t = threadIdx.x;
    const int nthreads = (int)blockDim.x - 32;   // consumers
    const int stride = (int)gridDim.x;

    if (t == 0) {
        for (int i = 0; i < nbuf; ++i) {
            mbar_init(full + i, 1);
            mbar_init(done + i, nthreads);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    pdl_launch_dependents();
    load_twiddles(tw, a.tw, L);
    __syncthreads();
    pdl_wait();

    if (t >= nthreads) {
        // ---------------- producer warp ----------------
        if (t != nthreads) return;
        auto perm_ld = [&](void* dst, const CUtensorMap* map, int c0, int gi, unsigned long long* bar) {
            if (ns == 2) {
                if (GROUPS) tma_load_4d(dst, map, c0, 0, 0, gi, bar);
                else tma_load_3d(dst, map, c0, 0, 0, bar);
            } else if (ns == 3) {
                if (GROUPS) tma_load_5d(dst, map, c0, 0, 0, 0, gi, bar);
                else tma_load_4d(dst, map, c0, 0, 0, 0, bar);
            } else {
                tma_load_5d(dst, map, c0, 0, 0, 0, 0, bar);
            }
        };
        auto perm_st = [&](const CUtensorMap* map, int c0, int gi, const void* src) {
            if (ns == 2) {
                if (GROUPS) tma_store_4d(map, c0, 0, 0, gi, src);
                else tma_store_3d(map, c0, 0, 0, src);
            } else if (ns == 3) {
                if (GROUPS) tma_store_5d(map, c0, 0, 0, 0, gi, src);
                else tma_store_4d(map, c0, 0, 0, 0, src);
            } else {
                tma_store_5d(map, c0, 0, 0, 0, 0, src);
            }
        };
        auto issue_load = [&](int tile, int slot) {
            if (a.reverse) tile = a.totalTiles - 1 - tile;
            const int gi = tile / a.tilesPerGroup;
            const int tt = tile - gi * a.tilesPerGroup;
            float4* dst = bufs + (size_t)slot * SLOT;
            mbar_expect_tx(full + slot, slot_bytes);
            if (MODE == 1) {
                perm_ld(dst, &perm_map, tt * 4 * TXP, gi, full + slot);
            } else {
                for (int r0 = 0; r0 < L; r0 += a.boxRows)
                    tma_load_3d(dst + (size_t)r0 * TXP, &nat_map, tt * 4 * TXP, r0, gi, full + slot);
            }
        };
        for (int k = 0; k < nbuf; ++k) {
            const int tile = (int)blockIdx.x + k * stride;
            if (tile < a.totalTiles) issue_load(tile, k);
        }
        int it = 0;
        for (int tile = blockIdx.x; tile < a.totalTiles; tile += stride, ++it) {
            const int slot = it % nbuf;
            float4* sm = bufs + (size_t)slot * SLOT;
            mbar_wait(done + slot, (unsigned)((it / nbuf) & 1));
            const int tq = a.reverse ? a.totalTiles - 1 - tile : tile;
            const int gi = tq / a.tilesPerGroup;
            const int tt = tq - gi * a.tilesPerGroup;
            if (MODE == 0) {
                perm_st(&perm_map, tt * 4 * TXP, gi, sm);
            } else {
                for (int r0 = 0; r0 < L; r0 += a.boxRows)
                    tma_store_3d(&nat_map, tt * 4 * TXP, r0, gi, sm + (size_t)r0 * TXP);
            }
            tma_commit();
            const int next = tile + nbuf * stride;
            if (next < a.totalTiles) {
                tma_wait_read<0>();
                issue_load(next, slot);
            }
        }
        tma_wait_read<0>();
        return;
    }

    // ---------------- consumers ----------------
    const int cp = t % TXP, w = t / TXP, W = nthreads / TXP;
    auto csync = [&] { asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory"); };
    int it = 0;
    for (int tile = blockIdx.x; tile < a.totalTiles; tile += stride, ++it) {
        const int slot = it % nbuf;
        float4* sm = bufs + (size_t)slot * SLOT;
        float4* none = nullptr;
        mbar_wait(full + slot, (unsigned)((it / nbuf) & 1));
        if (MODE == 0) {
            int Li = L;
#pragma unroll
            for (int s = 0; s < 4; ++s) {   // (unrolled: the radices stay in registers)
                if (s < ns) {
                    const int R = P.radix[s];
                    stage_dispatch<false, false>(R, sm, none, tw, L, Li, cp, w, W, TXP, true);
                    Li /= R;
                    if (s + 1 < ns) csync();
                }
            }
        }
        if (MODE == 1) {
            int Li = 1;
#pragma unroll
            for (int s = 3; s >= 0; --s) {
                if (s < ns) {
                    const int R = P.radix[s];
                    Li *= R;
                    stage_dispatch<true, false>(R, sm, none, tw, L, Li, cp, w, W, TXP, true);
                    if (s > 0) csync();
                }
            }
        }
        fence_async_smem();
        mbar_arrive(done + slot);
    }
}

bool permuted_spec_dyn(const ColArgs& a, long long ngroups, int boxFloats, MapSpec& s)
{
    const int ns = a.P.ns;
    const bool with_groups = ngroups > 1;
    s.rank = 1 + ns + (with_groups ? 1 : 0);
    if (ns < 2 || s.rank > 5) return false;
    s.dims[0] = (cuuint64_t)a.rowLen * 2;
    s.box[0] = (cuuint32_t)boxFloats;
    const cuuint64_t row = (cuuint64_t)a.stride * sizeof(float2);
    cuuint64_t kstride[4], acc = 1;
    for (int i = 0; i < ns; ++i) {
        if (a.P.radix[i] > 256) return false;
        kstride[i] = acc;
        acc *= (cuuint64_t)a.P.radix[i];
    }
    for (int d = 0; d < ns; ++d) {
        const int st = ns - 1 - d;
        s.dims[1 + d] = (cuuint64_t)a.P.radix[st];
        s.strides[d] = kstride[st] * row;
        s.box[1 + d] = (cuuint32_t)a.P.radix[st];
    }
    if (with_groups) {
        s.dims[1 + ns] = (cuuint64_t)ngroups;
        s.strides[ns] = (cuuint64_t)a.groupStride * sizeof(float2);
        s.box[1 + ns] = 1;
    }
    return true;
}

bool run_col_tma_dyn(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const int L = a.P.L;
    static const int min_len = env_int("FCB200_TMA_DYN_MINLEN", 384);
    if (mode == 2 || a.P.generic || a.P.big || a.P.ns < 2 || a.P.ns > 4 || L < min_len) return false;
    if (a.split || a.splitPeers || a.rowMask || a.groupList || a.winSlot || a.txp != 8) return false;
    if (!encode_fn()) return false;
    const int boxFloats = 32;
    const int tpg = (a.rowLen * 2 + boxFloats - 1) / boxFloats;
    const long long total = ngroups * tpg;
    if (total == 0) return true;
    if (total > 0x7fffffffLL) return false;
    if (ngroups > 1 && (a.groupStride % 2) != 0) return false;
    if ((a.stride % 2) != 0 || (reinterpret_cast<uintptr_t>(a.data) & 15) != 0) return false;
    int boxRows = std::min(L, 256);
    while (L % boxRows) --boxRows;
    if (L / boxRows > 16) return false;   // (a prime-ish length would need one box per few rows)
    const size_t slot = (size_t)L * 8 * sizeof(float4);
    const size_t fixed = (size_t)L * sizeof(float4) + 2 * kDynMaxBuf * sizeof(unsigned long long);
    int nbuf = (int)std::min<size_t>(3, ((size_t)kMaxDynSmem - fixed) / slot);
    if (nbuf < 2) return false;
    const size_t smem = (size_t)nbuf * slot + fixed;
    // one consumer per butterfly of the stage with the most butterflies
    int most = 0;
    for (int s = 0; s < a.P.ns; ++s) most = std::max(most, L / a.P.radix[s]);
    const int threads = std::min(kDynMaxThreads, std::max(64, ((most * 8 + 31) / 32) * 32));

    CUtensorMap nat, perm;
    if (!encode(&nat, a.data, natural_spec(a, ngroups, L, boxRows, boxFloats))) return false;
    MapSpec ps;
    if (!permuted_spec_dyn(a, ngroups, boxFloats, ps)) return false;
    if (!encode(&perm, a.data, ps)) return false;
    static const int rev_env = env_int("FCB200_TMA_REVERSE", 1);
    const int reverse = (rev_env == 1 && ngroups > 1) ? 1 : (rev_env == 2 ? 1 : 0);
    TmaArgs ta{a.P.tw, tpg, (int)total, boxRows, a.scale, reverse, nullptr, 0};
    auto go = [&](auto kernel) {
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = (int)std::min<long long>(total, (long long)sm_count_of_current_device());
        launch_pdl(a.pdl != 0, kernel, dim3(grid), dim3(threads + 32), smem, st, nat, perm, ta, a.P, nbuf);
        FC_CUDA_KERNEL();
    };
    if (ngroups > 1) {
        if (mode == 0) go(col_tma_dyn_kernel<0, true>);
        else go(col_tma_dyn_kernel<1, true>);
    } else {
        if (mode == 0) go(col_tma_dyn_kernel<0, false>);
        else go(col_tma_dyn_kernel<1, false>);
    }
    return true;
}

// fused z pass with the PSF spectrum derived on the fly from the window planes in a.H ([winPlanes][ny][xcp])
template <class P, int THREADS, int NBUF, int TXP>
bool run_col_otf_tma(const ColArgs& a, long long ngroups, int z0, cudaStream_t st, bool probe)
{
    constexpr int L = P::L;
    static_assert(L % 16 == 0, "on-the-fly PSF spectrum needs L % 16 == 0");
    // natural order in and out and no planner table: the kernel's radix sequence need not be the planner's
    if (a.P.L != L) return false;
    if (a.split || a.splitPeers || a.rowMask || a.groupList || ngroups != 1) return false;
    const int nh = a.winPlanes / 16;
    if (a.winPlanes % 16 != 0 || (nh != 1 && nh != 2 && nh != 4) || a.winPlanes > L) return false;
    if (!encode_fn()) return false;
    if ((a.stride % 2) != 0) return false;
    const int boxFloats = 4 * TXP;
    const int tpg = (a.rowLen * 2 + boxFloats - 1) / boxFloats;
    int boxRows = std::min(L, 256);
    while (L % boxRows) --boxRows;
    const size_t tile = (size_t)L * TXP * sizeof(float4), win = (size_t)a.winPlanes * TXP * sizeof(float4);
    constexpr bool hreg = P::ns == 2 && P::RL == 16 && P::L >= 384;   // H in registers: no H tile, no position table (col_tma_kernel)
    const size_t smem = (size_t)NBUF * (tile + win) + (hreg ? 0 : tile + (size_t)L * sizeof(int)) + (size_t)L * sizeof(float4) +
                        2 * NBUF * sizeof(unsigned long long);
    if (smem > (size_t)kMaxDynSmem) return false;
    if (probe) return true;
    if (tpg == 0) return true;
    if ((reinterpret_cast<uintptr_t>(a.data) & 15) != 0 || (reinterpret_cast<uintptr_t>(a.H) & 15) != 0) return false;
    CUtensorMap nat, wmap;
    if (!encode(&nat, a.data, natural_spec(a, 1, L, boxRows, boxFloats))) return false;
    MapSpec ws;
    ws.rank = 3;
    ws.dims[0] = (cuuint64_t)a.rowLen * 2;
    ws.dims[1] = (cuuint64_t)a.winPlanes;
    ws.dims[2] = 1;
    ws.strides[0] = (cuuint64_t)a.stride * sizeof(float2);
    ws.strides[1] = ws.strides[0] * (cuuint64_t)a.winPlanes;
    ws.box[0] = (cuuint32_t)boxFloats;
    ws.box[1] = (cuuint32_t)a.winPlanes;
    ws.box[2] = 1;
    if (!encode(&wmap, const_cast<float2*>(a.H), ws)) return false;
    TmaArgs ta{a.P.tw, tpg, tpg, boxRows, a.scale, 0, nullptr, z0};
    auto go = [&](auto kernel) {
        FC_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        int per_sm = 1;   // two-slot configurations of the short lengths fit twice: the pass is bound by its butterflies
        FC_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, THREADS + 32, smem));
        const int grid = std::min(tpg, sm_count_of_current_device() * std::max(1, per_sm));
        launch_pdl(a.pdl != 0, kernel, dim3(grid), dim3(THREADS + 32), smem, st, nat, nat, wmap, ta);
        FC_CUDA_KERNEL();
    };
    if (nh == 1) go(col_tma_kernel<3, P, THREADS, NBUF, TXP, false, 1>);
    else if (nh == 2) go(col_tma_kernel<3, P, THREADS, NBUF, TXP, false, 2>);
    else go(col_tma_kernel<3, P, THREADS, NBUF, TXP, false, 4>);
    return true;
}

}  // namespace

bool launch_col_otf_tma(const ColArgs& a, long long ngroups, int z0, cudaStream_t st, bool probe)
{
    const int on = env_int("FCB200_TMA", 1);
    static const int otf_on = env_int("FCB200_OTF_TMA", 1);
    if (!on || !otf_on || !static_enabled() || a.txp != 8) return false;
    static const long long max_stride = (long long)env_int("FCB200_TMA_MAXSTRIDE_KB", 2048) << 10;
    if (a.stride * (long long)sizeof(float2) > max_stride) return false;
    // L = 256 runs as (16,16) -- one shared-memory round trip instead of two next to the H derivation -- with two slots,
    // so that two CTAs fit an SM (the pass is bound by its butterflies): C3 0.195 -> 0.134 ms, below the pass that
    // reads a materialised spectrum (0.142 ms)
    // L = 384 as (24,16) for the same reason (one CTA per SM, three slots): 384^3 0.182 -> 0.156 ms ((16,24): 0.164)
    return run_col_otf_tma<P256, 128, 2, 8>(a, ngroups, z0, st, probe) ||
           run_col_otf_tma<SPlan<384, 24, 16>, 192, 3, 8>(a, ngroups, z0, st, probe) ||
           run_col_otf_tma<P512, 256, 4, 4>(a, ngroups, z0, st, probe) || run_col_otf_tma<P448, 512, 2, 8>(a, ngroups, z0, st, probe);
}

// FCB200_TMA: 0 = off, 1 (default) = the configurations measured to win, 2 = every configuration compiled below
bool launch_col_tma(const ColArgs& a, int mode, long long ngroups, cudaStream_t st)
{
    const int on = env_int("FCB200_TMA", 1);   // read on every call: the tests switch it within one process
    if (!on || !static_enabled()) return false;
    // Rows further apart than a 2 MiB page (z passes of very large planes) put every row of a tile on its own page;
    // the TMA path loses there (1024x1024x512, z stride 4.2 MB: fused pass 1.55 -> 2.8 ms), the register kernels cope.
    static const long long max_stride = (long long)env_int("FCB200_TMA_MAXSTRIDE_KB", 2048) << 10;
    if (a.stride * (long long)sizeof(float2) > max_stride) return false;
    if (a.txp == 4) {   // L = 2048: 32-byte row segments keep three 64 KB tiles in flight (2048x2048 y passes 1.29 -> 1.09 ms)
        if (plan_matches<P2048>(a.P) && mode != 2) return run_col_tma<P2048, 256, 3, 2>(a, mode, ngroups, st);
        return false;
    }
    if (a.txp != 8) return false;
    // measured (profiles/r02_tma_ab.jsonl): C3 y passes 0.117 -> 0.100 ms, fused z 0.163 -> 0.144 ms; 384^3 y 0.098 -> 0.081,
    // fused z 0.170 -> 0.143; 1024^2 y passes (64-byte rows) 1.05 -> 0.75 ms; 512^3 fused z (L = 512, 64-byte rows) 0.401 -> 0.367
    static const int t512 = env_int("FCB200_TMA_T512", 512);
    if (plan_matches<P512>(a.P) && mode != 2 && t512 == 256) return run_col_tma<P512, 256, 3, 8>(a, mode, ngroups, st);
    if (plan_matches<P512>(a.P) && mode != 2) return run_col_tma<P512, 512, 3, 8>(a, mode, ngroups, st);
    if (plan_matches<P512>(a.P) && mode == 2) return run_col_tma<P512, 256, 3, 4>(a, mode, ngroups, st);
    if (plan_matches<P256b>(a.P) && mode == 2) return run_col_tma<P256b, 256, 3, 8>(a, mode, ngroups, st);
    if (plan_matches<P256>(a.P)) return run_col_tma<P256, 128, 3, 8>(a, mode, ngroups, st);
    if (plan_matches<P384>(a.P) && mode != 2) return run_col_tma<P384, 192, 3, 8>(a, mode, ngroups, st);
    if (plan_matches<P384>(a.P) && mode == 2) return run_col_tma<P384, 192, 2, 8>(a, mode, ngroups, st);
    if (plan_matches<P1024>(a.P) && mode != 2) return run_col_tma<P1024, 256, 3, 4>(a, mode, ngroups, st);
    // 7-smooth extents of the caller-padded configurations: 560^2 y passes 0.164 -> 0.136 ms; the two-stage plans
    // (300, 420, 270, 448 on the y axis) with one consumer per butterfly of the larger stage
    if (plan_matches<P560>(a.P) && mode != 2) return run_col_tma<P560, 320, 3, 8, 1>(a, mode, ngroups, st);
    if (plan_matches<P300>(a.P) && mode != 2) return run_col_tma<P300, 160, 3, 8, 1>(a, mode, ngroups, st);
    if (plan_matches<P300>(a.P) && mode == 2) return run_col_tma<P300, 160, 2, 8, 2>(a, mode, ngroups, st);
    if (plan_matches<P420>(a.P) && mode != 2) return run_col_tma<P420, 192, 3, 8, 1>(a, mode, ngroups, st);
    if (plan_matches<P420>(a.P) && mode == 2) return run_col_tma<P420, 192, 2, 8, 2>(a, mode, ngroups, st);
    if (plan_matches<P270>(a.P) && mode != 2) return run_col_tma<P270, 160, 3, 8, 1>(a, mode, ngroups, st);
    if (plan_matches<P270>(a.P) && mode == 2) return run_col_tma<P270, 160, 2, 8, 2>(a, mode, ngroups, st);
    if (plan_matches<P448y>(a.P) && mode != 2) return run_col_tma<P448y, 224, 3, 8, 1>(a, mode, ngroups, st);
    // plain passes of long lengths without a compile-time plan: the same pipeline with run-time radices
    // (FCB200_TMA_DYN=0: the one-tile-per-CTA kernels)
    static const int dyn = env_int("FCB200_TMA_DYN", 1);
    if (dyn && !col_static_has_plan(a.P) && run_col_tma_dyn(a, mode, ngroups, st)) return true;
    if (on >= 2) {
        if (plan_matches<P448>(a.P) && mode != 2) return run_col_tma<P448, 512, 3, 8, 1>(a, mode, ngroups, st);
        if (plan_matches<P1024>(a.P) && mode == 2) return run_col_tma<P1024, 128, 3, 2>(a, mode, ngroups, st);
    }
    return false;
}

}  // namespace fcb200
